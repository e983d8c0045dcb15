"""Replay memory: the Buffer tensor store, uniform retrieval and class-balanced sampling.

Mirrors the reference surface (utils/buffer/buffer.py:8-41; utils/buffer/buffer_utils.py:9-26,
74-160) with a different split of work:
  * images and labels live on the GPU as in the reference ([mem,3,H,W] fp32, [mem] int64);
  * every *decision* about indices is taken on the host from a numpy mirror of the labels
    (no .item()/.tolist() round trips per sample: the reference does ~5000 of them to rebuild
    its class cache, buffer_utils.py:156-160);
  * rows move with the gather/scatter kernels of csrc/misc.cu.
"""
import numpy as np
import torch

from . import ops

input_size_match = {      # utils/setup_elements.py:11-17
    'cifar100': [3, 32, 32], 'cifar10': [3, 32, 32], 'core50': [3, 128, 128],
    'mini_imagenet': [3, 84, 84], 'openloris': [3, 50, 50],
}
n_classes = {             # utils/setup_elements.py:20-26
    'cifar100': 100, 'cifar10': 10, 'core50': 50, 'mini_imagenet': 100, 'openloris': 69,
}


# --------------------------------------------------------------------------- host index logic
def class_balanced_indices(labels, n_valid, n_smp_cls, excl_mask=None, rng=None):
    """Up to n_smp_cls uniformly random slots of every class present in labels[:n_valid],
    skipping slots flagged in excl_mask.  Vectorised restatement of
    ClassBalancedRandomSampling.sample (buffer_utils.py:81-121): one random key per slot, sort
    by (class, key), keep the first n of every class.  Output is class-major (ascending class
    id), random inside a class."""
    rng = np.random if rng is None else rng
    if n_smp_cls <= 0 or n_valid <= 0:
        return np.zeros(0, dtype=np.int64)
    valid = np.arange(n_valid) if excl_mask is None else np.flatnonzero(~excl_mask[:n_valid])
    if valid.size == 0:
        return np.zeros(0, dtype=np.int64)
    lab = labels[valid]
    order = np.argsort(lab + rng.random(valid.size), kind='stable')
    lab_sorted = lab[order]
    first = np.flatnonzero(np.r_[True, lab_sorted[1:] != lab_sorted[:-1]])
    counts = np.diff(np.r_[first, lab_sorted.size])
    rank = np.arange(lab_sorted.size) - np.repeat(first, counts)
    return valid[order[rank < n_smp_cls]].astype(np.int64)


def uniform_indices(n_filled, num_retrieve, excl_indices=None):
    """np.random.choice without replacement over the filled slots minus excl_indices -- the same
    numpy calls, in the same order, as random_retrieve (buffer_utils.py:9-17), so a seeded
    numpy stream yields the reference's indices."""
    filled = np.arange(n_filled)
    excl = [] if excl_indices is None else list(excl_indices)
    valid = np.setdiff1d(filled, np.array(excl))
    num_retrieve = min(num_retrieve, valid.shape[0])
    return np.random.choice(valid, num_retrieve, replace=False).astype(np.int64)


class _Pinned:
    """Reusable pinned host staging for small index / label uploads."""

    def __init__(self):
        self.buf = None

    def upload(self, arr, device):
        arr = np.ascontiguousarray(arr, dtype=np.int64)
        n = arr.size
        if device.type != 'cuda':
            return torch.from_numpy(arr.copy())
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(max(256, 2 * n), dtype=torch.int64).pin_memory()
        # a fresh pinned slice per call would race with an in-flight copy; rotate through the buffer
        self.buf[:n].copy_(torch.from_numpy(arr))
        out = torch.empty(n, dtype=torch.int64, device=device)
        out.copy_(self.buf[:n], non_blocking=False)
        return out


def to_device_i64(arr, device):
    """Small host int64 array -> device tensor."""
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64))
    return t.to(device) if torch.device(device).type == 'cuda' else t


# --------------------------------------------------------------------------- class-balanced sampler
class ClassBalancedRandomSampling:
    """Class-level state like the reference (buffer_utils.py:74-79): two buffers in one process
    share it and the ASER plugin constructors reset it (aser_retrieve.py:19, aser_update.py:20)."""
    class_index_cache = None     # dict class -> set(slot)   (kept for parity with the reference's API)
    class_num_cache = None       # np.int64 [num_class]
    labels_host = None           # np.int64 [mem] mirror the sampler draws from
    n_valid = 0

    @classmethod
    def reset(cls):
        cls.class_index_cache = None
        cls.class_num_cache = None
        cls.labels_host = None
        cls.n_valid = 0

    @classmethod
    def sample_indices(cls, n_smp_cls, excl_indices=None):
        if cls.labels_host is None:
            raise RuntimeError('ClassBalancedRandomSampling.update_cache has not been called')
        excl_mask = None
        if excl_indices is not None and len(excl_indices) > 0:
            excl_mask = np.zeros(cls.labels_host.shape[0], dtype=bool)
            excl_mask[np.asarray(list(excl_indices), dtype=np.int64)] = True
        # only slots registered through update_cache take part (the reference samples from its cache)
        member = cls._member_mask()
        if excl_mask is None:
            excl_mask = ~member
        else:
            excl_mask |= ~member
        return class_balanced_indices(cls.labels_host, cls.labels_host.shape[0], n_smp_cls, excl_mask)

    @classmethod
    def _member_mask(cls):
        if cls._member is None or cls._member.shape[0] != cls.labels_host.shape[0]:
            cls._member = np.zeros(cls.labels_host.shape[0], dtype=bool)
        return cls._member

    _member = None

    @classmethod
    def sample(cls, buffer_x, buffer_y, n_smp_cls, excl_indices=None, device='cpu'):
        """Reference signature (buffer_utils.py:81): returns (x, y, sample_ind)."""
        ind = cls.sample_indices(n_smp_cls, excl_indices)
        ind_t = to_device_i64(ind, buffer_x.device)
        if buffer_x.is_cuda:
            return ops.gather_rows(buffer_x, ind_t), ops.gather_rows(buffer_y, ind_t), ind_t
        return buffer_x[ind_t], buffer_y[ind_t], ind_t

    @classmethod
    def update_cache(cls, buffer_y, num_class, new_y=None, ind=None, device='cpu', labels_host=None):
        """Incremental update (new_y/ind given, buffer_utils.py:140-154) or full rebuild from the
        label buffer (buffer_utils.py:155-160).  Accepts host arrays; device tensors are copied
        once (a sync) only when no host mirror is supplied."""
        def host(a):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                return a.detach().cpu().numpy().astype(np.int64)
            return np.asarray(a, dtype=np.int64)
        if cls.class_index_cache is None or new_y is None:
            n = buffer_y.shape[0]
            if cls.class_index_cache is None:
                cls.class_index_cache = {}
                cls.class_num_cache = np.zeros(num_class, dtype=np.int64)
                cls.labels_host = np.zeros(n, dtype=np.int64)
                cls._member = np.zeros(n, dtype=bool)
        if new_y is not None:
            new_y, ind = host(new_y), host(ind)
            for i, ny in zip(ind.tolist(), new_y.tolist()):
                if cls._member[i]:
                    oy = int(cls.labels_host[i])
                    cls.class_index_cache[oy].discard(i)
                    cls.class_num_cache[oy] -= 1
                cls.class_index_cache.setdefault(ny, set()).add(i)
                cls.class_num_cache[ny] += 1
                cls.labels_host[i] = ny
                cls._member[i] = True
        else:
            lab = host(labels_host) if labels_host is not None else host(buffer_y)
            cls.labels_host = lab.copy()
            cls._member = np.ones(lab.shape[0], dtype=bool)
            cache = {}
            for c in np.unique(lab).tolist():
                cache[c] = set(np.flatnonzero(lab == c).tolist())
            cls.class_index_cache = cache
            # the reference leaves class_num_cache untouched on this path (buffer_utils.py:155-160)


# --------------------------------------------------------------------------- buffer
def random_retrieve(buffer, num_retrieve, excl_indices=None, return_indices=False):
    """buffer_utils.py:9-26."""
    idx = uniform_indices(buffer.current_index, num_retrieve, excl_indices)
    buffer.last_random_idx = idx          # host record of the draw (tests replay it)
    x, y, idx_t = buffer.gather(idx)
    if return_indices:
        return x, y, idx_t
    return x, y


class Buffer(torch.nn.Module):
    """Same attributes as the reference Buffer (buffer.py:8-41): buffer_img, buffer_label
    (registered buffers), current_index, n_seen_so_far, model, params, device; plus labels_host,
    the numpy mirror every index decision reads."""

    def __init__(self, model, params, update_methods=None, retrieve_methods=None):
        super().__init__()
        self.params = params
        self.model = model
        self.cuda = self.params.cuda
        self.current_index = 0
        self.n_seen_so_far = 0
        use_cuda = torch.cuda.is_available()     # the reference ignores params.cuda here (buffer.py:22-23)
        self.device = 'cuda' if use_cuda else 'cpu'
        buffer_size = params.mem_size
        print('buffer has %d slots' % buffer_size)
        input_size = input_size_match[params.data]
        dev = torch.device(self.device)
        self.register_buffer('buffer_img', torch.zeros((buffer_size, *input_size), dtype=torch.float32, device=dev))
        self.register_buffer('buffer_label', torch.zeros(buffer_size, dtype=torch.int64, device=dev))
        self.labels_host = np.zeros(buffer_size, dtype=np.int64)
        if update_methods is None or retrieve_methods is None:
            from . import registry
            update_methods = update_methods or registry.update_methods
            retrieve_methods = retrieve_methods or registry.retrieve_methods
        self.update_method = update_methods[params.update](params)
        self.retrieve_method = retrieve_methods[params.retrieve](params)
        if getattr(self.params, 'buffer_tracker', False):
            raise NotImplementedError('buffer_tracker belongs to the match/mem_match retrievals, outside the replay path')

    def update(self, x, y, **kwargs):
        return self.update_method.update(buffer=self, x=x, y=y, **kwargs)

    def retrieve(self, **kwargs):
        return self.retrieve_method.retrieve(buffer=self, **kwargs)

    # ---- row movement
    def gather(self, idx_host):
        """(x [n,3,H,W], y [n], idx device tensor) for host slot indices."""
        idx_t = to_device_i64(idx_host, self.buffer_img.device)
        if self.buffer_img.is_cuda:
            return ops.gather_rows(self.buffer_img, idx_t), ops.gather_rows(self.buffer_label, idx_t), idx_t
        return self.buffer_img[idx_t], self.buffer_label[idx_t], idx_t

    def write(self, slots_host, x_rows, y_rows, y_host):
        """buffer_img[slots] = x_rows; buffer_label[slots] = y_rows (+ host mirror)."""
        slots_host = np.asarray(slots_host, dtype=np.int64)
        if slots_host.size == 0:
            return
        idx_t = to_device_i64(slots_host, self.buffer_img.device)
        if self.buffer_img.is_cuda:
            ops.scatter_rows(self.buffer_img, idx_t, x_rows)
            ops.scatter_rows(self.buffer_label, idx_t, y_rows)
        else:
            self.buffer_img[idx_t] = x_rows
            self.buffer_label[idx_t] = y_rows
        self.labels_host[slots_host] = np.asarray(y_host, dtype=np.int64)
