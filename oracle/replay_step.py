"""Oracle: one full replay step on the CPU, as the reference executes it.

Restates the per-minibatch loop bodies of agents/exp_replay.py:34-92 (ER with random / MIR /
ASER retrieval and reservoir / ASER update) and agents/scr.py:40-63 (SCR) on top of the other
oracle modules, INCLUDING the work the reference does and throws away (the two discarded
backward passes of the ASER branch, exp_replay.py:55,77,81; the candidate features computed
twice, aser_retrieve.py:64,76; deep features in chunks of 64, utils/utils.py:68-83), because this
module is also the CPU baseline that bench.py times (`--impl reference`, `cpu_baseline`).
Test infrastructure only -- see oracle/__init__.py.

Random choices (which buffer slots are sampled) can be injected through `choices` so that a
GPU trajectory can be replayed decision by decision ("index-exactness").
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import aser as oaser
from . import knn_sv as oknn
from . import resnet as oresnet
from . import supcon as osup


class ReplayState:
    """Model + optimizer + buffer state of one learner."""

    def __init__(self, spec, params, bn, mem_size, in_shape, num_classes, lr=0.1, weight_decay=0.0):
        self.spec = spec
        self.params = params
        self.bn = bn
        self.lr = lr
        self.wd = weight_decay
        self.mem_size = mem_size
        self.num_classes = num_classes
        self.buffer_img = torch.zeros((mem_size,) + tuple(in_shape), dtype=torch.float32)
        self.buffer_label = torch.zeros(mem_size, dtype=torch.long)
        self.current_index = 0
        self.n_seen_so_far = 0
        self.class_num_cache = np.zeros(num_classes, dtype=np.int64)
        self.class_index_cache = {}
        self.log = {}

    # ---- class cache (buffer_utils.py:140-154)
    def cache_update(self, ind, new_y):
        for i, ny in zip(np.asarray(ind).tolist(), np.asarray(new_y).tolist()):
            for c, s in self.class_index_cache.items():
                if i in s:
                    s.discard(i)
                    self.class_num_cache[c] -= 1
                    break
            self.class_index_cache.setdefault(ny, set()).add(i)
            self.class_num_cache[ny] += 1


def _features_eval_chunked(st, x):
    """mini_batch_deep_features: eval mode, no_grad, chunks of 64 (utils/utils.py:45-90)."""
    outs = []
    with torch.no_grad():
        for s in range(0, x.shape[0], 64):
            outs.append(oresnet.features(st.spec, st.params, st.bn, x[s:s + 64], train=False))
    return torch.cat(outs) if outs else torch.zeros((0, st.spec.dim_in))


def _knn_sv_images(st, eval_x, eval_y, cand_x, cand_y, k):
    """compute_knn_sv on images (aser_utils.py:7-61): features of eval+cand, then the SV matrix."""
    feats = _features_eval_chunked(st, torch.cat((eval_x, cand_x)))
    ef, cf = feats[:eval_x.shape[0]].numpy(), feats[eval_x.shape[0]:].numpy()
    sv, _, _ = oknn.knn_sv_matrix(ef, eval_y.numpy(), cf, cand_y.numpy(), k, dtype=np.float32)
    return sv


def _random_indices(st, n, excl=None):
    valid = np.setdiff1d(np.arange(st.current_index), np.array([] if excl is None else list(excl)))
    n = min(n, valid.shape[0])
    return np.random.choice(valid, n, replace=False).astype(np.int64)


def _cbrs_indices(st, n_smp_cls, excl=None):
    """ClassBalancedRandomSampling.sample (buffer_utils.py:81-121): per-class python loop."""
    excl = set() if excl is None else set(np.asarray(excl).tolist())
    out = []
    for ind_set in st.class_index_cache.values():
        if ind_set:
            valid = list(ind_set - excl)
            perm = torch.randperm(len(valid))
            out += [valid[i] for i in perm[:n_smp_cls].tolist()]
    return np.asarray(out, dtype=np.int64)


def _train_fwd_bwd(st, x, y, keep_grad=True):
    """model.forward in train mode + mean CE + backward (exp_replay.py:40-55)."""
    loss, logits, grads = oresnet.ce_loss_and_grads(st.spec, st.params, st.bn, x, y)
    return loss, logits, grads


def _reservoir_update(st, x, y, choices):
    place_left = max(0, st.mem_size - st.current_index)
    written = []
    if place_left:
        off = min(place_left, x.shape[0])
        st.buffer_img[st.current_index:st.current_index + off] = x[:off]
        st.buffer_label[st.current_index:st.current_index + off] = y[:off]
        written = list(range(st.current_index, st.current_index + off))
        st.current_index += off
        st.n_seen_so_far += off
        if off == x.shape[0]:
            return written
    x, y = x[place_left:], y[place_left:]
    draws = choices.get('reservoir_draws') if choices else None
    if draws is None:
        draws = torch.FloatTensor(x.shape[0]).uniform_(0, st.n_seen_so_far).long().numpy()
    st.log['reservoir_draws'] = np.asarray(draws)
    st.n_seen_so_far += x.shape[0]
    slots, src = oaser.reservoir_slots(draws, st.mem_size)
    if slots:
        st.buffer_img[slots] = x[src]
        st.buffer_label[slots] = y[src]
    return slots


def _aser_retrieve(st, cur_x, cur_y, num_retrieve, k, aser_type, n_smp_cls, choices):
    cand_ind = choices['ret_cand_ind'] if choices and 'ret_cand_ind' in choices else _cbrs_indices(st, n_smp_cls)
    cand_x, cand_y = st.buffer_img[cand_ind], st.buffer_label[cand_ind]
    sv_adv = _knn_sv_images(st, cur_x, cur_y, cand_x, cand_y, k)                       # aser_retrieve.py:64
    sv_coop = None
    coop_ind = np.zeros(0, dtype=np.int64)
    if aser_type != 'neg_sv':
        coop_ind = choices['ret_coop_ind'] if choices and 'ret_coop_ind' in choices else \
            _cbrs_indices(st, n_smp_cls, excl=cand_ind)
        sv_coop = _knn_sv_images(st, st.buffer_img[coop_ind], st.buffer_label[coop_ind], cand_x, cand_y, k)  # :75-76
    pos = oaser.retrieve_indices(sv_adv, sv_coop, aser_type, num_retrieve)
    st.log.update(ret_cand_ind=np.asarray(cand_ind), ret_coop_ind=np.asarray(coop_ind), ret_pos=pos,
                  ret_score=oaser.retrieve_score(sv_adv, sv_coop, aser_type))
    return cand_x[pos], cand_y[pos]


def _aser_update(st, x, y, k, n_smp_cls, choices):
    place_left = st.mem_size - st.current_index
    if place_left:
        n_fit = min(place_left, x.shape[0])
        st.cache_update(np.arange(st.current_index, st.current_index + n_fit), y[:n_fit].numpy())
        _reservoir_update(st, x[:n_fit], y[:n_fit], None)
    if st.current_index != st.mem_size:
        return
    cur_x, cur_y = x[place_left:], y[place_left:]
    if choices and 'upd_threshold' in choices:
        threshold = choices['upd_threshold']
    else:
        threshold = torch.tensor(1).float().uniform_(0, 1 / st.num_classes).item()
    minority = oaser.minority_positions(cur_y.numpy(), st.class_num_cache, st.mem_size, threshold)
    eval_ind = choices['upd_eval_ind'] if choices and 'upd_eval_ind' in choices else _cbrs_indices(st, n_smp_cls)
    n_total = int(choices['n_total_smp']) if choices and 'n_total_smp' in choices else None
    cand_ind = choices['upd_cand_ind'] if choices and 'upd_cand_ind' in choices else \
        _random_indices(st, n_total if n_total is not None else int(1.5 * st.num_classes), excl=eval_ind)
    eval_x = torch.cat((st.buffer_img[eval_ind], cur_x[minority]))
    eval_y = torch.cat((st.buffer_label[eval_ind], cur_y[minority]))
    cand_x = torch.cat((st.buffer_img[cand_ind], cur_x))
    cand_y = torch.cat((st.buffer_label[cand_ind], cur_y))
    sv = _knn_sv_images(st, eval_x, eval_y, cand_x, cand_y, k)
    sv_sum = sv.sum(0)
    ind_cur, ind_buffer = oaser.update_partition(sv_sum, len(cand_ind), cand_ind)
    st.n_seen_so_far += cur_x.shape[0]
    st.log.update(upd_eval_ind=np.asarray(eval_ind), upd_cand_ind=np.asarray(cand_ind), upd_threshold=threshold,
                  upd_ind_cur=ind_cur, upd_ind_buffer=ind_buffer, upd_sv_sum=sv_sum)
    if len(ind_cur):
        st.cache_update(ind_buffer, cur_y[ind_cur].numpy())
        st.buffer_img[ind_buffer] = cur_x[ind_cur]
        st.buffer_label[ind_buffer] = cur_y[ind_cur]


def er_step(st, batch_x, batch_y, retrieve='random', update='random', eps_mem_batch=10, k=3, aser_type='asvm',
            n_smp_cls=1, subsample=50, choices=None):
    """One iteration of agents/exp_replay.py:34-92 (mem_iters = 1).  Returns the loss that was
    optimised (combined loss in the ASER branch, stream-batch loss otherwise)."""
    aser_branch = (update == 'ASER' or retrieve == 'ASER')
    loss, logits, grads = _train_fwd_bwd(st, batch_x, batch_y)                         # :40-55
    # ---- retrieve (:58)
    if retrieve == 'random' or (retrieve == 'ASER' and st.n_seen_so_far <= st.mem_size):
        idx = choices['ret_idx'] if choices and 'ret_idx' in choices else _random_indices(st, eps_mem_batch)
        st.log['ret_idx'] = np.asarray(idx)
        mem_x, mem_y = st.buffer_img[idx], st.buffer_label[idx]
    elif retrieve == 'MIR':
        idx = choices['mir_idx'] if choices and 'mir_idx' in choices else _random_indices(st, subsample)
        sub_x, sub_y = st.buffer_img[idx], st.buffer_label[idx]
        if sub_x.shape[0] > 0:
            scores = oresnet.mir_scores(st.spec, st.params, st.bn, grads, st.lr, sub_x, sub_y)
            top = np.argsort(-scores.numpy(), kind='stable')[:eps_mem_batch]
            st.log.update(mir_idx=np.asarray(idx), mir_scores=scores.numpy(), mir_top=top)
            mem_x, mem_y = sub_x[top], sub_y[top]
        else:
            mem_x, mem_y = sub_x, sub_y
    else:
        mem_x, mem_y = _aser_retrieve(st, batch_x, batch_y, eps_mem_batch, k, aser_type, n_smp_cls, choices)
    # ---- replay forward/backward (:59-77): gradients accumulate onto the stream-batch gradients
    if mem_x.shape[0] > 0:
        loss_m, _, grads_m = _train_fwd_bwd(st, mem_x, mem_y)
        for key in grads:
            if grads[key] is not None and grads_m[key] is not None:
                grads[key] = grads[key] + grads_m[key]
    if aser_branch:
        # zero_grad + combined forward/backward (:79-87): the two gradients above are discarded
        combined = torch.cat((mem_x, batch_x))
        labels = torch.cat((mem_y, batch_y))
        loss, _, grads = _train_fwd_bwd(st, combined, labels)
    oresnet.sgd_step(st.params, grads, st.lr, st.wd)                                   # :87 / :89
    # ---- update (:92)
    if update == 'ASER':
        _aser_update(st, batch_x, batch_y, k, n_smp_cls, choices)
    else:
        st.log['written'] = _reservoir_update(st, batch_x, batch_y, choices)
    return float(loss)


def scr_step(st, batch_x, batch_y, eps_mem_batch=100, temperature=0.07, transform=None, choices=None):
    """One iteration of agents/scr.py:40-63 with random retrieval / reservoir update.
    `transform` maps the combined batch to its augmented view (identity when None: kornia is
    absent and unpinned)."""
    loss_val = None
    idx = choices['ret_idx'] if choices and 'ret_idx' in choices else _random_indices(st, eps_mem_batch)
    st.log['ret_idx'] = np.asarray(idx)
    mem_x, mem_y = st.buffer_img[idx], st.buffer_label[idx]
    if mem_x.shape[0] > 0:
        combined = torch.cat((mem_x, batch_x))
        labels = torch.cat((mem_y, batch_y))
        aug = combined if transform is None else transform(combined)
        leaves = OrderedDict((k_, v.detach().clone().requires_grad_(True)) for k_, v in st.params.items())
        f1 = oresnet.forward(st.spec, leaves, st.bn, combined, train=True)
        f2 = oresnet.forward(st.spec, leaves, st.bn, aug, train=True)
        feats = torch.stack((f1, f2), dim=1)
        loss, dfeat = osup.supcon_loss_and_grad(feats.detach().numpy(), labels.numpy(), temperature, dtype=np.float32)
        feats.backward(torch.tensor(dfeat, dtype=torch.float32))
        grads = OrderedDict((k_, v.grad) for k_, v in leaves.items())
        oresnet.sgd_step(st.params, grads, st.lr, st.wd)
        loss_val = float(loss)
    st.log['written'] = _reservoir_update(st, batch_x, batch_y, choices)
    return loss_val
