"""Retrieval plugins with the reference's constructor / retrieve(buffer, **kwargs) surface:
Random_retrieve (utils/buffer/random_retrieve.py:3-9), MIR_retrieve
(utils/buffer/mir_retrieve.py:8-64) and ASER_retrieve (utils/buffer/aser_retrieve.py:8-92).

Index decisions are host-side numpy on the label mirror; arithmetic is CUDA through the C ABI:
one eval-mode feature pass for current batch + cooperative samples + candidates (the reference
runs the candidates through the network twice, aser_retrieve.py:64,76), two fused kNN-SV
launches, one ranking launch, row gathers.
"""
import numpy as np
import torch

from . import ops
from .engine import ce_loss
from .memory import ClassBalancedRandomSampling, n_classes, random_retrieve, to_device_i64
from .nets import engine_of


class Random_retrieve(object):
    def __init__(self, params):
        super().__init__()
        self.num_retrieve = params.eps_mem_batch

    def retrieve(self, buffer, **kwargs):
        return random_retrieve(buffer, self.num_retrieve)


class MIR_retrieve(object):
    """Maximally interfered retrieval: rank a random subsample by the loss increase a virtual SGD
    step on the current batch gradient would cause (mir_retrieve.py:15-30).  The virtual weights
    theta - lr*grad are written to a second arena by one fused kernel (no copy.deepcopy, no
    per-parameter loops); both forwards run in train mode as in the reference, the live model's
    running statistics move once, the virtual copy's are discarded."""

    def __init__(self, params, **kwargs):
        super().__init__()
        self.params = params
        self.subsample = params.subsample
        self.num_retrieve = params.eps_mem_batch
        if self.subsample > ops.RANK_MAX:
            raise ValueError('MIR subsample %d exceeds the ranking kernel limit %d (b200ocl_rank_desc)'
                             % (self.subsample, ops.RANK_MAX))

    def retrieve(self, buffer, **kwargs):
        sub_x, sub_y = random_retrieve(buffer, self.subsample)
        if sub_x.size(0) == 0:
            return sub_x, sub_y
        eng = engine_of(buffer.model)
        virt = eng.virtual_state()
        virt.bn_stats.copy_(eng.state.bn_stats)                       # deepcopy'd buffers (mir_retrieve.py:41)
        eng.sgd_step(self.params.learning_rate, 0.0, dst=virt)        # theta' = theta - lr * grad (:43-46)
        logits_pre, _ = eng.forward_train(sub_x, slot=2)
        pre = ce_loss(logits_pre, sub_y, want_grad=False, want_per_sample=True)['per_sample']
        logits_post, _ = eng.forward_train(sub_x, state=virt, slot=2)
        post = ce_loss(logits_post, sub_y, want_grad=False, want_per_sample=True)['per_sample']
        big_ind = ops.rank_desc(post, self.num_retrieve, sa=1.0, b=pre, sb=-1.0)   # scores = post - pre
        self.last_scores = (pre, post)
        self.last_top = big_ind
        return ops.gather_rows(sub_x, big_ind), ops.gather_rows(sub_y, big_ind)


def aser_retrieve_select(eng_features, cur_f, cur_y, coop_f, coop_y, cand_f, cand_y, k, aser_type, num_retrieve):
    """Candidate positions retrieved by ASER given deep features (aser_retrieve.py:60-91).
    Returns a device int64 tensor of positions into the candidate set, in rank order."""
    adv = ops.knn_sv(cur_f, cur_y, cand_f, cand_y, k, want_sum=(aser_type != 'asv'), want_min=(aser_type == 'asv'))
    if aser_type == 'neg_sv':
        return ops.rank_desc(adv['sum'], num_retrieve, sa=-1.0)
    coop = ops.knn_sv(coop_f, coop_y, cand_f, cand_y, k, want_sum=(aser_type != 'asv'), want_max=(aser_type == 'asv'))
    if aser_type == 'asv':
        return ops.rank_desc(coop['max'], num_retrieve, sa=1.0, b=adv['min'], sb=-1.0)
    # asvm (and anything else): mean over eval rows = column sum / row count
    return ops.rank_desc(coop['sum'], num_retrieve, sa=1.0 / coop_f.shape[0], b=adv['sum'], sb=-1.0 / cur_f.shape[0])


class ASER_retrieve(object):
    def __init__(self, params, **kwargs):
        super().__init__()
        self.num_retrieve = params.eps_mem_batch
        self.device = 'cuda' if torch.cuda.is_available() else 'cpu'
        self.k = params.k
        self.mem_size = params.mem_size
        self.aser_type = params.aser_type
        self.n_smp_cls = int(params.n_smp_cls)
        self.out_dim = n_classes[params.data]
        self.is_aser_upt = params.update == 'ASER'
        if self.n_smp_cls * self.out_dim > min(ops.KNN_MAX_CAND, ops.RANK_MAX):
            raise ValueError('ASER retrieve: n_smp_cls*num_classes = %d candidates exceed the kNN-SV / ranking kernel limit %d'
                             % (self.n_smp_cls * self.out_dim, min(ops.KNN_MAX_CAND, ops.RANK_MAX)))
        ClassBalancedRandomSampling.reset()

    def retrieve(self, buffer, **kwargs):
        if buffer.n_seen_so_far <= self.mem_size:
            # random retrieval until the buffer has been filled once (aser_retrieve.py:24-26)
            return random_retrieve(buffer, self.num_retrieve)
        return self._retrieve_by_knn_sv(buffer, kwargs['x'], kwargs['y'], self.num_retrieve)

    def _retrieve_by_knn_sv(self, buffer, cur_x, cur_y, num_retrieve):
        eng = engine_of(buffer.model)
        CB = ClassBalancedRandomSampling
        if not self.is_aser_upt:
            CB.update_cache(buffer.buffer_label, self.out_dim, labels_host=buffer.labels_host)
        cand_ind = CB.sample_indices(self.n_smp_cls)
        coop_ind = np.zeros(0, dtype=np.int64)
        if self.aser_type != 'neg_sv':
            coop_ind = CB.sample_indices(self.n_smp_cls, excl_indices=cand_ind)
        n_cur, n_coop, n_cand = cur_x.shape[0], coop_ind.size, cand_ind.size
        self.last_choices = {'ret_cand_ind': cand_ind, 'ret_coop_ind': coop_ind}
        # one batch [cur | coop | cand], one eval-mode feature pass
        batch = torch.empty((n_cur + n_coop + n_cand,) + tuple(cur_x.shape[1:]), dtype=torch.float32,
                            device=cur_x.device)
        batch[:n_cur].copy_(cur_x)
        idx_t = to_device_i64(np.concatenate([coop_ind, cand_ind]), cur_x.device)
        ops.gather_rows(buffer.buffer_img, idx_t, out=batch[n_cur:])
        feats = eng.features_eval(batch)
        coop_y = to_device_i64(buffer.labels_host[coop_ind], cur_x.device)
        cand_y = to_device_i64(buffer.labels_host[cand_ind], cur_x.device)
        pos = aser_retrieve_select(eng, feats[:n_cur], cur_y, feats[n_cur:n_cur + n_coop], coop_y,
                                   feats[n_cur + n_coop:], cand_y, self.k, self.aser_type,
                                   min(num_retrieve, n_cand))
        cand_x = batch[n_cur + n_coop:]
        self.last_pos = pos
        return ops.gather_rows(cand_x, pos), ops.gather_rows(cand_y, pos)
