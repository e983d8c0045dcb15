"""Memory-sharded kNN Shapley values (BASELINE config 5).

Eval rows (the replay memory's features) are independent units: a row's sort and recurrence need
every candidate but no other row, and the ASER plugins consume only reductions over rows
(aser_retrieve.py:79-86, aser_update.py:80).  So each rank runs the fused kernel on its shard of
eval rows against the replicated candidate block and contributes a [3, C] block of column partials
(sum, max, min); ONE all-gather moves them (12 KB per rank at C = 1000), every rank combines them
in rank order (deterministic) and ranks the candidates.  No other collective is on the path.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops


def shard_bounds(n_rows, rank, world):
    """Contiguous, balanced row range of `rank`."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _local_partials(eval_f, eval_y, cand_f, cand_y, k, kernel):
    C = cand_f.shape[0]
    part = torch.empty((3, C), dtype=torch.float32, device=cand_f.device)
    if eval_f.shape[0] == 0:
        part[0].zero_()
        part[1].fill_(-torch.finfo(torch.float32).max)
        part[2].fill_(torch.finfo(torch.float32).max)
        return part
    out = kernel(eval_f, eval_y, cand_f, cand_y, k)
    part[0].copy_(out['sum']); part[1].copy_(out['max']); part[2].copy_(out['min'])
    return part


def _cuda_kernel(eval_f, eval_y, cand_f, cand_y, k):
    return ops.knn_sv(eval_f, eval_y, cand_f, cand_y, k, want_sum=True, want_max=True, want_min=True)


def knn_sv_sharded(eval_f_local, eval_y_local, cand_f, cand_y, k, group=None, kernel=None):
    """Column reductions of the global SV matrix given this rank's shard of eval rows.
    Returns {'sum','max','min'} [C] tensors, identical on every rank."""
    kernel = kernel or _cuda_kernel
    part = _local_partials(eval_f_local, eval_y_local, cand_f, cand_y, k, kernel)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {'sum': part[0], 'max': part[1], 'min': part[2]}
    world = dist.get_world_size(group)
    flat = torch.empty((world * part.shape[0], part.shape[1]), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(flat, part.contiguous(), group=group)         # the one collective on the path
    gathered = flat.view(world, part.shape[0], part.shape[1])
    total = gathered[:, 0].to(torch.float64).cumsum(0)[-1].to(torch.float32)  # rank order, fp64 accumulate
    return {'sum': total, 'max': gathered[:, 1].max(0).values, 'min': gathered[:, 2].min(0).values}


def aser_scores_sharded(eval_f_local, eval_y_local, n_eval_total, cand_f, cand_y, k, n_top, group=None, kernel=None):
    """-(mean SV) ranking of the candidates against a sharded evaluation memory: indices of the n_top
    candidates with the largest summed SV (aser_update.py:80-93 semantics over a sharded memory)."""
    red = knn_sv_sharded(eval_f_local, eval_y_local, cand_f, cand_y, k, group=group, kernel=kernel)
    if red['sum'].is_cuda:
        return ops.rank_desc(red['sum'], n_top), red
    order = np.argsort(-red['sum'].numpy(), kind='stable')[:n_top]
    return torch.from_numpy(order), red
