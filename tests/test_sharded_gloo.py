"""CPU, world_size 2, gloo: orchestration of the memory-sharded kNN-SV path (shard bounds, the single
all-gather, deterministic combine) with the oracle injected as the per-shard kernel, and the
data-parallel gradient averaging hook of the learners."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import knn_sv as oknn


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_kernel(eval_f, eval_y, cand_f, cand_y, k):
    sv, _, _ = oknn.knn_sv_matrix(eval_f.numpy(), eval_y.numpy(), cand_f.numpy(), cand_y.numpy(), k, dtype=np.float32)
    return {'sum': torch.from_numpy(sv.sum(0)), 'max': torch.from_numpy(sv.max(0)), 'min': torch.from_numpy(sv.min(0))}


def _worker(rank, world, port, n_rows, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from b200ocl import sharded
    rs = np.random.RandomState(0)
    E, C, d, k = n_rows, 37, 12, 3
    ef = np.maximum(rs.standard_normal((E, d)), 0).astype(np.float32)
    cf = np.maximum(rs.standard_normal((C, d)), 0).astype(np.float32)
    ey, cy = rs.randint(0, 5, E), rs.randint(0, 5, C)
    lo, hi = sharded.shard_bounds(E, rank, world)
    red = sharded.knn_sv_sharded(torch.from_numpy(ef[lo:hi]), torch.from_numpy(ey[lo:hi]), torch.from_numpy(cf),
                                 torch.from_numpy(cy), k, kernel=_oracle_kernel)
    top, _ = sharded.aser_scores_sharded(torch.from_numpy(ef[lo:hi]), torch.from_numpy(ey[lo:hi]), E,
                                         torch.from_numpy(cf), torch.from_numpy(cy), k, 10, kernel=_oracle_kernel)
    # data-parallel gradient averaging as bench.py wires it
    g = torch.full((8,), float(rank + 1))
    dist.all_reduce(g)
    if rank == 0:
        sv, _, _ = oknn.knn_sv_matrix(ef, ey, cf, cy, k, dtype=np.float32)
        out.put({'sum_err': float(np.abs(red['sum'].numpy() - sv.sum(0)).max()),
                 'max_ok': bool(np.array_equal(red['max'].numpy(), sv.max(0))),
                 'min_ok': bool(np.array_equal(red['min'].numpy(), sv.min(0))),
                 'top_ok': bool(np.array_equal(top.numpy(), np.argsort(-sv.sum(0), kind='stable')[:10])),
                 'grad': g.tolist()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_rows', [50, 1, 7])
def test_sharded_knn_sv_world2(n_rows):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res['sum_err'] < 1e-5 and res['max_ok'] and res['min_ok'] and res['top_ok'], res
    assert res['grad'] == [3.0] * 8


def test_shard_bounds_cover():
    from b200ocl import sharded
    for n in [0, 1, 7, 50000]:
        for w in [1, 2, 4, 8]:
            spans = [sharded.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
