"""Memory-sharded kNN-SV sweep (BASELINE config 5) on N GPUs: one process per GPU (torchrun), NCCL.
Each rank scores its shard of the 50k x 512 buffer features against the replicated 1k candidates with
the fused kernel, ONE all-gather moves the [3, C] column partials, every rank combines them in rank
order.  Checks the result against the unsharded kernel on rank 0 and prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_sweep.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import ops, sharded  # noqa: E402


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    E = int(os.environ.get('SWEEP_ROWS', 50000))
    C, d, k = int(os.environ.get('SWEEP_CAND', 1000)), int(os.environ.get('SWEEP_DIM', 512)), 3
    g = torch.Generator(device='cuda').manual_seed(0)          # same data on every rank
    ef = torch.relu(torch.randn(E, d, device='cuda', generator=g))
    cf = torch.relu(torch.randn(C, d, device='cuda', generator=g))
    ey = torch.randint(0, 100, (E,), device='cuda', generator=g)
    cy = torch.randint(0, 100, (C,), device='cuda', generator=g)
    lo, hi = sharded.shard_bounds(E, rank, world)
    ef_l, ey_l = ef[lo:hi].contiguous(), ey[lo:hi].contiguous()

    def run():
        return sharded.aser_scores_sharded(ef_l, ey_l, E, cf, cy, k, 100)
    for _ in range(3):
        run()
    times = []
    for _ in range(10):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        top, red = run()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    times.sort()
    ms = times[len(times) // 2]
    ok = True
    err = 0.0
    if rank == 0:
        full = ops.knn_sv(ef, ey, cf, cy, k, want_sum=True, want_max=True, want_min=True)
        err = float((full['sum'] - red['sum']).abs().max())
        ok = err < 2e-4 and torch.equal(full['max'], red['max']) and torch.equal(full['min'], red['min'])
        # ranking may differ only where the summed SVs are within the fp32 combine noise
        top_full = ops.rank_desc(full['sum'], 100)
        same = float((top_full == top).float().mean())
        print(json.dumps({'workload': 'kNN-SV sweep 50000x512 eval, 1000 cand, k=3, eval rows sharded', 'n_gpus': world,
                          'ms': ms, 'rows_per_s': E / (ms * 1e-3), 'max_abs_err_vs_unsharded': err,
                          'top100_identical_fraction': same, 'ok': bool(ok)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
