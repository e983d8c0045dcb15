"""Kernel-level timings on the GPU box (CUDA events, warm-up, L2-sized inputs noted).
Writes gpurun_out/microbench.json.  Not the headline bench (that is bench.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import ops  # noqa: E402


def timeit(fn, warm=3, iters=10):
    try:
        return _timeit(fn, warm, iters)
    except Exception as e:  # keep the other rows
        return {'error': str(e)[:200]}


def _timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {'ms_median': ts[len(ts) // 2], 'ms_min': ts[0], 'ms_max': ts[-1]}


def time_graphed(fn, reps=20, iters=10):
    """Device time per call with the host out of the picture: `reps` calls captured into one CUDA graph,
    the replay timed with events (median of `iters`), divided by reps."""
    try:
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps)
        ts.sort()
        return {'ms_median': ts[len(ts) // 2], 'ms_min': ts[0], 'ms_max': ts[-1], 'how': 'cuda graph of %d calls' % reps}
    except Exception as e:
        return {'error': str(e)[:200]}


def main():
    res = {}
    g = torch.Generator(device='cuda').manual_seed(0)
    for (E, C, d) in [(10, 100, 160), (100, 100, 160), (110, 160, 160), (5000, 200, 160), (50000, 1000, 512),
                      (200, 2000, 160), (1000, 50000, 512)]:      # the last two: scratch-line kernel (C > 1024)
        ef = torch.relu(torch.randn(E, d, device='cuda', generator=g))
        cf = torch.relu(torch.randn(C, d, device='cuda', generator=g))
        ey = torch.randint(0, 100, (E,), device='cuda', generator=g)
        cy = torch.randint(0, 100, (C,), device='cuda', generator=g)
        r = timeit(lambda: ops.knn_sv(ef, ey, cf, cy, 3, want_sum=True))
        if 'ms_median' in r:
            r['alg_bytes'] = 4 * d * (E + C) + 8 * (E + C) + 4 * C
            r['GBps'] = r['alg_bytes'] / (r['ms_median'] * 1e-3) / 1e9
            r['direct_form_TFLOPs'] = 3.0 * E * C * d / (r['ms_median'] * 1e-3) / 1e12
        res['knn_sv_%dx%dx%d' % (E, C, d)] = r
    for B in [110, 1024, 4096, 8192]:
        f = torch.nn.functional.normalize(torch.randn(B, 2, 128, device='cuda', generator=g), dim=2)
        y = torch.randint(0, 100, (B,), device='cuda', generator=g)
        r = time_graphed(lambda: ops.supcon(f, y, 0.07), reps=20 if B <= 1024 else 3)
        if 'ms_median' in r:
            A = 2 * B
            r['alg_bytes'] = 2 * 4 * A * 128 + 8 * B + 4
            r['GBps'] = r['alg_bytes'] / (r['ms_median'] * 1e-3) / 1e9
            r['TFLOPs_6A2d'] = 6.0 * A * A * 128 / (r['ms_median'] * 1e-3) / 1e12
            r['per_call_with_python_ms'] = timeit(lambda: ops.supcon(f, y, 0.07)).get('ms_median')
        res['supcon_B%d' % B] = r
    v = torch.randn(160, device='cuda')
    res['rank_desc_160'] = timeit(lambda: ops.rank_desc(v, 10))
    src = torch.randn(5000, 3, 32, 32, device='cuda')
    idx = torch.randperm(5000, device='cuda')[:110]
    res['gather_110x12KB'] = timeit(lambda: ops.gather_rows(src, idx))
    p, gr = torch.randn(1109240, device='cuda'), torch.randn(1109240, device='cuda')
    res['sgd_1.1M'] = timeit(lambda: ops.sgd_step(p, gr, 0.1))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'microbench.json'), 'w') as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
