"""Error of the halo-strip tensor-core convolution against fp64 for per-tap promotion and for 3-tap TMEM chains
(B200OCL_TCP_CHAIN=3), next to the CUDA-core kernels.  GPU box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_conv import run_conv, reference  # noqa: E402


def main():
    for (N, H, W, cin, cout) in [(110, 32, 32, 20, 20), (110, 16, 16, 40, 40), (110, 8, 8, 80, 80), (110, 4, 4, 160, 160)]:
        g = torch.Generator(device='cuda').manual_seed(N + H)
        w = torch.randn(cout, cin, 3, 3, device='cuda', generator=g) / np.sqrt(9 * cin)
        x = torch.relu(torch.randn(N, H, W, cin, device='cuda', generator=g))
        ref = reference(x, w, 0)
        out = {}
        for name, path, chain in (('cuda-core', 1, '1'), ('tc per tap', 3, '1'), ('tc 3-tap chain', 3, '3')):
            os.environ['B200OCL_TCP_CHAIN'] = chain
            got = run_conv(x, w, 0, path)
            d = (got.double() - ref).abs()
            out[name] = (float(d.max() / ref.abs().max()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))
        print('%dx%d cin %d: ' % (H, W, cin) + '  '.join('%s max %.2e rms %.2e' % (k, v[0], v[1]) for k, v in out.items()), flush=True)
    os.environ['B200OCL_TCP_CHAIN'] = '1'


if __name__ == '__main__':
    main()
