"""Run the fused SupCon kernel a few times (for ncu captures): python tools/supcon_prof.py [B ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import ops  # noqa: E402

for B in [int(v) for v in sys.argv[1:]] or [110, 4096]:
    g = torch.Generator(device='cuda').manual_seed(0)
    f = torch.nn.functional.normalize(torch.randn(B, 2, 128, device='cuda', generator=g), dim=2)
    y = torch.randint(0, 100, (B,), device='cuda', generator=g)
    for _ in range(3):
        loss, grad = ops.supcon(f, y, 0.07)
    torch.cuda.synchronize()
    print(B, float(loss))
