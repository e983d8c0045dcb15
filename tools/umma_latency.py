"""Time the single-CTA tcgen05 self-test kernel to separate MMA cost from staging cost:
mode 0 issues 4 dependent MMAs per 32-wide K block, mode 1 issues 12 (3xTF32)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200ocl import _native
from b200ocl.ops import _stream
lib = _native.lib()
for N in (32, 80, 256):
    for K in (1440, 2880):
        a = torch.rand(128, K, device='cuda'); b = torch.rand(N, K, device='cuda')
        d = torch.empty(128, N, device='cuda'); st = torch.zeros(1, dtype=torch.int32, device='cuda')
        for mode in (0, 1):
            for _ in range(3):
                lib.b200ocl_selftest_umma_tf32(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, mode, st.data_ptr(), _stream())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                lib.b200ocl_selftest_umma_tf32(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, mode, st.data_ptr(), _stream())
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            nblk = K // 32
            print('N=%3d K=%4d mode=%d  %.1f us/launch  %.0f ns per K block  (%d MMAs per block)' % (N, K, mode, us, 1e3 * us / nblk, 4 if mode == 0 else 12))
