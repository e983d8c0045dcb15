"""Which descriptor form reads a row-shifted window of a swizzled patch correctly?  (GPU box)
Prints, per (start_row, sbo_rows, base-offset mode), whether the tcgen05 result equals numpy exactly."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import _native  # noqa: E402
from b200ocl.ops import _stream  # noqa: E402


def main():
    lib = _native.lib()
    rs = np.random.RandomState(0)
    rows, N = 320, 32
    P = (rs.randint(-8, 9, (rows, 32)) / 8.0).astype(np.float32)
    B = (rs.randint(-8, 9, (N, 32)) / 8.0).astype(np.float32)
    p, b = torch.tensor(P).cuda(), torch.tensor(B).cuda()
    for sbo in (8, 16, 10):
        for start in (0, 1, 2, 7, 16, 17, 18, 34):
            if start + 15 * sbo + 8 > rows:
                continue
            idx = np.array([start + (i // 8) * sbo + (i % 8) for i in range(128)])
            ref = (P[idx].astype(np.float64) @ B.astype(np.float64).T).astype(np.float32)
            out = []
            for mode in (0, 1):
                d = torch.full((128, N), float('nan'), device='cuda')
                st = torch.full((1,), -1, dtype=torch.int32, device='cuda')
                rc = lib.b200ocl_selftest_umma_window(p.data_ptr(), b.data_ptr(), d.data_ptr(), rows, start, sbo, mode, N,
                                                      st.data_ptr(), _stream())
                _native.check(rc, 'window')
                torch.cuda.synchronize()
                D = d.cpu().numpy()
                bad_rows = int((np.abs(D - ref).max(axis=1) > 0).sum())
                out.append('mode%d: status %d, %3d/128 rows wrong' % (mode, int(st), bad_rows))
            print('sbo_rows %2d start_row %2d | %s' % (sbo, start, ' | '.join(out)), flush=True)


if __name__ == '__main__':
    main()
