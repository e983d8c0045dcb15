"""Layout type 1 (SWIZZLE_128B_BASE32B), rows stored plain: where does the hardware look?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import _native
from b200ocl.ops import _stream
lib = _native.lib()
ROWS = 200

def run(PA, PB, a_lbo, a_sbo, b_lbo, b_sbo, flags, N=32):
    a, b = torch.tensor(PA).cuda().contiguous(), torch.tensor(PB).cuda().contiguous()
    d = torch.full((128, N), float('nan'), device='cuda')
    status = torch.full((1,), -1, dtype=torch.int32, device='cuda')
    rc = lib.b200ocl_selftest_umma_mn(a.data_ptr(), b.data_ptr(), d.data_ptr(), ROWS, ROWS, 0, a_lbo, a_sbo, 0, b_lbo,
                                      b_sbo, 1 | flags, N, status.data_ptr(), _stream())
    _native.check(rc, 'umma_mn'); torch.cuda.synchronize()
    return d.cpu().numpy(), int(status)

ones = np.ones((ROWS, 32), dtype=np.float32)
rs = np.random.RandomState(0)
R = (rs.randint(-8, 9, (ROWS, 32)) / 8.0).astype(np.float32)
for flags, name in ((3 << 8 | 1 << 10 | 1 << 11, 'type1 plain both-MN'), (3 << 8 | 1 << 10, 'type1 sw128-stored both-MN'),
                    (1 << 8 | 1 << 10 | 1 << 11, 'type1 plain A-MN only'), (3 << 8 | 1 << 11, 'type2 plain both-MN')):
    D, st = run(R, R, 8, 8, 8, 8, flags)
    print(name, 'status', st, 'nan', int(np.isnan(D).sum()), 'zeros', int((D == 0).sum()), 'D[0,:4]', D[0, :4])
flags = 3 << 8 | 1 << 10 | 1 << 11
for (lbo, sbo) in ((8, 4), (8, 8), (4, 8)):
    print('== A one-hot map, type1 plain, lbo_rows=%d sbo_rows=%d: (r, c) -> D row : value' % (lbo, sbo))
    for r in range(0, 20):
        line = []
        for c in range(0, 32):
            P = np.zeros((ROWS, 32), dtype=np.float32); P[r, c] = 1
            D, st = run(P, ones, lbo, sbo, 8, 4, flags)
            rows = np.flatnonzero(np.nan_to_num(np.abs(D)).sum(1) > 0)
            line.append('.' if rows.size == 0 else ('%d' % rows[0] if rows.size == 1 else '%d+' % rows[0]))
        print('r=%2d ' % r + ' '.join('%3s' % t for t in line))
