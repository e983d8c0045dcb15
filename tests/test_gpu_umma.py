"""GPU: the tcgen05 / TMEM path (descriptor encodings of csrc/umma.cuh) against numpy."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_selftest(A, B, mode):
    from b200ocl import _native
    from b200ocl.ops import _stream
    lib = _native.lib()
    a, b = torch.tensor(A).cuda().contiguous(), torch.tensor(B).cuda().contiguous()
    N, K = B.shape
    d = torch.full((128, N), float('nan'), device='cuda')
    status = torch.full((1,), -1, dtype=torch.int32, device='cuda')
    rc = lib.b200ocl_selftest_umma_tf32(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, mode, status.data_ptr(), _stream())
    _native.check(rc, 'b200ocl_selftest_umma_tf32')
    torch.cuda.synchronize()
    assert int(status) == 0, 'MMA completion barrier timed out'
    return d.cpu().numpy()


@pytest.mark.parametrize('N,K', [(16, 32), (32, 64), (80, 192), (160, 96), (256, 32), (48, 736)])
def test_umma_tf32_exact_on_representable_inputs(N, K):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    rs = np.random.RandomState(N + K)
    A = (rs.randint(-8, 9, (128, K)) / 8.0).astype(np.float32)       # exactly representable in TF32
    B = (rs.randint(-8, 9, (N, K)) / 8.0).astype(np.float32)
    D = run_selftest(A, B, 0)
    np.testing.assert_array_equal(D, (A.astype(np.float64) @ B.astype(np.float64).T).astype(np.float32))


@pytest.mark.parametrize('N,K', [(32, 192), (80, 736), (160, 1440)])
def test_umma_3xtf32_is_fp32_grade(N, K):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    rs = np.random.RandomState(N)
    A = np.maximum(rs.standard_normal((128, K)), 0).astype(np.float32)
    B = (rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    D1 = run_selftest(A, B, 0)
    D3 = run_selftest(A, B, 1)
    scale = np.abs(ref).max()
    e1 = np.abs(D1 - ref).max() / scale
    e3 = np.abs(D3 - ref).max() / scale
    fp32 = np.abs((A @ B.T) - ref).max() / scale
    print('K=%d  single-pass TF32 %.2e   3xTF32 %.2e   numpy fp32 %.2e' % (K, e1, e3, fp32))
    assert e1 > 1e-5          # a single TF32 pass is visibly lossy
    # One long TMEM accumulation: the tensor core's accumulate truncates, the error grows ~6.4e-9 * K
    # (measured 1.6e-6 / 5.3e-6 / 9.2e-6 at K = 192 / 736 / 1440).  conv_tc.cu therefore promotes every
    # 32-wide K block into fp32 registers; its parity is covered by tests/test_gpu_net.py.
    assert e3 < 1.2e-8 * K + 1e-6, (e3, fp32)


def run_mn(PA, PB, a_row0, a_lbo, a_sbo, b_row0, b_lbo, b_sbo, ksteps, N, layout_type=1):
    from b200ocl import _native
    from b200ocl.ops import _stream
    lib = _native.lib()
    a, b = torch.tensor(PA).cuda().contiguous(), torch.tensor(PB).cuda().contiguous()
    d = torch.full((128, N), float('nan'), device='cuda')
    status = torch.full((1,), -1, dtype=torch.int32, device='cuda')
    rc = lib.b200ocl_selftest_umma_mn(a.data_ptr(), b.data_ptr(), d.data_ptr(), PA.shape[0], PB.shape[0], a_row0, a_lbo, a_sbo,
                                      b_row0, b_lbo, b_sbo, ksteps, N, layout_type, status.data_ptr(), _stream())
    _native.check(rc, 'b200ocl_selftest_umma_mn')
    torch.cuda.synchronize()
    assert int(status) == 0, 'MMA completion barrier timed out'
    return d.cpu().numpy()


def mn_rows(row0, sbo, ksteps):
    return np.concatenate([row0 + g * sbo + np.arange(4) for g in range(2 * ksteps)])


def mn_reference(PA, PB, a_row0, a_lbo, a_sbo, b_row0, b_lbo, b_sbo, ksteps, N):
    D = np.zeros((128, N), dtype=np.float64)
    for j in range(4):
        A = PA[mn_rows(a_row0 + j * a_lbo, a_sbo, ksteps)].astype(np.float64)            # [K, 32 channels]
        for q in range((N + 31) // 32):
            nb = min(32, N - 32 * q)
            B = PB[mn_rows(b_row0 + q * b_lbo, b_sbo, ksteps), :nb].astype(np.float64)   # [K, nb]
            D[32 * j:32 * j + 32, 32 * q:32 * q + nb] = A.T @ B
    return D.astype(np.float32)


@pytest.mark.parametrize('a_row0,a_lbo,a_sbo,b_row0,b_lbo,b_sbo,ksteps,N', [
    (0, 8, 4, 0, 8, 4, 1, 32),        # M blocks in consecutive 8-row groups, K over 8 consecutive rows
    (0, 16, 8, 0, 16, 8, 2, 32),      # both strides doubled: LBO = block stride, SBO = 4-row K group stride
    (0, 1, 4, 0, 8, 4, 1, 32),        # M block j = the strip shifted by j rows (leading byte offset 128)
    (3, 1, 4, 5, 8, 4, 4, 32),        # starts that are not multiples of the 4-row swizzle period
    (77, 1, 4, 2, 8, 4, 16, 32),      # a full 128-position tile at a kernel-row offset
    (2, 1, 4, 1, 136, 4, 3, 48),      # N = 48: second N block 136 rows further
    (1, 1, 4, 3, 136, 4, 16, 160),
])
def test_umma_mn_major_strips(a_row0, a_lbo, a_sbo, b_row0, b_lbo, b_sbo, ksteps, N):
    """MN-major tf32 operands (layout type SWIZZLE_128B_BASE32B) read in place from strips: the addressing wgrad_tc.cu
    relies on -- one 128-byte row per K index, 32-byte chunks XOR-ed with the absolute row index mod 4."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    rs = np.random.RandomState(a_row0 + 7 * a_lbo + N)
    rows_a = a_row0 + 3 * a_lbo + 2 * ksteps * a_sbo + 9
    rows_b = b_row0 + ((N + 31) // 32 - 1) * b_lbo + 2 * ksteps * b_sbo + 7
    PA = (rs.randint(-8, 9, (rows_a, 32)) / 8.0).astype(np.float32)      # exactly representable in TF32
    PB = (rs.randint(-8, 9, (rows_b, 32)) / 8.0).astype(np.float32)
    D = run_mn(PA, PB, a_row0, a_lbo, a_sbo, b_row0, b_lbo, b_sbo, ksteps, N)
    ref = mn_reference(PA, PB, a_row0, a_lbo, a_sbo, b_row0, b_lbo, b_sbo, ksteps, N)
    bad = np.argwhere(D != ref)
    assert bad.size == 0, ('first mismatches (row, col):', bad[:6].tolist(), 'mismatching rows:', sorted(set(bad[:, 0].tolist()))[:40],
                           'cols:', sorted(set(bad[:, 1].tolist()))[:40])


def test_umma_mn_major_needs_the_32_byte_base_swizzle():
    """Measured behaviour this stack builds on: with layout type 2 (SWIZZLE_128B, 16-byte base -- the type of the K-major
    tiles) a kind::tf32 instruction with MN-major operands completes and leaves D all zeros; layout type 1
    (SWIZZLE_128B_BASE32B) is the one that reads the operands."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    rs = np.random.RandomState(5)
    PA = (rs.randint(1, 9, (64, 32)) / 8.0).astype(np.float32)
    PB = (rs.randint(1, 9, (64, 32)) / 8.0).astype(np.float32)
    D2 = run_mn(PA, PB, 0, 8, 4, 0, 8, 4, 1, 32, layout_type=2)
    assert not D2.any(), 'layout type 2 now reads MN-major tf32 operands: revisit umma.cuh'
    D1 = run_mn(PA, PB, 0, 8, 4, 0, 8, 4, 1, 32, layout_type=1)
    np.testing.assert_array_equal(D1, mn_reference(PA, PB, 0, 8, 4, 0, 8, 4, 1, 32))
