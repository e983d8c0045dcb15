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
