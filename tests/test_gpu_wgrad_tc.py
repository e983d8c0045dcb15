"""GPU: the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu, both operands read in place from strips) against an fp64
weight gradient of torch's conv2d -- the cuDNN call behind loss.backward() for nn.Conv2d (models/resnet.py:11-12)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def wgrad_tc(x_nhwc, dz_nhwc):
    from b200ocl import _native
    from b200ocl.ops import _stream, _workspace
    lib = _native.lib()
    N, H, W, cin = x_nhwc.shape
    cout = dz_nhwc.shape[3]
    nbytes = lib.b200ocl_wgrad_tc_selftest_workspace_bytes(N, H, W, cin, cout)
    assert nbytes > 0
    ws = _workspace(nbytes, x_nhwc.device)
    dw = torch.full((cout, cin, 3, 3), float('nan'), device=x_nhwc.device)
    rc = lib.b200ocl_wgrad_tc_selftest(x_nhwc.data_ptr(), dz_nhwc.data_ptr(), dw.data_ptr(), N, H, W, cin, cout, ws.data_ptr(),
                                       ws.numel(), _stream())
    _native.check(rc, 'b200ocl_wgrad_tc_selftest')
    torch.cuda.synchronize()
    return dw


@pytest.mark.parametrize('N,H,W,cin,cout', [
    (1, 4, 4, 160, 160), (3, 4, 4, 160, 160), (10, 8, 8, 80, 80), (20, 16, 16, 40, 40), (10, 32, 32, 20, 20),
    (110, 32, 32, 20, 20), (110, 4, 4, 160, 160), (37, 16, 16, 40, 40), (7, 11, 11, 80, 160), (5, 21, 21, 40, 80),
    (2, 37, 5, 8, 12), (220, 8, 8, 80, 80),
])
def test_wgrad_tc_matches_fp64(N, H, W, cin, cout):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    g = torch.Generator().manual_seed(N * 1000 + H * 10 + cin)
    x = torch.relu(torch.randn(N, cin, H, W, generator=g))                  # post-ReLU activations
    dz = torch.randn(N, cout, H, W, generator=g) / (N * H * W) ** 0.5
    w = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, padding=1).backward(dz.double())
    ref = w.grad
    got = wgrad_tc(x.permute(0, 2, 3, 1).contiguous().cuda(), dz.permute(0, 2, 3, 1).contiguous().cuda()).cpu().double()
    assert torch.isfinite(got).all()
    err = float((got - ref).abs().max() / ref.abs().max())
    rms = float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())
    print('N=%d %dx%d %d->%d  max %.2e  rms %.2e' % (N, H, W, cin, cout, err, rms))
    assert err < 5e-6 and rms < 2e-6, (err, rms)
    again = wgrad_tc(x.permute(0, 2, 3, 1).contiguous().cuda(), dz.permute(0, 2, 3, 1).contiguous().cuda()).cpu().double()
    assert torch.equal(again, got)                                          # deterministic
