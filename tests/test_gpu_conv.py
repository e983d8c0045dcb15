"""GPU: every 3x3 convolution kernel family against an fp64 reference convolution, one layer at a time
(b200ocl_conv_selftest forces the path): CUDA-core kernels, tcgen05 with im2col tiles (conv_tc.cu), tcgen05
fed from a halo patch (conv_tcp.cu); forward and data gradient, raw and accumulate."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PATHS = {'cuda_core': 1, 'tc_im2col': 2, 'tc_patch': 3}


def run_conv(x_nhwc, w, dgrad, path, accumulate=None, train=False):
    from b200ocl import _native
    from b200ocl.ops import _stream
    lib = _native.lib()
    N, H, W, _ = x_nhwc.shape
    cout, cin = w.shape[0], w.shape[1]
    out = torch.zeros(N, H, W, cin if dgrad else cout, device='cuda') if accumulate is None else accumulate.clone()
    stats = torch.full((4 * cout,), float('nan'), device='cuda')
    nbytes = lib.b200ocl_conv_selftest_workspace_bytes(N, cin, cout, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    ws.view(torch.float32).fill_(float('nan'))        # nothing may be read before it is written
    mode = 2 if train else (0 if accumulate is None else 1)
    rc = lib.b200ocl_conv_selftest(x_nhwc.data_ptr(), w.data_ptr(), out.data_ptr(), N, H, W, cin, cout, int(dgrad), path,
                                   mode, stats.data_ptr() if train else None, ws.data_ptr(), nbytes, _stream())
    _native.check(rc, 'b200ocl_conv_selftest')
    torch.cuda.synchronize()
    return (out, stats) if train else out


def reference(x_nhwc, w, dgrad):
    x = x_nhwc.permute(0, 3, 1, 2).double()
    if dgrad:
        y = torch.nn.functional.conv_transpose2d(x, w.double(), padding=1)
    else:
        y = torch.nn.functional.conv2d(x, w.double(), padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


SHAPES = [  # N, H, W, cin, cout
    (10, 32, 32, 20, 20), (110, 32, 32, 20, 20), (7, 16, 16, 40, 40), (110, 16, 16, 40, 40),
    (10, 8, 8, 80, 80), (7, 8, 8, 80, 80), (110, 8, 8, 80, 80), (3, 16, 16, 20, 40), (5, 8, 8, 40, 80),
    (1, 8, 8, 80, 80), (2, 32, 32, 20, 20), (64, 16, 16, 160, 160), (8, 8, 8, 80, 80), (8, 16, 16, 40, 40), (8, 32, 32, 20, 20),
    (10, 4, 4, 160, 160), (110, 4, 4, 160, 160), (7, 4, 4, 160, 160), (5, 21, 21, 40, 40), (3, 11, 11, 80, 80), (2, 6, 6, 160, 160),
]


@pytest.mark.parametrize('path', sorted(PATHS))
@pytest.mark.parametrize('dgrad', [0, 1])
@pytest.mark.parametrize('shape', SHAPES)
def test_conv3x3_matches_fp64(shape, dgrad, path):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    N, H, W, cin, cout = shape
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + H + cin + dgrad)
    w = torch.randn(cout, cin, 3, 3, device='cuda', generator=g) / np.sqrt(9 * cin)
    x = torch.randn(N, H, W, cout if dgrad else cin, device='cuda', generator=g)
    if not dgrad:
        x = torch.relu(x)
    ref = reference(x, w, dgrad)
    got = run_conv(x, w, dgrad, PATHS[path])
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, err
    # accumulate mode adds onto an existing tensor
    base = torch.randn(got.shape, device='cuda', generator=g)
    got2 = run_conv(x, w, dgrad, PATHS[path], accumulate=base)
    err2 = float((got2.double() - (ref + base.double())).abs().max() / ref.abs().max())
    assert err2 < 5e-6, err2


@pytest.mark.parametrize('path', sorted(PATHS))
@pytest.mark.parametrize('shape', SHAPES)
def test_conv3x3_train_statistics(shape, path):
    """Train-mode epilogue: raw output + batch mean / invstd / running statistics (fp64 sums in the kernels)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    N, H, W, cin, cout = shape
    g = torch.Generator(device='cuda').manual_seed(N * 77 + H + cin)
    w = torch.randn(cout, cin, 3, 3, device='cuda', generator=g) / np.sqrt(9 * cin)
    x = torch.relu(torch.randn(N, H, W, cin, device='cuda', generator=g))
    ref = reference(x, w, 0)
    got, stats = run_conv(x, w, 0, PATHS[path], train=True)
    assert float((got.double() - ref).abs().max() / ref.abs().max()) < 5e-6
    z = ref.reshape(-1, cout)
    mean, var = z.mean(0), z.var(0, unbiased=False)
    st = stats.double().reshape(4, cout)
    scale = float(z.abs().max())
    assert float((st[0] - mean).abs().max()) < 2e-6 * scale
    assert float((st[1] - 1.0 / torch.sqrt(var + 1e-5)).abs().max() / (1.0 / torch.sqrt(var + 1e-5)).abs().max()) < 1e-5
    assert float((st[2] - 0.1 * mean).abs().max()) < 2e-6 * scale
    n = z.shape[0]
    unb = var * n / max(n - 1, 1)
    assert float((st[3] - 0.1 * unb).abs().max() / (0.1 * unb).abs().max()) < 1e-5
