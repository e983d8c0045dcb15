"""_PinnedRing offset logic on the CPU (uploads ending exactly on a half / on nbytes)."""
import numpy as np
import pytest
import torch

from b200ocl import memory


class _Ev:
    def __init__(self, log, half):
        self.log, self.half = log, half

    def synchronize(self):
        self.log.append(('wait', self.half))


def _ring(nbytes, log):
    r = memory._PinnedRing(nbytes)
    r.buf = torch.empty(nbytes, dtype=torch.uint8)                    # no pin_memory without a driver
    r._alloc = lambda shape, dtype, device: torch.empty(shape, dtype=dtype)
    r._copy = lambda out, view: out.view(torch.uint8).reshape(-1).copy_(view)

    def record():
        log.append(('record', r.cur))
        return _Ev(log, r.cur)
    r._record = record
    return r


@pytest.mark.parametrize('sizes', [[80, 40, 120], [128], [64, 64], [16] * 9, [256, 16, 240]])
def test_ring_exact_boundaries(sizes):
    log = []
    r = _ring(512, log)                       # halves of 256 bytes
    rs = np.random.RandomState(0)
    for rep in range(200):
        for n in sizes:
            a = rs.randint(0, 255, n).astype(np.uint8)
            out = r.upload(a, 'cpu')
            assert np.array_equal(out.numpy(), a)
            assert 0 <= r.used <= r.half
    # every entry into a half that was left before waits for that half's event
    left = set()
    for kind, h in log:
        if kind == 'record':
            left.add(h)
        else:
            assert h in left


def test_ring_random_sizes_never_overflow():
    log = []
    r = _ring(1 << 12, log)
    rs = np.random.RandomState(1)
    for _ in range(60000):
        n = int(rs.choice([8, 80, 40, 120, 800, 1200, 2048, 16, 24]))
        a = np.full(n, n & 255, dtype=np.uint8)
        out = r.upload(a, 'cpu')
        assert out.numel() == n and int(out[0]) == (n & 255)


def test_ring_typed_upload_roundtrip():
    r = _ring(1024, [])
    a = np.arange(13, dtype=np.int64) * 7
    assert np.array_equal(r.upload(a, 'cpu').numpy(), a)
    f = np.linspace(0, 1, 9, dtype=np.float32).reshape(3, 3)
    assert np.array_equal(r.upload(f, 'cpu').numpy(), f)
