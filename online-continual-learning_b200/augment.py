"""SCR's augmentation pipeline (agents/scr.py:18-24) as one CUDA kernel (csrc/augment.cu).

Random parameters are drawn here, on the host, per sample, following the parameter generators of
the four kornia 0.4.1 modules the reference composes; the kernel applies them.  kornia itself is
absent from the image, so the arithmetic is parity-unpinned -- SCR parity tests inject a fixed
second view instead.
"""
import math

import numpy as np
import torch

from . import _native
from .memory import to_device
from .ops import _need_cuda, _stream


class Identity(torch.nn.Module):
    def forward(self, x):
        return x


def draw_params(n, height, width, rng=None, scale=(0.2, 1.0), ratio=(3. / 4., 4. / 3.), flip_p=0.5,
                jitter=(0.4, 0.4, 0.4, 0.1), jitter_p=0.8, gray_p=0.2):
    """[n,12] float32 parameter block for b200ocl_scr_augment."""
    rng = np.random if rng is None else rng
    out = np.zeros((n, 12), dtype=np.float32)
    # RandomResizedCrop: area ~ U(scale)*H*W, log-uniform aspect; 10 tries, else the whole image
    tries = 10
    area = rng.uniform(scale[0], scale[1], (tries, n)) * height * width
    aspect = np.exp(rng.uniform(math.log(ratio[0]), math.log(ratio[1]), (tries, n)))
    tw, th = np.floor(np.sqrt(area * aspect)), np.floor(np.sqrt(area / aspect))
    ok = (tw >= 1) & (tw <= width) & (th >= 1) & (th <= height)
    first = ok.argmax(axis=0)                      # first successful try per sample
    any_ok = ok.any(axis=0)
    cols = np.arange(n)
    w = np.where(any_ok, tw[first, cols], float(width))
    h = np.where(any_ok, th[first, cols], float(height))
    out[:, 2], out[:, 3] = w, h
    out[:, 0] = np.floor(rng.uniform(0, 1, n) * (width - w + 1))
    out[:, 1] = np.floor(rng.uniform(0, 1, n) * (height - h + 1))
    out[:, 4] = rng.uniform(0, 1, n) < flip_p
    out[:, 5] = rng.uniform(0, 1, n) < jitter_p
    out[:, 6] = rng.uniform(1 - jitter[0], 1 + jitter[0], n) - 1.0         # additive brightness
    out[:, 7] = rng.uniform(1 - jitter[1], 1 + jitter[1], n)
    out[:, 8] = rng.uniform(1 - jitter[2], 1 + jitter[2], n)
    out[:, 9] = rng.uniform(-jitter[3], jitter[3], n)                        # hue shift in turns
    order = rng.permutation(4)                                               # one order per batch
    out[:, 10] = float(sum(int(op) << (2 * i) for i, op in enumerate(order)))
    out[:, 11] = rng.uniform(0, 1, n) < gray_p
    return out


class SCRTransform(torch.nn.Module):
    """RandomResizedCrop(size, scale=(0.2,1)) -> RandomHorizontalFlip -> ColorJitter(0.4,0.4,0.4,0.1,p=0.8)
    -> RandomGrayscale(p=0.2) on a batch of NCHW images."""

    def __init__(self, size, scale=(0.2, 1.0)):
        super().__init__()
        self.size = tuple(size)
        self.scale = scale

    def forward(self, x, params=None):
        _need_cuda(x)
        x = x.detach().to(torch.float32).contiguous()
        n, c, h, w = x.shape
        if c != 3 or (h, w) != self.size:
            raise ValueError('expected [N,3,%d,%d] images' % self.size)
        if params is None:
            params = draw_params(n, h, w, scale=self.scale)
        p = to_device(np.ascontiguousarray(params, dtype=np.float32), x.device)
        out = torch.empty_like(x)
        rc = _native.lib().b200ocl_scr_augment(x.data_ptr(), out.data_ptr(), p.data_ptr(), n, h, w, _stream())
        _native.check(rc, 'b200ocl_scr_augment')
        return out
