"""Oracle: SCR's second view (agents/scr.py:18-24: RandomResizedCrop -> RandomHorizontalFlip -> ColorJitter ->
RandomGrayscale, kornia 0.4.1 modules) applied with GIVEN per-sample parameters, in numpy float64.

PARITY UNPINNED: kornia is a third-party dependency of the reference (requirements.txt: kornia==0.4.1) that is
neither vendored in /root/reference nor installed in this image, and the reference holds no test vectors for it.
This module restates the published definitions of the four modules -- crop_by_boxes (corner-to-corner box
mapping, bilinear resampling), hflip, the colour operations of kornia.enhance (additive brightness,
multiplicative contrast, HSV saturation scaling and hue rotation, each clamped to [0,1], applied in a random
order drawn once per batch) and rgb_to_grayscale (ITU-R 601 weights) -- so that the CUDA kernel's arithmetic
is checked against an independent implementation; it is NOT checked against kornia itself.  Anchors in the
reference: the call site agents/scr.py:18-24 (module list and arguments) and agents/scr.py:47-48,55-56 (the
transform is applied to the concatenated batch).  Test infrastructure only -- see oracle/__init__.py.

Parameter block per sample (the layout of online-continual-learning_b200/augment.py::draw_params):
0 x0, 1 y0, 2 crop_w, 3 crop_h, 4 flip, 5 jitter_on, 6 brightness delta, 7 contrast factor, 8 saturation factor,
9 hue shift (turns), 10 order code (four 2-bit op ids, first op in the low bits), 11 gray.
"""
import numpy as np


def _rgb_to_hsv(c):
    r, g, b = c
    mx, mn = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    d = mx - mn
    s = np.where(mx > 0, d / np.where(mx > 0, mx, 1.0), 0.0)
    dd = np.where(d > 0, d, 1.0)
    h = np.where(mx == r, (g - b) / dd, np.where(mx == g, 2.0 + (b - r) / dd, 4.0 + (r - g) / dd)) / 6.0
    h = np.where(d > 0, h - np.floor(h), 0.0)
    return h, s, mx


def _hsv_to_rgb(h, s, v):
    h6 = h * 6.0
    i = np.floor(h6)
    f = h6 - i
    p, q, t = v * (1 - s), v * (1 - s * f), v * (1 - s * (1 - f))
    k = i.astype(np.int64) % 6
    r = np.choose(k, [v, q, p, p, t, v])
    g = np.choose(k, [t, v, v, q, p, p])
    b = np.choose(k, [p, p, t, v, v, q])
    return np.stack([r, g, b])


def scr_view(x, params):
    """x [N,3,H,W] in [0,1], params [N,12] -> the augmented batch (float64)."""
    x = np.asarray(x, dtype=np.float64)
    params = np.asarray(params, dtype=np.float64)
    n, _, H, W = x.shape
    out = np.empty_like(x)
    for i in range(n):
        p = params[i]
        # crop box (x0, y0, w, h) resampled to H x W: output corner pixels sit on the box corners
        ox = np.arange(W, dtype=np.float64)
        if p[4] > 0.5:
            ox = W - 1 - ox                                   # flipping the result = reading mirrored columns
        sx = p[0] + (ox * (p[2] - 1) / (W - 1) if W > 1 else 0 * ox)
        sy = p[1] + (np.arange(H, dtype=np.float64) * (p[3] - 1) / (H - 1) if H > 1 else np.zeros(H))
        fx, fy = np.clip(sx, 0, W - 1), np.clip(sy, 0, H - 1)
        x0, y0 = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
        x1, y1 = np.minimum(x0 + 1, W - 1), np.minimum(y0 + 1, H - 1)
        ax, ay = (fx - x0)[None, None, :], (fy - y0)[None, :, None]
        im = x[i]
        top = (1 - ax) * im[:, y0][:, :, x0] + ax * im[:, y0][:, :, x1]
        bot = (1 - ax) * im[:, y1][:, :, x0] + ax * im[:, y1][:, :, x1]
        c = (1 - ay) * top + ay * bot
        if p[5] > 0.5:
            code = int(p[10])
            for _ in range(4):
                op = code & 3
                code >>= 2
                if op == 0:
                    c = np.clip(c + p[6], 0, 1)
                elif op == 1:
                    c = np.clip(c * p[7], 0, 1)
                else:
                    h, s, v = _rgb_to_hsv(c)
                    if op == 2:
                        s = np.clip(s * p[8], 0, 1)
                    else:
                        h = h + p[9]
                        h = h - np.floor(h)
                    c = _hsv_to_rgb(h, s, v)
        if p[11] > 0.5:
            g = 0.299 * c[0] + 0.587 * c[1] + 0.114 * c[2]
            c = np.stack([g, g, g])
        out[i] = c
    return out
