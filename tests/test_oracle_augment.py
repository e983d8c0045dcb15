"""CPU: oracle/augment.py (SCR's second view with given parameters; kornia itself is absent -- parity unpinned)
checked against independent implementations of its parts: torch's grid_sample for the corner-aligned bilinear
crop and the standard library's colorsys for the HSV operations."""
import colorsys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import augment as oaug


def _identity_params(n, H, W):
    p = np.zeros((n, 12))
    p[:, 2], p[:, 3], p[:, 7], p[:, 8] = W, H, 1, 1
    return p


def test_crop_matches_grid_sample():
    rs = np.random.RandomState(0)
    for (H, W) in [(32, 32), (84, 84), (8, 12)]:
        x = rs.rand(5, 3, H, W)
        p = _identity_params(5, H, W)
        p[:, 0] = rs.randint(0, W // 2, 5)
        p[:, 1] = rs.randint(0, H // 2, 5)
        p[:, 2] = np.array([rs.randint(1, W - int(p[i, 0]) + 1) for i in range(5)])
        p[:, 3] = np.array([rs.randint(1, H - int(p[i, 1]) + 1) for i in range(5)])
        p[::2, 4] = 1
        out = oaug.scr_view(x, p)
        for i in range(5):
            # sampling grid in grid_sample's align_corners=True convention: -1 / +1 are the centres of the corner pixels
            xs = p[i, 0] + np.arange(W) * (p[i, 2] - 1) / max(W - 1, 1)
            ys = p[i, 1] + np.arange(H) * (p[i, 3] - 1) / max(H - 1, 1)
            gx, gy = np.meshgrid(2 * xs / (W - 1) - 1, 2 * ys / (H - 1) - 1)
            grid = torch.from_numpy(np.stack([gx, gy], -1))[None]
            ref = F.grid_sample(torch.from_numpy(x[i:i + 1]), grid, mode='bilinear', padding_mode='border', align_corners=True)[0].numpy()
            if p[i, 4] > 0.5:
                ref = ref[:, :, ::-1]
            np.testing.assert_allclose(out[i], ref, rtol=0, atol=1e-12)


def test_colour_ops_match_colorsys():
    rs = np.random.RandomState(1)
    x = rs.rand(2, 3, 6, 6)
    x[0, :, 0, 0] = 0.5                      # a grey pixel: hue undefined, saturation 0
    x[0, :, 0, 1] = 0.0                      # black
    for order in [(0, 1, 2, 3), (3, 2, 1, 0), (2, 0, 3, 1)]:
        p = _identity_params(2, 6, 6)
        p[:, 5] = 1
        p[:, 6], p[:, 7], p[:, 8], p[:, 9] = [0.2, -0.3], [1.3, 0.7], [0.6, 1.4], [0.08, -0.05]
        p[:, 10] = sum(op << (2 * i) for i, op in enumerate(order))
        out = oaug.scr_view(x, p)
        for n in range(2):
            for yy in range(6):
                for xx in range(6):
                    c = list(x[n, :, yy, xx])
                    for op in order:
                        if op == 0:
                            c = [min(max(v + p[n, 6], 0), 1) for v in c]
                        elif op == 1:
                            c = [min(max(v * p[n, 7], 0), 1) for v in c]
                        else:
                            h, s, v = colorsys.rgb_to_hsv(*c)
                            if op == 2:
                                s = min(max(s * p[n, 8], 0), 1)
                            else:
                                h = (h + p[n, 9]) % 1.0
                            c = list(colorsys.hsv_to_rgb(h, s, v))
                    np.testing.assert_allclose(out[n, :, yy, xx], c, rtol=0, atol=1e-12)


def test_grayscale_and_identity():
    rs = np.random.RandomState(2)
    x = rs.rand(3, 3, 5, 7)
    p = _identity_params(3, 5, 7)
    p[1, 4] = 1
    p[2, 11] = 1
    out = oaug.scr_view(x, p)
    np.testing.assert_array_equal(out[0], x[0])
    np.testing.assert_array_equal(out[1], x[1][:, :, ::-1])
    g = 0.299 * x[2, 0] + 0.587 * x[2, 1] + 0.114 * x[2, 2]
    np.testing.assert_allclose(out[2], np.stack([g, g, g]), rtol=0, atol=1e-15)
