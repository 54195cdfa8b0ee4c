"""TEST INFRASTRUCTURE (checker only): numpy restatement of the a-trous shading-map filter defined in iris_b200/csrc/denoise.cuh.

Stands where bake_shading.py:81,126-131 calls `mitsuba.OptixDenoiser` -- a learned network in an absent third-party library:
PARITY WITH THE REFERENCE'S DENOISER IS UNPINNED (no weights, no golden vectors exist for it).  What this file pins is that the
CUDA filter computes the documented formula (Dammertz et al. 2010):
  level i: taps q = p + 2^i (dx,dy), dx,dy in -2..2, h = (1/16,1/4,3/8,1/4,1/16)
  w = h(dx) h(dy) exp(-|c_p-c_q|^2/(sigma_c 2^-i)^2) max(0,n_p.n_q)^sigma_n exp(-|x_p-x_q|^2/sigma_x^2);  out = sum w c_q / sum w
  zero guide normal: pixel passes through and is never a tap."""
import numpy as np


def atrous(image, normal=None, position=None, iterations=5, sigma_c=1.0, sigma_n=32.0, sigma_x=1.0):
    src = np.asarray(image, np.float32)
    H, W, _ = src.shape
    h = np.array([0.0625, 0.25, 0.375, 0.25, 0.0625], np.float32)
    valid = np.ones((H, W), bool) if normal is None else (np.asarray(normal) != 0).any(-1)
    for i in range(iterations):
        step = 1 << i
        sc = np.float32(sigma_c) * np.float32(2.0 ** -i)
        acc = np.zeros_like(src)
        wsum = np.zeros((H, W), np.float32)
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                oy, ox = dy * step, dx * step
                y0, y1 = max(0, -oy), min(H, H - oy)
                x0, x1 = max(0, -ox), min(W, W - ox)
                if y0 >= y1 or x0 >= x1:
                    continue
                P = (slice(y0, y1), slice(x0, x1))
                Q = (slice(y0 + oy, y1 + oy), slice(x0 + ox, x1 + ox))
                dc = src[Q] - src[P]
                w = h[dx + 2] * h[dy + 2] * np.exp(-(dc * dc).sum(-1) / (sc * sc))
                if normal is not None:
                    w = w * np.maximum((normal[P] * normal[Q]).sum(-1), 0.0) ** np.float32(sigma_n) * valid[Q]
                if position is not None:
                    dp = position[Q] - position[P]
                    w = w * np.exp(-(dp * dp).sum(-1) / np.float32(sigma_x * sigma_x))
                w = w.astype(np.float32)
                acc[P] += src[Q] * w[..., None]
                wsum[P] += w
        out = np.where((wsum > 0)[..., None], acc / np.maximum(wsum, 1e-38)[..., None], src)
        out[~valid] = src[~valid]
        src = out.astype(np.float32)
    return src
