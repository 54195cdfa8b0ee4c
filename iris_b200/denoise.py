"""Shading-map denoiser on the device: the post-process of bake_shading.py:81,126-131,190-203 (`mitsuba.OptixDenoiser`).

OptiX's denoiser is a learned network in an absent third-party library, so there is no parity to pin; `atrous` is a deterministic
edge-avoiding a-trous wavelet filter (definition: csrc/denoise.cuh, checker: oracle/denoise.py) that can use the primary-hit
normals / positions the bake already holds as guides.  `iris_b200.compat.mitsuba.OptixDenoiser` calls it."""
from __future__ import annotations

import torch

from . import _capi as C

DEFAULTS = dict(iterations=5, sigma_c=None, sigma_n=32.0, sigma_x=None)


def atrous(image, normal=None, position=None, iterations=5, sigma_c=None, sigma_n=32.0, sigma_x=None):
    """image (H,W,3) float32 CUDA tensor -> filtered copy.  normal / position: optional (H,W,3) guides (zero normal = no hit).
    sigma_c defaults to the image's mean luminance-like level (mean of the positive pixels), sigma_x to 5 % of the guide's extent."""
    if not image.is_cuda:
        raise RuntimeError("iris_b200.denoise.atrous needs CUDA tensors (there is no CPU path)")
    img = image.contiguous().float()
    H, W = int(img.shape[0]), int(img.shape[1])
    nrm = None if normal is None else normal.to(img.device).contiguous().float()
    pos = None if position is None else position.to(img.device).contiguous().float()
    if sigma_c is None:
        m = img[img > 0]
        sigma_c = float(m.mean()) if m.numel() else 1.0
    if sigma_x is None:
        sigma_x = 1.0
        if pos is not None and pos.numel():
            ext = float((pos.reshape(-1, 3).max(0).values - pos.reshape(-1, 3).min(0).values).max())
            sigma_x = max(0.05 * ext, 1e-6)
    out = torch.empty_like(img)
    lib = C.lib()
    ws = torch.empty(max(lib.iris_denoise_workspace_bytes(H, W), 16), dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        C.check(lib.iris_denoise_atrous(C.ptr(img), C.ptr(nrm), C.ptr(pos), H, W, int(iterations), float(sigma_c), float(sigma_n), float(sigma_x),
                                        C.ptr(out), C.ptr(ws), ws.numel(), C.stream_ptr()))
    return out
