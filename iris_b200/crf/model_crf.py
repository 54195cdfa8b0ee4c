"""EmorCRF with the reference's interface (crf/model_crf.py:32-121): per-channel camera response f0 + weight @ basis over the
EMoR basis (Grossberg & Nayar), applied to HDR radiance after the estimator.  forward / its adjoint run on the CUDA path
(iris_crf_forward / iris_crf_backward); `inverse` (the LDR -> HDR map slf_bake.py:127 and slf_refine.py:97 apply to every view before
the SLF scatter) builds the 3 x 1024 inverse table with a few torch ops on the device (`get_inv_crf`: monotonised response,
interpolated on its own non-uniform grid) and applies it with the same interpolation kernel; `initialize_weight` fits the weights
to a given response; the regularisers are one-line torch expressions.

The EMoR tables are data of the reference checkout (crf/emor.txt, crf/invemor.txt: blocks of a `name =` line followed by 256
lines of 4 numbers); they are read from a path, never bundled here."""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _capi as C


def parse_emor_file(path):
    """crf/emor.py:19-38 -> (names, vectors (n_curves, 1024)): E, f0, h(1)..h(25)."""
    lines = [ln.strip() for ln in open(path)]
    stride = 1 + 256
    names, vectors = [], []
    for i in range(len(lines) // stride):
        names.append(lines[i * stride].split("=")[0].strip())
        vectors.append(np.array(" ".join(lines[i * stride + 1:(i + 1) * stride]).split(), np.float32))
    return np.array(names), np.stack(vectors)


class _CRFApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hdr, exposure, crf):
        hdr = hdr.contiguous().float()
        exposure = exposure.reshape(-1).contiguous().float()
        crf = crf.contiguous().float()
        n = hdr.shape[0]
        if exposure.numel() not in (1, n):
            raise ValueError("EmorCRF: exposure must hold one value or one per row (got %d for %d rows)" % (exposure.numel(), n))
        stride = 0 if exposure.numel() == 1 else 1
        ldr = torch.empty_like(hdr)
        with torch.cuda.device(hdr.device):
            C.check(C.lib().iris_crf_forward(C.ptr(hdr), C.ptr(exposure), stride, C.ptr(crf), crf.shape[1], n, C.ptr(ldr), C.stream_ptr()))
        ctx.save_for_backward(hdr, exposure, crf)
        ctx.stride = stride
        return ldr

    @staticmethod
    def backward(ctx, d_ldr):
        hdr, exposure, crf = ctx.saved_tensors
        d_ldr = d_ldr.contiguous().float()
        d_hdr = torch.empty_like(hdr) if ctx.needs_input_grad[0] else None
        d_crf = torch.zeros_like(crf) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(hdr.device):
            C.check(C.lib().iris_crf_backward(C.ptr(hdr), C.ptr(exposure), ctx.stride, C.ptr(crf), crf.shape[1], C.ptr(d_ldr), hdr.shape[0],
                                              C.ptr(d_hdr), C.ptr(d_crf), C.stream_ptr()))
        return d_hdr, None, d_crf


class EmorCRF(nn.Module):
    def __init__(self, dim=11, emor_path=None, tables=None):
        """tables: optional (f0 (1024,), basis (>=dim,1024)) arrays; otherwise read from emor_path (default ./crf/emor.txt, where the
        reference resolves it from the working directory, crf/emor.py:14-15)."""
        super().__init__()
        self.dim = dim
        if tables is None:
            _, vectors = parse_emor_file(emor_path or os.path.join(os.getcwd(), "crf", "emor.txt"))
            f0, basis = vectors[1], vectors[2:2 + dim]
        else:
            f0, basis = np.asarray(tables[0], np.float32), np.asarray(tables[1], np.float32)[:dim]
        self.register_buffer("f0", torch.as_tensor(f0).float()[None])
        self.register_buffer("basis", torch.as_tensor(basis).float())
        self.weight = nn.Parameter(torch.zeros(3, dim))

    def get_crf(self):
        return self.f0 + self.weight @ self.basis

    def forward(self, hdr, exposure):
        """(n,3) HDR radiance, exposure (n,1) or scalar tensor -> (n,3) LDR."""
        exposure = exposure if torch.is_tensor(exposure) else torch.tensor([float(exposure)], device=hdr.device)
        return _CRFApply.apply(hdr, exposure.to(hdr.device), self.get_crf())

    def get_inv_crf(self):
        """crf/model_crf.py:45-55 (with mono_increase_constraint, :22-30): per channel, the response is made monotone (finite
        differences shifted to be non-negative, renormalised, integrated) and inverted by linear interpolation on its own knots."""
        crf = self.get_crf().detach()
        n = crf.shape[1]
        x = torch.linspace(0, 1, n, device=crf.device)
        diff = crf[:, 1:] - crf[:, :-1]
        dmin = diff.min(dim=1, keepdim=True).values
        diff = diff + torch.where(dmin < 0, -dmin, torch.zeros_like(dmin))
        diff = diff / diff.sum(dim=1, keepdim=True)
        mono = torch.cat([torch.zeros(3, 1, device=crf.device), torch.cumsum(diff, 1)], 1).contiguous()
        q = x.expand(3, n).contiguous()
        right = torch.searchsorted(mono, q).clamp_max(n - 1)                 # first knot >= x (torch.bucketize, right=False)
        left = (right - 1).clamp_min(0)
        dl = (q - mono.gather(1, left)).clamp_min(0)
        dr = (mono.gather(1, right) - q).clamp_min(0)
        both = (dl == 0) & (dr == 0)
        dl = torch.where(both, torch.ones_like(dl), dl)
        dr = torch.where(both, torch.ones_like(dr), dr)
        return (x[left] * dr + x[right] * dl) / (dl + dr)

    def inverse(self, ldr, exposure):
        """crf/model_crf.py:88-105: (n,3) LDR -> HDR = inv_crf(clip(ldr,0,1)) / exposure."""
        exposure = exposure if torch.is_tensor(exposure) else torch.tensor([float(exposure)], device=ldr.device)
        one = torch.ones(1, device=ldr.device)
        with torch.no_grad():
            hdr = _CRFApply.apply(ldr.detach(), one, self.get_inv_crf())
        return hdr / exposure.to(ldr.device)

    def cal_weight_fitting_crf(self, crf):
        """crf/model_crf.py:61-66: least-squares weights (3,dim) of the response `crf` (3,n_bins) over the basis."""
        B = self.basis.detach().cpu().double().numpy().T
        y = (np.asarray(crf, np.float64) - self.f0.detach().cpu().double().numpy()).T
        return np.linalg.solve(B.T @ B, B.T @ y).T

    def initialize_weight(self, crf):
        self.weight = nn.Parameter(torch.as_tensor(self.cal_weight_fitting_crf(crf), dtype=torch.float32).to(self.weight.device))

    def reg_weight(self):
        return torch.mean(self.weight ** 2)

    def reg_monotonically_increasing(self):
        crf = self.get_crf()
        return torch.sum(torch.relu(-(crf[:, 1:] - crf[:, :-1])))

    def reg_smoothness(self):
        crf = self.get_crf()
        return torch.mean((crf[:, :-2] + crf[:, 2:] - 2 * crf[:, 1:-1]) ** 2)
