"""EmorCRF with the reference's interface (crf/model_crf.py:32-121): per-channel camera response f0 + weight @ basis over the
EMoR basis (Grossberg & Nayar), applied to HDR radiance after the estimator.  forward / its adjoint run on the CUDA path
(iris_crf_forward / iris_crf_backward); the regularisers and the inverse are small torch expressions.

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

    def reg_weight(self):
        return torch.mean(self.weight ** 2)

    def reg_monotonically_increasing(self):
        crf = self.get_crf()
        return torch.sum(torch.relu(-(crf[:, 1:] - crf[:, :-1])))

    def reg_smoothness(self):
        crf = self.get_crf()
        return torch.mean((crf[:, :-2] + crf[:, 2:] - 2 * crf[:, 1:-1]) ** 2)
