"""oracle/trig.py -- TEST INFRASTRUCTURE, not product code.

torch front-end of oracle/trig.c: sin / cos / asin / acos as the explicit fp32 polynomial code that
iris_b200/csrc/trig.cuh restates operation for operation, so that the oracle and the CUDA kernels sample
bit-identical directions (reference call sites: model/brdf.py:28-29,50-51, utils/ops.py:32-44).  No gradient
flows through the samplers in the reference (uniforms and detached roughness in, directions out), so plain
tensor-in / tensor-out functions are enough.

`patched_torch()` swaps torch.sin / cos / asin / acos (and sqrt, which torch's CPU build does not round correctly) for these during a
run of the REFERENCE's own code
(oracle/refharness.py) -- "the reference with this libm" -- which is how the `*_st` golden vectors are made.
"""
from __future__ import annotations

import contextlib
import ctypes

import numpy as np
import torch

from .intersect import _lib as _load

_READY = False


def _lib():
    global _READY
    L = _load()
    if not _READY:
        L.oracle_sincos.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_asin.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.oracle_acos.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.oracle_sqrt.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _READY = True
    return L


def _in(x):
    return np.ascontiguousarray(x.detach().to(torch.float32).numpy())


def sincos(x):
    a = _in(x)
    s, c = np.empty_like(a), np.empty_like(a)
    _lib().oracle_sincos(a.ctypes.data, a.size, s.ctypes.data, c.ctypes.data)
    return torch.from_numpy(s).reshape(x.shape), torch.from_numpy(c).reshape(x.shape)


def sin(x):
    return sincos(x)[0]


def cos(x):
    return sincos(x)[1]


def _unary(name, x):
    a = _in(x)
    y = np.empty_like(a)
    getattr(_lib(), name)(a.ctypes.data, a.size, y.ctypes.data)
    return torch.from_numpy(y).reshape(x.shape)


def asin(x):
    return _unary("oracle_asin", x)


def acos(x):
    return _unary("oracle_acos", x)


def sqrt(x):
    """IEEE square root (torch's CPU kernel is one ulp off on 0.6 % of fp32 inputs; CUDA's and C's are exact)."""
    return _unary("oracle_sqrt", x)


@contextlib.contextmanager
def patched_torch():
    """torch.sin / cos / asin / acos -> the shared definitions, for fp32 CPU tensors without grad (everything
    else falls through to torch's own)."""
    orig = {k: getattr(torch, k) for k in ("sin", "cos", "asin", "acos", "sqrt")}
    mine = {"sin": sin, "cos": cos, "asin": asin, "acos": acos, "sqrt": sqrt}
    base_sqrt = torch.Tensor.sqrt

    def wrap(k):
        def f(x, *a, **kw):
            if torch.is_tensor(x) and x.dtype == torch.float32 and not x.requires_grad and not a and not kw and x.device.type == "cpu":
                return mine[k](x)
            return orig[k](x, *a, **kw)
        return f
    for k in orig:
        setattr(torch, k, wrap(k))

    def tensor_sqrt(self):
        if self.dtype == torch.float32 and not self.requires_grad and self.device.type == "cpu":
            return sqrt(self)
        return base_sqrt(self)
    torch.Tensor.sqrt = tensor_sqrt                       # the reference writes x.sqrt() (model/brdf.py:28,50, model/emitter.py:241)
    try:
        yield
    finally:
        for k, v in orig.items():
            setattr(torch, k, v)
        torch.Tensor.sqrt = base_sqrt
