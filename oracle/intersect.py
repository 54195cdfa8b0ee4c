"""oracle/intersect.py -- TEST INFRASTRUCTURE, not product code.

ctypes front-end of oracle/intersect.c (the CPU closest-hit restatement of reference
utils/path_tracing.py:17-48).  Build with `make -C oracle`.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        L.oracle_scene_create.restype = ctypes.c_void_p
        L.oracle_scene_create.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        L.oracle_scene_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_intersect.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6
        _LIB = L
    return _LIB


class OracleScene:
    """Triangle mesh + CPU closest-hit.  mode: 'bvh' (default) or 'brute'."""

    def __init__(self, vertices, faces):
        self.vertices = np.ascontiguousarray(vertices, np.float32)
        self.faces = np.ascontiguousarray(faces, np.int32)
        self._h = _lib().oracle_scene_create(self.vertices.ctypes.data, len(self.vertices),
                                             self.faces.ctypes.data, len(self.faces))

    def __del__(self):
        try:
            if self._h:
                _lib().oracle_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def intersect_raw(self, o, d, mode="bvh"):
        """o,d (N,3) float32 arrays -> dict(t, prim, uv, p, n, tie)."""
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = len(o)
        out = {
            "t": np.empty(n, np.float32), "prim": np.empty(n, np.int32), "uv": np.empty((n, 2), np.float32),
            "p": np.empty((n, 3), np.float32), "n": np.empty((n, 3), np.float32), "tie": np.empty(n, np.uint8),
        }
        _lib().oracle_intersect(self._h, 0 if mode == "brute" else 1, o.ctypes.data, d.ctypes.data, n,
                                out["t"].ctypes.data, out["prim"].ctypes.data, out["uv"].ctypes.data,
                                out["p"].ctypes.data, out["n"].ctypes.data, out["tie"].ctypes.data)
        return out

    def ray_intersect(self, xs, ds, mode="bvh"):
        """Same return convention as reference utils/path_tracing.py:17-48, on torch CPU tensors:
        positions (N,3), normals (N,3) flipped toward -ds, uvs (N,2), idx (N,) int64 (-1 = miss), valid (N,) bool."""
        import torch
        r = self.intersect_raw(xs.detach().numpy(), ds.detach().numpy(), mode)
        idx = torch.from_numpy(r["prim"].astype(np.int64))
        return (torch.from_numpy(r["p"]), torch.from_numpy(r["n"]), torch.from_numpy(r["uv"]), idx, idx >= 0)
