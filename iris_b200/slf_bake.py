"""Voxel surface-light-field bake on the device: the three passes of the reference's slf_bake.py:70-145 over the training views,
written against `iris_slf_*` (include/iris_b200.h).

    baker = SLFBaker(H=256, device="cuda:0")
    for view in views:  baker.observe_bounds(positions, valid)            # pass 1   slf_bake.py:73-85
    baker.set_bounds_from_observed(dataset="synthetic")                   #          :87-93  (x1.1 rules)
    for view in views:  baker.mark(positions, valid)                      # pass 2   :96-113
    baker.build_index()                                                   #          VoxelSLF.__init__, model/slf.py:26-37
    for view in views:  baker.scatter_add(positions, radiance, valid)     # pass 3   :118-135
    vslf = baker.finalize()                                               #          :138-145  -> dict in the reference's vslf.npz layout

`SLFBaker.from_vslf(state_dict)` is the entry of slf_refine.py:85-108: the occupancy mask and bounds of an existing vslf.npz are kept,
the radiance is re-accumulated from scratch (there with the trained CRF's `inverse`, iris_b200.crf.EmorCRF.inverse) and averaged.

`finalize()` returns {'mask', 'voxel_min', 'voxel_max', 'weight': {'inds', 'radiance', 'count'}} -- what `torch.save(..., 'vslf.npz')`
writes in the reference and what SLFEmitter / iris_b200.core.ShadingTables.set_slf load; `device_tables()` hands the int32 index
grid and the radiance table to the estimators without the round trip through int64.
"""
from __future__ import annotations

import ctypes

import torch

from . import _capi as C


class SLFBaker:
    def __init__(self, H=256, device="cuda:0"):
        self.H = int(H)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SLFBaker runs on a CUDA device (no CPU path)")
        self._state = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._first = True
        self.voxel_min = self.voxel_max = None
        self.occupancy = torch.zeros(self.H ** 3, dtype=torch.int32, device=self.device)
        self.inds = None
        self.n_cells = 0
        self.sum = self.count = None

    @classmethod
    def from_vslf(cls, state_dict, device="cuda:0"):
        """slf_refine.py:85-87: VoxelSLF(state_dict['mask'], state_dict['voxel_min'], state_dict['voxel_max']) -- same voxels, same
        index grid (rank of the occupied voxels in raster order, model/slf.py:29-32), radiance and count back to zero.  state_dict is the
        dict `torch.load('vslf.npz')` returns (or finalize()'s)."""
        mask = state_dict["mask"]
        mask = mask if torch.is_tensor(mask) else torch.as_tensor(mask)
        b = cls(int(mask.shape[0]), device)
        if tuple(mask.shape) != (b.H, b.H, b.H):
            raise ValueError("vslf mask must be (H,H,H)")
        b.set_bounds(state_dict["voxel_min"], state_dict["voxel_max"])
        b.occupancy = mask.reshape(-1).to(device=b.device, dtype=torch.int32).contiguous()
        b._first = False
        b.build_index()
        return b

    @staticmethod
    def _prep(positions, valid):
        positions = positions.reshape(-1, 3).float().contiguous()
        if valid is not None:
            valid = valid.reshape(-1).to(torch.uint8).contiguous()
        return positions, valid

    # ---- pass 1
    def observe_bounds(self, positions, valid=None):
        positions, valid = self._prep(positions, valid)
        with torch.cuda.device(self.device):
            C.check(C.lib().iris_slf_bounds(C.ptr(positions), C.ptr(valid), positions.shape[0], 1 if self._first else 0, C.ptr(self._state), C.stream_ptr()))
        self._first = False

    def observed_bounds(self):
        lo, hi = self._state[:2].tolist()
        return lo, hi

    def set_bounds(self, voxel_min, voxel_max):
        self.voxel_min, self.voxel_max = float(voxel_min), float(voxel_max)

    def set_bounds_from_observed(self, dataset="synthetic"):
        """slf_bake.py:87-93, in fp32 like the reference's 0-dim tensors.  The reference's running min / max start from 1000. and 0.0
        (:73-74), i.e. the bounds are min(1000, min p) and max(0, max p): kept, so that a scene on one side of the origin gets the same
        grid."""
        lo, hi = (torch.tensor(v, dtype=torch.float32) for v in self.observed_bounds())
        lo, hi = torch.minimum(lo, torch.tensor(1000.0)), torch.maximum(hi, torch.tensor(0.0))
        if dataset in ("synthetic", "real"):
            lo, hi = 1.1 * lo, 1.1 * hi
        else:
            c = lo + hi
            lo, hi = c + (lo - c) * 1.1, c + (hi - c) * 1.1
        self.set_bounds(lo.item(), hi.item())

    def _grid(self):
        if self.voxel_min is None:
            raise RuntimeError("set the voxel bounds first (set_bounds / set_bounds_from_observed)")
        vmin = torch.tensor(self.voxel_min, dtype=torch.float32)
        return float(vmin), float(torch.tensor(self.voxel_max, dtype=torch.float32) - vmin)

    # ---- pass 2
    def mark(self, positions, valid=None):
        positions, valid = self._prep(positions, valid)
        vmin, vrange = self._grid()
        with torch.cuda.device(self.device):
            C.check(C.lib().iris_slf_mark(C.ptr(positions), C.ptr(valid), positions.shape[0], vmin, vrange, self.H, C.ptr(self.occupancy), C.stream_ptr()))

    def build_index(self):
        lib = C.lib()
        ws = torch.empty(lib.iris_slf_index_workspace_bytes(self.H), dtype=torch.uint8, device=self.device)
        self.inds = torch.empty(self.H ** 3, dtype=torch.int32, device=self.device)
        n = C.c_i64()
        with torch.cuda.device(self.device):
            C.check(lib.iris_slf_index(C.ptr(self.occupancy), self.H, C.ptr(self.inds), ctypes.byref(n), C.ptr(ws), ws.numel(), C.stream_ptr()))
        self.n_cells = int(n.value)
        self.sum = torch.zeros(max(self.n_cells, 1), 3, device=self.device)[: self.n_cells]
        self.count = torch.zeros(max(self.n_cells, 1), dtype=torch.int32, device=self.device)[: self.n_cells]
        return self.n_cells

    # ---- pass 3
    def scatter_add(self, positions, radiance, valid=None):
        if self.inds is None:
            raise RuntimeError("build_index() first")
        positions, valid = self._prep(positions, valid)
        radiance = radiance.reshape(-1, 3).float().contiguous()
        vmin, vrange = self._grid()
        with torch.cuda.device(self.device):
            C.check(C.lib().iris_slf_accumulate(C.ptr(positions), C.ptr(valid), C.ptr(radiance), positions.shape[0], vmin, vrange, self.H, C.ptr(self.inds),
                                                C.ptr(self.sum), C.ptr(self.count), C.stream_ptr()))

    def finalize(self):
        with torch.cuda.device(self.device):
            C.check(C.lib().iris_slf_finalize(C.ptr(self.sum), C.ptr(self.count), self.n_cells, C.stream_ptr()))
        H = self.H
        return {"mask": (self.occupancy > 0).view(H, H, H), "voxel_min": self.voxel_min, "voxel_max": self.voxel_max,
                "weight": {"inds": self.inds.view(H, H, H).long(), "radiance": self.sum, "count": self.count.long()}}

    def device_tables(self):
        """(inds int32 (H,H,H), radiance (n_cells,3)) as the estimators read them; call after finalize()."""
        return self.inds.view(self.H, self.H, self.H), self.sum
