"""VoxelSLF -- voxel-grid surface light field (reference model/slf.py:16-70): dense (H,H,H) index grid into a compact
(n_occ,3) radiance table, nearest-voxel query.  Buffers `inds`, `radiance`, `count` keep the reference's state_dict layout
(the device tables narrow `inds` to int32 on upload, iris_b200.core.ShadingTables.set_slf)."""
from __future__ import annotations

import torch
import torch.nn as nn


class VoxelSLF(nn.Module):
    def __init__(self, mask, voxel_min, voxel_max):
        super().__init__()
        H = mask.shape[0]
        self.H = H
        self.voxel_min = voxel_min
        self.voxel_max = voxel_max
        flat = mask.reshape(-1).bool()
        n_occ = int(flat.sum())
        table = torch.full((H * H * H,), -1, dtype=torch.long)
        table[flat] = torch.arange(n_occ)               # raster order over (z,y,x) == the reference's torch.where(mask) order
        self.register_buffer("inds", table.view(H, H, H))
        self.register_buffer("radiance", torch.zeros(n_occ, 3))
        self.register_buffer("count", torch.zeros(n_occ, dtype=torch.long))

    def spatial_idx(self, x):
        """model/slf.py:41-54: table row of the voxel containing x (-1 = empty)."""
        cell = ((x - self.voxel_min) / (self.voxel_max - self.voxel_min) * self.H).long().clamp_(0, self.H - 1)
        return self.inds.view(-1)[(cell[..., 2] * self.H + cell[..., 1]) * self.H + cell[..., 0]]

    def scatter_add(self, x, radiance):
        """model/slf.py:56-61 (baking)."""
        idx = self.spatial_idx(x)
        self.radiance.scatter_add_(0, idx[..., None].expand_as(radiance), radiance)
        self.count.scatter_add_(0, idx, torch.ones_like(idx))

    def forward(self, x):
        idx = self.spatial_idx(x)
        rad = self.radiance[idx]
        rad[idx == -1] = 0
        return {"rgb": rad}
