"""oracle/slf.py -- TEST INFRASTRUCTURE.  torch (CPU) restatement of the surface-light-field bake:

    scene bounds + 1.1 scaling         slf_bake.py:70-93
    occupancy histogram / mask         slf_bake.py:96-113
    VoxelSLF (index grid, scatter_add) model/slf.py:16-61
    mean pooling                       slf_bake.py:138

Pinned: tests/golden/make_golden.py runs the bake on seeded points through the reference's OWN VoxelSLF class (model/slf.py imported
from /root/reference) and stores the result in tests/golden/slf.npz; tests/test_oracle_cpu.py checks this restatement against it."""
import torch


def bounds(views, dataset="synthetic"):
    voxel_min, voxel_max = 1000.0, 0.0
    for pos, valid in views:
        if not valid.any():
            continue
        p = pos[valid]
        voxel_min = min(voxel_min, p.min())
        voxel_max = max(voxel_max, p.max())
    if dataset in ("synthetic", "real"):
        voxel_min, voxel_max = 1.1 * voxel_min, 1.1 * voxel_max
    else:
        c = voxel_min + voxel_max
        voxel_min, voxel_max = c + (voxel_min - c) * 1.1, c + (voxel_max - c) * 1.1
    return voxel_min, voxel_max            # 0-dim fp32 tensors, like the reference's


def occupancy(views, voxel_min, voxel_max, H):
    hist = torch.zeros(H ** 3)
    for pos, valid in views:
        if not valid.any():
            continue
        p = (pos[valid] - voxel_min) / (voxel_max - voxel_min)
        p = (p * H).long().clamp(0, H - 1)
        inds = p[..., 0] + p[..., 1] * H + p[..., 2] * H * H
        hist.scatter_add_(0, inds, torch.ones_like(inds).float())
    return (hist.reshape(H, H, H) > 0)


class VoxelSLF:
    def __init__(self, mask, voxel_min, voxel_max):
        H = mask.shape[0]
        self.H, self.voxel_min, self.voxel_max = H, voxel_min, voxel_max
        kk, jj, ii = torch.where(mask)
        self.inds = -torch.ones(H, H, H, dtype=torch.long)
        self.inds[kk, jj, ii] = torch.arange(len(ii))
        self.radiance = torch.zeros(len(ii), 3)
        self.count = torch.zeros(len(ii), dtype=torch.long)

    def spatial_idx(self, x):
        x_ = (x - self.voxel_min) / (self.voxel_max - self.voxel_min)
        x_ = (x_ * self.H).long().clamp(0, self.H - 1)
        return self.inds[x_[..., 2], x_[..., 1], x_[..., 0]]

    def scatter_add(self, x, radiance):
        idx = self.spatial_idx(x)
        self.radiance.scatter_add_(0, idx[..., None].expand_as(radiance), radiance)
        self.count.scatter_add_(0, idx, torch.ones_like(idx))


def bake(views, radiances, H, dataset="synthetic", slf_cls=VoxelSLF):
    """views: list of (positions (n,3), valid (n,) bool); radiances: list of (n,3).  Returns the vslf.npz dict of slf_bake.py:140-145."""
    voxel_min, voxel_max = bounds(views, dataset)
    mask = occupancy(views, voxel_min, voxel_max, H)
    vslf = slf_cls(mask, voxel_min.item(), voxel_max.item())
    for (pos, valid), rad in zip(views, radiances):
        if not valid.any():
            continue
        vslf.scatter_add(pos[valid], rad[valid])
    vslf.radiance = vslf.radiance / vslf.count[..., None].float().clamp_min(1)
    return {"mask": mask, "voxel_min": voxel_min.item(), "voxel_max": voxel_max.item(),
            "weight": {"inds": vslf.inds, "radiance": vslf.radiance, "count": vslf.count}}
