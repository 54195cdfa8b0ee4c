"""The reference's estimator surface (utils/path_tracing.py) on the CUDA path: same names, argument order and return
conventions, so `from utils.path_tracing import ...` in the drivers can be pointed here (iris_b200.compat.install)."""
from __future__ import annotations

import torch

from .. import ops


def ray_intersect(scene, xs, ds):
    """utils/path_tracing.py:17-48."""
    return ops.ray_intersect(scene, xs, ds)


def path_tracing_single(scene, emitter_net, material_net, rays_o, rays_d, dx_du, dy_dv, spp):
    """utils/path_tracing.py:320-407: (B,3) radiance, differentiable wrt emitter_net.radiance and material_net.mlp.params."""
    if rays_o.shape[0] == 0:
        return torch.zeros_like(rays_o)
    return ops.path_tracing_single(scene, emitter_net, material_net, rays_o, rays_d, dx_du, dy_dv, spp)


def path_tracing(scene, emitter_net, material_net, rays_o, rays_d, dx_du, dy_dv, spp, indir_depth):
    """utils/path_tracing.py:214-318."""
    if rays_o.shape[0] == 0:
        return torch.zeros_like(rays_o)
    return ops.path_tracing(scene, emitter_net, material_net, rays_o, rays_d, dx_du, dy_dv, spp, indir_depth)


def path_tracing_det_diff(scene, emitter_net, material_net, positions, wis, normals, uvs, triangle_idxs, spp, indir_depth):
    """utils/path_tracing.py:50-124."""
    return ops.path_tracing_det(scene, emitter_net, material_net, None, positions, wis, normals, triangle_idxs, spp, indir_depth)


def path_tracing_det_spec(scene, emitter_net, material_net, roughness_level, positions, wis, normals, uvs, triangle_idxs, spp, indir_depth):
    """utils/path_tracing.py:127-212."""
    return ops.path_tracing_det(scene, emitter_net, material_net, roughness_level, positions, wis, normals, triangle_idxs, spp, indir_depth)


def trace_indirect(scene, emitter_net, material_net, position, wo, normal, indir_depth):
    """utils/path_tracing.py:409-502."""
    return ops.trace_indirect(scene, emitter_net, material_net, position, wo, normal, indir_depth)
