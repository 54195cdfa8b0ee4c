"""Operator layer: the reference's call surface on top of the C ABI.

 * `tables_for(emitter_net, material_net)` builds (and caches on the modules) the device tables the kernels read, from ANY
   object that carries the reference's buffers -- the reference's own `SLFEmitter` / `NGPBRDF` instances or the mirrors in
   `iris_b200.model` (duck typing on `is_emitter`, `emitter_vertices`, `emitter_area`, `radiance`, `slf.inds`, `slf.radiance`,
   `slf.voxel_min/max`, `mlp.params`, `voxel_min/max`).
 * the estimators, the field and the ray cast go through `torch.ops.iris_b200.*` (iris_b200/library.py: torch.library custom ops with
   schemas, fake implementations and registered adjoints over the same C ABI): gradients flow to `emitter_net.radiance` (rows
   [0,K), the reference's quirk) and to `material_net.mlp.params`.
 * uniforms: Philox keyed by a seed drawn from torch's global generator per call (reproducible under `torch.manual_seed`,
   like the reference's `torch.rand`), or an explicit `(N,D)` buffer via `iris_b200.ops.inject_samples(U)` for parity runs.
"""
from __future__ import annotations

import contextlib
import ctypes

import numpy as np
import torch

from . import _capi as C
from . import core
from . import library as L_

_INJECT = []


@contextlib.contextmanager
def inject_samples(U):
    """Run the estimators inside with the explicit (N,D) uniform buffer U (identical injected sequences)."""
    _INJECT.append(U)
    try:
        yield
    finally:
        _INJECT.pop()


def next_sampler(device):
    if _INJECT:
        return core.Sampler(U=_INJECT[-1].to(device))
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # advances torch's global generator like torch.rand would
    from . import dist as idist
    return core.Sampler(seed=idist.rank_seed(seed))              # ranks seeded alike still draw independent sample streams


# ------------------------------------------------------------------------------------------------ scene handle
def as_scene(scene):
    """Accept an iris_b200.core.Scene, or the object returned by the mitsuba compat shim's load_dict."""
    if isinstance(scene, core.Scene):
        return scene
    s = getattr(scene, "iris_scene", None)
    if s is None:
        raise TypeError("scene must be an iris_b200 Scene (use iris_b200.compat.mitsuba.load_dict or iris_b200.core.Scene)")
    return s


# ------------------------------------------------------------------------------------------------ tables
def _ver(t):
    return (t.data_ptr(), t._version, tuple(t.shape), str(t.device))


def tables_for(emitter_net, material_net=None, device=None):
    dev = torch.device(device) if device is not None else emitter_net.radiance.device
    cache = emitter_net.__dict__.setdefault("_iris_cache", {})
    T = cache.get("tables")
    if T is None or T.device != dev:
        T = core.ShadingTables(dev)
        cache.clear()
        cache["tables"] = T
    key = (_ver(emitter_net.is_emitter), _ver(emitter_net.emitter_vertices), _ver(emitter_net.emitter_area))
    if cache.get("emitter_key") != key:
        T.set_emitter(emitter_net.is_emitter, emitter_net.emitter_vertices, emitter_net.emitter_area, emitter_net.radiance)
        cache["emitter_key"] = key
    r = emitter_net.radiance.detach()
    T.t["radiance"] = r if (r.device == dev and r.dtype == torch.float32 and r.is_contiguous()) else r.to(dev, torch.float32).contiguous()
    slf = emitter_net.slf
    key = (_ver(slf.inds), _ver(slf.radiance), float(slf.voxel_min), float(slf.voxel_max))
    if cache.get("slf_key") != key:
        T.set_slf(slf.inds, slf.radiance, slf.voxel_min, slf.voxel_max)
        cache["slf_key"] = key
    if material_net is not None and hasattr(material_net, "mlp"):
        p = material_net.mlp.params
        key = (_ver(p), float(material_net.voxel_min), float(material_net.voxel_max))
        if cache.get("field_key") != key:
            T.set_field(p, material_net.voxel_min, material_net.voxel_max)
            cache["field_key"] = key
    return T


# ------------------------------------------------------------------------------------------------ ray_intersect
def ray_intersect(scene, xs, ds):
    """utils/path_tracing.py:17-48: positions, normals (flipped toward -ds), uvs, idx (int64, -1 = miss), valid."""
    sc = as_scene(scene)
    shape = xs.shape[:-1]
    t, prim, uv, p, n = torch.ops.iris_b200.intersect(L_.handle(sc), xs.reshape(-1, 3).contiguous().float(), ds.reshape(-1, 3).contiguous().float())
    idx = prim.long()
    return p.reshape(*shape, 3), n.reshape(*shape, 3), uv.reshape(*shape, 2), idx.reshape(shape), (idx >= 0).reshape(shape)


# ------------------------------------------------------------------------------------------------ path_tracing_single
def _single(scene, tables, radiance, params, rays, spp, sampler):
    """torch.ops.iris_b200.single_forward (iris_b200/library.py): forward kernel + registered replay adjoint.  Gradients flow to
    `radiance` (rows [0,K), the reference's quirk) and to `params` (material_net.mlp.params)."""
    need_rad = radiance is not None and radiance.requires_grad and torch.is_grad_enabled()
    need_par = params is not None and params.requires_grad and torch.is_grad_enabled()
    rec_mode = 2 if need_par else (1 if need_rad else 0)
    L, _, _ = torch.ops.iris_b200.single_forward(radiance, params, rays, L_.handle(scene), L_.handle(tables), int(spp), sampler.seed, sampler.lane_offset,
                                                 sampler.U, rec_mode)
    return L


def pack_rays(rays_o, rays_d, dx_du, dy_dv):
    return torch.cat([rays_o, rays_d, dx_du, dy_dv], -1).float().contiguous()


def path_tracing_single(scene, emitter_net, material_net, rays_o, rays_d, dx_du, dy_dv, spp):
    """utils/path_tracing.py:320-407 as one fused forward (+ replay adjoint through autograd)."""
    sc = as_scene(scene)
    T = tables_for(emitter_net, material_net, rays_o.device)
    params = material_net.mlp.params if hasattr(material_net, "mlp") else None
    return _single(sc, T, emitter_net.radiance, params, pack_rays(rays_o, rays_d, dx_du, dy_dv), spp, next_sampler(rays_o.device))


def path_tracing(scene, emitter_net, material_net, rays_o, rays_d, dx_du, dy_dv, spp, indir_depth):
    """utils/path_tracing.py:214-318.  Forward only: the reference itself calls it under no_grad (render.py:171-176,
    train_brdf_crf.py:360-370) and detaches everything past the first bounce (:313-315)."""
    T = tables_for(emitter_net, material_net, rays_o.device)
    smp = next_sampler(rays_o.device)
    return torch.ops.iris_b200.path_tracing(pack_rays(rays_o, rays_d, dx_du, dy_dv), L_.handle(as_scene(scene)), L_.handle(T), int(spp), int(indir_depth), smp.seed, smp.U)


def path_tracing_det(scene, emitter_net, material_net, roughness_level, positions, wis, normals, triangle_idxs, spp, indir_depth):
    """utils/path_tracing.py:50-124 (roughness_level None) and :127-212."""
    T = tables_for(emitter_net, material_net, positions.device)
    if positions.shape[0] == 0:
        z = torch.zeros_like(positions)
        return z if roughness_level is None else (z, z.clone())
    mode = 0 if roughness_level is None else 1
    return core.path_tracing_det(as_scene(scene), T, mode, 0.0 if roughness_level is None else float(roughness_level), positions, wis, normals,
                                 triangle_idxs, spp, indir_depth, next_sampler(positions.device))


def trace_indirect(scene, emitter_net, material_net, position, wo, normal, indir_depth):
    """utils/path_tracing.py:409-502."""
    T = tables_for(emitter_net, material_net, position.device)
    if position.shape[0] == 0:
        return torch.zeros_like(position)
    return core.trace_indirect(as_scene(scene), T, position, wo, normal, indir_depth, next_sampler(position.device))


# ------------------------------------------------------------------------------------------------ bake
def bake_diffuse(scene, emitter_net, position, normal, spp):
    """The chunk loop of bake_shading.py:105-123 as one launch: Ld (B,3)."""
    T = tables_for(emitter_net, None, position.device)
    smp = next_sampler(position.device)
    return torch.ops.iris_b200.bake(position.contiguous().float(), normal.contiguous().float(), None, L_.handle(as_scene(scene)), L_.handle(T), 0, 1.0, int(spp),
                                    smp.seed, smp.U)[0]


def bake_specular(scene, emitter_net, position, wo, normal, roughness, spp):
    """The chunk loop of bake_shading.py:165-188 for one roughness level: (Ls0, Ls1)."""
    T = tables_for(emitter_net, None, position.device)
    smp = next_sampler(position.device)
    return torch.ops.iris_b200.bake(position.contiguous().float(), normal.contiguous().float(), wo.contiguous().float(), L_.handle(as_scene(scene)), L_.handle(T),
                                    1, float(roughness), int(spp), smp.seed, smp.U)


# ------------------------------------------------------------------------------------------------ BRDF field
def field(material_net, position):
    """NGPBRDF.forward (model/brdf.py:243-260): dict(albedo (N,3), roughness (N,1), metallic (N,1))."""
    T = _field_tables(material_net, position.device)
    p = material_net.mlp.params
    shape = position.shape[:-1]
    keep = bool(p.requires_grad and torch.is_grad_enabled())            # the adjoint reuses the encoded inputs
    mat, _ = torch.ops.iris_b200.field_forward(p, position.reshape(-1, 3).float().contiguous(), L_.handle(T), keep)
    return {"albedo": mat[:, 0:3].reshape(*shape, 3), "roughness": mat[:, 3:4].reshape(*shape, 1), "metallic": mat[:, 4:5].reshape(*shape, 1)}


# ------------------------------------------------------------------------------------------------ training-step shading (train_brdf_crf.py)
def _field_tables(material_net, dev):
    cache = material_net.__dict__.setdefault("_iris_cache", {})
    T = cache.get("tables")
    p = material_net.mlp.params
    key = (_ver(p), float(material_net.voxel_min), float(material_net.voxel_max))
    if T is None or T.device != dev or cache.get("field_key") != key:
        T = core.ShadingTables(dev).set_field(p, material_net.voxel_min, material_net.voxel_max)
        cache["tables"], cache["field_key"] = T, key
    return T


class BrdfShading(torch.autograd.Function):
    """field forward -> shading from the baked maps, and their adjoints, as ONE autograd node: the (n,5) material tensor is
    returned as well, so regularisers written in torch on albedo / roughness / metallic add their gradient to the same d_mat
    before the single field adjoint runs."""

    @staticmethod
    def forward(ctx, params, tables, position, diffuse, specular0, specular1):
        mat, enc = core.field_forward(tables, position, want_encoded=True)
        L = core.brdf_shading_forward(mat, diffuse, specular0, specular1)
        ctx.tables = tables
        ctx.n_params = params.numel()
        ctx.save_for_backward(position, mat, diffuse, specular0, specular1, enc)
        return L, mat

    @staticmethod
    def backward(ctx, dL, d_mat_up):
        position, mat, diffuse, specular0, specular1, enc = ctx.saved_tensors
        d_mat = d_mat_up.contiguous().float().clone() if d_mat_up is not None else None
        d_mat = core.brdf_shading_backward(mat, diffuse, specular0, specular1, dL, d_mat)
        d = torch.zeros(ctx.n_params, device=dL.device, dtype=torch.float32)
        core.field_backward(ctx.tables, position, d_mat, d, encoded=enc)
        return d, None, None, None, None, None


def brdf_shading(material_net, positions, diffuse, specular0, specular1):
    """The shading block of train_brdf_crf.py:193-206 fused with NGPBRDF.forward.  Returns (L (n,3), mat dict) where the dict has the
    reference's keys (albedo (n,3), roughness (n,1), metallic (n,1)), all differentiable w.r.t. material_net.mlp.params."""
    position = positions.reshape(-1, 3).float().contiguous()
    T = _field_tables(material_net, position.device)
    L, mat = BrdfShading.apply(material_net.mlp.params, T, position, diffuse, specular0, specular1)
    return L, {"albedo": mat[:, 0:3], "roughness": mat[:, 3:4], "metallic": mat[:, 4:5]}


class _LerpSpecular(torch.autograd.Function):
    @staticmethod
    def forward(ctx, specular, roughness):
        n = specular.shape[0]
        mat = torch.ones(n, 5, device=specular.device)              # albedo = metallic = 1  ->  kd = 0, ks = 1
        mat[:, 3] = roughness.reshape(-1)
        zero3, zeroS = torch.zeros(n, 3, device=specular.device), torch.zeros_like(specular)
        ctx.save_for_backward(mat, zero3, specular, zeroS)
        return core.brdf_shading_forward(mat, zero3, specular, zeroS)

    @staticmethod
    def backward(ctx, dL):
        mat, zero3, specular, zeroS = ctx.saved_tensors
        d_mat = core.brdf_shading_backward(mat, zero3, specular, zeroS, dL)
        return None, d_mat[:, 3:4]


def lerp_specular(specular, roughness):
    """utils/ops.py:99-119 on the CUDA kernel (gradient to roughness; the baked maps are data)."""
    return _LerpSpecular.apply(specular.float().contiguous(), roughness.float())
