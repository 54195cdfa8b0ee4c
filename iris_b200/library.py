"""torch.library registration of the C-ABI entry points (namespace `iris_b200`).

Each op is a thin call into libiris_b200.so (iris_b200.core) with an explicit schema, a fake (meta) implementation that only
computes output shapes -- so FakeTensorMode / torch.compile tracing see through the calls without a GPU -- and, for the two
differentiable paths, `register_autograd` formulas that pair the forward with the library's hand-written adjoint:

    iris_b200::intersect(scene, o, d)                                -> t, prim, uv, p, n            ray_intersect, utils/path_tracing.py:17-48
    iris_b200::single_forward(radiance, params, rays, scene, tables, spp, seed, lane_offset, U, rec_mode)
                                                                     -> L, record, encoded           path_tracing_single, :320-407
    iris_b200::single_backward(dL, record, encoded, tables, spp, n_rad_rows, n_params)
                                                                     -> d_radiance, d_params         its replay adjoint
    iris_b200::field_forward(params, position, tables, keep)         -> mat, encoded                 NGPBRDF.forward, model/brdf.py:243-260
    iris_b200::field_backward(d_mat, position, encoded, tables, n_params) -> d_params
    iris_b200::bake(position, normal, wo, scene, tables, mode, roughness, spp, seed, U) -> out0, out1   bake_shading.py:108-123,168-188
    iris_b200::path_tracing(rays, scene, tables, spp, indir_depth, seed, U) -> L                      path_tracing, :214-318

Opaque objects (the BVH scene, the device tables) cross the dispatcher as integer handles into a weak registry: an op schema can
only carry tensors and scalars, and the objects stay owned by whoever created them.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import core

_OBJECTS = weakref.WeakValueDictionary()


def handle(obj) -> int:
    """Integer handle of a Scene / ShadingTables for the op schemas (valid while the object is alive)."""
    h = id(obj)
    _OBJECTS[h] = obj
    return h


def _obj(h: int):
    try:
        return _OBJECTS[h]
    except KeyError:
        raise RuntimeError("iris_b200: stale scene / tables handle (the object was garbage-collected)") from None


def _sampler(seed: int, lane_offset: int, U: Optional[Tensor]):
    return core.Sampler(U=U, seed=seed, lane_offset=lane_offset)


# ------------------------------------------------------------------------------------------------ ray_intersect
@torch.library.custom_op("iris_b200::intersect", mutates_args=(), device_types="cuda")
def intersect(scene: int, o: Tensor, d: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    return _obj(scene).intersect_raw(o, d)


@intersect.register_fake
def _(scene, o, d):
    n = o.shape[0]
    return (o.new_empty(n), o.new_empty(n, dtype=torch.int32), o.new_empty(n, 2), o.new_empty(n, 3), o.new_empty(n, 3))


# ------------------------------------------------------------------------------------------------ path_tracing_single
# rec_mode: 0 inference, 1 record for the emitter gradient, 2 record + encoded field inputs (field gradient as well)
@torch.library.custom_op("iris_b200::single_forward", mutates_args=(), device_types="cuda")
def single_forward(radiance: Tensor, params: Optional[Tensor], rays: Tensor, scene: int, tables: int, spp: int, seed: int, lane_offset: int,
                   U: Optional[Tensor], rec_mode: int) -> Tuple[Tensor, Tensor, Tensor]:
    L, rec = core.single_forward(_obj(scene), _obj(tables), rays, spp, _sampler(seed, lane_offset, U), want_record=rec_mode > 0,
                                 want_encoded=rec_mode == 2, workspace=_workspace(rays.device, core.C.lib().iris_single_workspace_bytes(rays.shape[0], spp)))
    empty = rays.new_empty(0, dtype=torch.uint8)
    enc = getattr(rec, "encoded", None) if rec is not None else None
    return L, (rec if rec is not None else empty), (enc if enc is not None else empty)


@single_forward.register_fake
def _(radiance, params, rays, scene, tables, spp, seed, lane_offset, U, rec_mode):
    B = rays.shape[0]
    u8 = dict(dtype=torch.uint8)
    return (rays.new_empty(B, 3), rays.new_empty(96 * B * spp if rec_mode > 0 else 0, **u8), rays.new_empty(128 * B * spp if rec_mode == 2 else 0, **u8))


@torch.library.custom_op("iris_b200::single_backward", mutates_args=(), device_types="cuda")
def single_backward(dL: Tensor, record: Tensor, encoded: Tensor, tables: int, spp: int, n_rad_rows: int, n_params: int) -> Tuple[Tensor, Tensor]:
    T = _obj(tables)
    d_par = torch.zeros(n_params, device=dL.device, dtype=torch.float32) if n_params > 0 else None
    ws = _workspace(dL.device, core.C.lib().iris_single_workspace_bytes(dL.shape[0], spp)) if n_params > 0 else None
    g = core.single_backward(T, dL, spp, record, want_radiance=n_rad_rows > 0, d_params=d_par, workspace=ws,
                             encoded=encoded if encoded.numel() else None)
    d_rad = torch.zeros(max(n_rad_rows, 0), 3, device=dL.device, dtype=torch.float32)      # (F,3): only rows [0,K) receive gradient
    if n_rad_rows > 0:
        d_rad[:T.K] = g
    return d_rad, (d_par if d_par is not None else dL.new_empty(0))


@single_backward.register_fake
def _(dL, record, encoded, tables, spp, n_rad_rows, n_params):
    return dL.new_empty(max(n_rad_rows, 0), 3), dL.new_empty(max(n_params, 0))


def _single_setup(ctx, inputs, output):
    radiance, params, rays, scene, tables, spp, seed, lane_offset, U, rec_mode = inputs
    L, rec, enc = output
    ctx.tables, ctx.spp = tables, spp
    ctx.n_rad = radiance.shape[0] if (radiance is not None and radiance.requires_grad and rec_mode > 0) else 0
    ctx.n_par = params.numel() if (params is not None and params.requires_grad and rec_mode == 2) else 0
    ctx.rad_shape = None if radiance is None else radiance.shape
    ctx.par_shape = None if params is None else params.shape
    ctx.save_for_backward(rec, enc)


def _single_backward(ctx, dL, d_rec, d_enc):
    rec, enc = ctx.saved_tensors
    d_rad = d_par = None
    if ctx.n_rad or ctx.n_par:
        a, b = torch.ops.iris_b200.single_backward(dL.contiguous(), rec, enc, ctx.tables, ctx.spp, ctx.n_rad, ctx.n_par)
        d_rad = a.view(ctx.rad_shape) if ctx.n_rad else None
        d_par = b.view(ctx.par_shape) if ctx.n_par else None
    return d_rad, d_par, None, None, None, None, None, None, None, None


single_forward.register_autograd(_single_backward, setup_context=_single_setup)


# ------------------------------------------------------------------------------------------------ NGPBRDF.forward
@torch.library.custom_op("iris_b200::field_forward", mutates_args=(), device_types="cuda")
def field_forward(params: Tensor, position: Tensor, tables: int, keep: bool) -> Tuple[Tensor, Tensor]:
    if keep:
        return core.field_forward(_obj(tables), position, want_encoded=True)
    return core.field_forward(_obj(tables), position), position.new_empty(0, 64, dtype=torch.float16)


@field_forward.register_fake
def _(params, position, tables, keep):
    n = position.shape[0]
    return position.new_empty(n, 5), position.new_empty(n if keep else 0, 64, dtype=torch.float16)


@torch.library.custom_op("iris_b200::field_backward", mutates_args=(), device_types="cuda")
def field_backward(d_mat: Tensor, position: Tensor, encoded: Tensor, tables: int, n_params: int) -> Tensor:
    d = torch.zeros(n_params, device=d_mat.device, dtype=torch.float32)
    n = position.shape[0]
    ws = _workspace(d_mat.device, core.C.lib().iris_field_backward_workspace_bytes(n))
    core.field_backward(_obj(tables), position, d_mat, d, workspace=ws, encoded=encoded if encoded.numel() else None)
    return d


@field_backward.register_fake
def _(d_mat, position, encoded, tables, n_params):
    return d_mat.new_empty(n_params)


def _field_setup(ctx, inputs, output):
    params, position, tables, keep = inputs
    ctx.tables, ctx.n_params, ctx.shape = tables, params.numel(), params.shape
    ctx.save_for_backward(position, output[1])


def _field_backward(ctx, d_mat, d_enc):
    position, enc = ctx.saved_tensors
    d = torch.ops.iris_b200.field_backward(d_mat.contiguous(), position, enc, ctx.tables, ctx.n_params)
    return d.view(ctx.shape), None, None, None


field_forward.register_autograd(_field_backward, setup_context=_field_setup)


# ------------------------------------------------------------------------------------------------ forward-only estimators
@torch.library.custom_op("iris_b200::bake", mutates_args=(), device_types="cuda")
def bake(position: Tensor, normal: Tensor, wo: Optional[Tensor], scene: int, tables: int, mode: int, roughness: float, spp: int, seed: int,
         U: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    out = core.bake(_obj(scene), _obj(tables), mode, roughness, position, normal, wo, spp, _sampler(seed, 0, U))
    return (out, position.new_empty(0, 3)) if mode == 0 else out


@bake.register_fake
def _(position, normal, wo, scene, tables, mode, roughness, spp, seed, U):
    B = position.shape[0]
    return position.new_empty(B, 3), position.new_empty(B if mode == 1 else 0, 3)


@torch.library.custom_op("iris_b200::path_tracing", mutates_args=(), device_types="cuda")
def path_tracing(rays: Tensor, scene: int, tables: int, spp: int, indir_depth: int, seed: int, U: Optional[Tensor]) -> Tensor:
    ws = _workspace(rays.device, core.C.lib().iris_wave_workspace_bytes(rays.shape[0] * spp))
    return core.path_tracing(_obj(scene), _obj(tables), rays, spp, indir_depth, _sampler(seed, 0, U), workspace=ws)


@path_tracing.register_fake
def _(rays, scene, tables, spp, indir_depth, seed, U):
    return rays.new_empty(rays.shape[0], 3)


# ------------------------------------------------------------------------------------------------ workspaces
# One scratch buffer per (device, stream), grown on demand and reused by every call: launches are stream-ordered on the caller's stream, and a
# workspace is dead when its call returns (records / encoded inputs, which live from forward to backward, are separate tensors).
_WS = {}


def _workspace(device, n_bytes):
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    w = _WS.get(key)
    if w is None or w.numel() < n_bytes or w.device != device:
        w = torch.empty(max(int(n_bytes), 16), dtype=torch.uint8, device=device)
        _WS[key] = w
    return w
