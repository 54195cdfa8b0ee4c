"""Host-side objects over the C ABI: Scene (mesh + BVH in HBM), ShadingTables (emitter / SLF / BRDF-field tables in the
layout IrisShadeParams wants) and thin launch helpers.  torch supplies device memory and the current stream only.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _capi as C

N_MLP = 9216
N_LEVELS = 32


def field_levels():
    """(scale f32[32], res u32[32], size u32[32], offset u32[32], n_entries) as the library lays the hash grid out."""
    sc = np.empty(N_LEVELS, np.float32)
    res, size, off = (np.empty(N_LEVELS, np.uint32) for _ in range(3))
    n = C.lib().iris_field_levels(sc.ctypes.data, res.ctypes.data, size.ctypes.data, off.ctypes.data)
    return sc, res, size, off, int(n)


class Scene:
    """One triangle mesh with its 8-wide compressed BVH on a CUDA device.

    Stands where the reference holds a Mitsuba scene (`mitsuba.load_dict`, train_emitter.py:57-63); prim index = row of `faces`.
    """

    def __init__(self, vertices, faces, device=0, builder="auto"):
        """builder: 1 / "lbvh" = on the device (Morton hierarchy + binned-SAH treelets + SAH-optimal 8-wide collapse, 32 ms per million
        triangles), 0 / "sah" = host binned SAH (0.7-1 s per million), "auto" = device build, host build if the device tree is deeper than
        the traversal stack allows (pathological duplicates).  Hits are identical whatever the builder; the device tree traces as fast
        as the host one or faster, also on scan-like irregular meshes (profiles/r2t_*)."""
        C.require_cuda()
        builder = {"sah": 0, "lbvh": 1}.get(builder, builder)
        v = np.ascontiguousarray(np.asarray(vertices, np.float32).reshape(-1, 3))
        f = np.ascontiguousarray(np.asarray(faces, np.int32).reshape(-1, 3))
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = C.lib().iris_scene_create(v.ctypes.data, len(v), f.ctypes.data, len(f), self.device.index, 1 if builder == "auto" else int(builder), ctypes.byref(h))
            if rc == -1 and builder == "auto" and b"deeper" in C.lib().iris_last_error():
                rc = C.lib().iris_scene_create(v.ctypes.data, len(v), f.ctypes.data, len(f), self.device.index, 0, ctypes.byref(h))
            C.check(rc)
        self._h = h
        self.n_faces = len(f)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and C is not None and getattr(C, "_LIB", None) is not None:      # (module globals may be gone at interpreter exit)
            C._LIB.iris_scene_destroy(h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def stats(self):
        s = C.IrisSceneStats()
        C.check(C.lib().iris_scene_stats(self._h, ctypes.byref(s)))
        return {k: (list(getattr(s, k)) if k.startswith("bounds") else getattr(s, k)) for k, _ in s._fields_}

    def intersect_raw(self, o, d):
        """o,d (N,3) float32 CUDA tensors -> t (N), prim (N) int32, uv (N,2), p (N,3), n (N,3)."""
        o = o.contiguous().float()
        d = d.contiguous().float()
        n = o.shape[0]
        dev = o.device
        t = torch.empty(n, device=dev)
        prim = torch.empty(n, dtype=torch.int32, device=dev)
        uv = torch.empty(n, 2, device=dev)
        p = torch.empty(n, 3, device=dev)
        nr = torch.empty(n, 3, device=dev)
        with torch.cuda.device(dev):
            C.check(C.lib().iris_intersect(self._h, C.ptr(o), C.ptr(d), n, C.ptr(t), C.ptr(prim), C.ptr(uv), C.ptr(p), C.ptr(nr), C.stream_ptr()))
        return t, prim, uv, p, nr


class Sampler:
    """Uniform-sample source of a launch: an explicit (N,D) CUDA buffer (parity runs) or Philox (seed, lane_offset)."""

    def __init__(self, U=None, seed=0, lane_offset=0):
        self.U = None if U is None else U.contiguous().float()
        self.seed = int(seed)
        self.lane_offset = int(lane_offset)

    def c(self):
        s = C.IrisSampler()
        s.U = None if self.U is None else self.U.data_ptr()
        s.stride = 0 if self.U is None else self.U.shape[1]
        s.seed = self.seed
        s.lane_offset = self.lane_offset
        return s


def sampler_fill(seed, lane_offset, n, dims, device):
    out = torch.empty(n, dims, device=device)
    with torch.cuda.device(device):
        C.check(C.lib().iris_sampler_fill(int(seed), int(lane_offset), n, dims, C.ptr(out), C.stream_ptr()))
    return out


class ShadingTables:
    """Device tables behind IrisShadeParams, built from the reference's on-disk / state_dict formats.

    emitter: dict(is_emitter (F) bool, emitter_vertices (K,3,3), emitter_area (K), emitter_radiance (F|K,3))  [extract_emitter_ldr.py:109-115]
    slf:     dict(voxel_min, voxel_max, weight=dict(inds (H,H,H) int64, radiance (n_occ,3)))               [slf_bake.py:140-145]
    field:   flat fp32 tcnn parameter vector [mlp 9216 | grid] + the voxel bounds NGPBRDF maps positions with   [model/brdf.py:243-260]
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self.t = {}
        self.K = 0
        self.F = 0
        self.H = 0
        self.slf_vmin = self.slf_range = 0.0
        self.field_vmin = self.field_range = 0.0

    def _dev(self, x, dtype):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=self.device, dtype=dtype).contiguous()

    def set_emitter(self, is_emitter, emitter_vertices, emitter_area, radiance):
        is_em = torch.as_tensor(np.asarray(is_emitter) if not torch.is_tensor(is_emitter) else is_emitter).bool().cpu()
        F = len(is_em)
        K = int(is_em.sum())
        eof = torch.full((F,), -1, dtype=torch.int32)
        eof[is_em] = torch.arange(K, dtype=torch.int32)
        self.F, self.K = F, K
        self.t["emitter_of_face"] = eof.to(self.device)
        self.t["face_of_emitter"] = torch.arange(F, dtype=torch.int32)[is_em].to(self.device)
        self.t["emitter_vertices"] = self._dev(emitter_vertices, torch.float32)
        area = self._dev(emitter_area, torch.float32)
        self.t["emitter_area"] = area
        ones = torch.ones_like(area)
        pdf = ones / ones.abs().sum().clamp_min(1e-12)                     # NF.normalize(p=1), model/emitter.py:170
        self.t["emitter_pdf"] = pdf.contiguous()
        self.t["emitter_cdf"] = pdf.cumsum(-1).contiguous()                  # fp32 cumsum table
        self.set_radiance(radiance)
        return self

    def set_radiance(self, radiance):
        r = radiance if torch.is_tensor(radiance) else torch.as_tensor(np.asarray(radiance))
        self.t["radiance"] = r.detach().to(device=self.device, dtype=torch.float32).contiguous()
        return self

    def set_slf(self, inds, radiance, voxel_min, voxel_max):
        inds = inds if torch.is_tensor(inds) else torch.as_tensor(np.asarray(inds))
        self.H = int(inds.shape[0])
        self.t["slf_inds"] = inds.to(device=self.device, dtype=torch.int32).contiguous()       # int64 -> int32 (128 -> 64 MiB at H=256)
        self.t["slf_radiance"] = self._dev(radiance, torch.float32)
        self.slf_vmin = float(np.float32(voxel_min))
        self.slf_range = float(np.float32(float(voxel_max) - float(voxel_min)))
        return self

    def set_field(self, params, voxel_min, voxel_max):
        p = params.detach().to(device=self.device, dtype=torch.float32)
        self.t["mlp_f16"] = p[:N_MLP].half().contiguous()
        self.t["grid_f16"] = p[N_MLP:].half().contiguous()
        self.field_vmin = float(np.float32(voxel_min))
        self.field_range = float(np.float32(float(voxel_max) - float(voxel_min)))
        return self

    def c(self):
        P = C.IrisShadeParams()
        g = lambda k: (self.t[k].data_ptr() if k in self.t else None)
        for k in ("emitter_of_face", "face_of_emitter", "emitter_vertices", "emitter_area", "emitter_pdf", "emitter_cdf", "radiance",
                  "slf_inds", "slf_radiance", "grid_f16", "mlp_f16"):
            setattr(P, k, g(k))
        P.n_emitters, P.n_faces, P.slf_H = self.K, self.F, self.H
        P.slf_vmin, P.slf_range = self.slf_vmin, self.slf_range
        P.field_vmin, P.field_range = self.field_vmin, self.field_range
        return P

    @classmethod
    def from_dicts(cls, device, emitter_dict, slf_dict, params=None, field_bounds=None):
        T = cls(device)
        T.set_emitter(emitter_dict["is_emitter"], emitter_dict["emitter_vertices"], emitter_dict["emitter_area"], emitter_dict["emitter_radiance"])
        T.set_slf(slf_dict["weight"]["inds"], slf_dict["weight"]["radiance"], slf_dict["voxel_min"], slf_dict["voxel_max"])
        if params is not None:
            vmin, vmax = field_bounds if field_bounds is not None else (slf_dict["voxel_min"], slf_dict["voxel_max"])
            T.set_field(params, vmin, vmax)
        return T


# ----------------------------------------------------------------------------------------------- launches
def bake(scene, tables, mode, roughness, position, normal, wo, spp, sampler):
    """bake_shading.py inner loops: mode 0 -> Ld (B,3); mode 1 -> (Ls0, Ls1)."""
    position = position.contiguous().float()
    normal = normal.contiguous().float()
    wo = None if wo is None else wo.contiguous().float()
    B = position.shape[0]
    dev = position.device
    out0 = torch.empty(B, 3, device=dev)
    out1 = torch.empty(B, 3, device=dev) if mode == 1 else None
    P, S = tables.c(), sampler.c()
    with torch.cuda.device(dev):
        C.check(C.lib().iris_bake(scene.handle, ctypes.byref(P), mode, float(roughness), C.ptr(position), C.ptr(normal), C.ptr(wo), B, int(spp),
                                  ctypes.byref(S), C.ptr(out0), C.ptr(out1), C.stream_ptr()))
    return out0 if mode == 0 else (out0, out1)


def bsdf_sample(mode, u, wo, normal, mat=None, roughness=0.0):
    """BaseBRDF.sample_diffuse (mode 0) / sample_specular (1) / sample_brdf (2), model/brdf.py:78-210, one lane per row.
    u (n,2) [mode 2: (n,3) = sample1 | sample2]; returns wi (n,3), pdf (n), w0 (n,3), w1 (n,3) [mode 1 only, else None]."""
    u = u.contiguous().float()
    normal = normal.contiguous().float()
    wo = None if wo is None else wo.contiguous().float()
    mat = None if mat is None else mat.contiguous().float()
    n, dev = normal.shape[0], normal.device
    wi = torch.empty(n, 3, device=dev)
    pdf = torch.empty(n, device=dev)
    w0 = torch.empty(n, 3, device=dev)
    w1 = torch.empty(n, 3, device=dev) if mode == 1 else None
    with torch.cuda.device(dev):
        C.check(C.lib().iris_bsdf_sample(int(mode), C.ptr(u), u.shape[1] if u.dim() == 2 else 0, C.ptr(wo), C.ptr(normal), C.ptr(mat), float(roughness), n,
                                         C.ptr(wi), C.ptr(pdf), C.ptr(w0), C.ptr(w1), C.stream_ptr()))
    return wi, pdf, w0, w1


def field_forward(tables, position, want_encoded=False):
    """NGPBRDF.forward: mat (n,5).  want_encoded: also return the (n,64) fp16 hash-grid features for field_backward(encoded=...)."""
    position = position.contiguous().float()
    n = position.shape[0]
    mat = torch.empty(n, 5, device=position.device)
    enc = torch.empty(n, 64, dtype=torch.float16, device=position.device) if want_encoded else None
    P = tables.c()
    with torch.cuda.device(position.device):
        C.check(C.lib().iris_field_forward(ctypes.byref(P), C.ptr(position), n, C.ptr(mat), C.ptr(enc), C.stream_ptr()))
    return (mat, enc) if want_encoded else mat


def field_backward(tables, position, d_mat, d_params=None, workspace=None, encoded=None):
    """Adjoint of NGPBRDF.forward: accumulates into d_params (flat fp32 [mlp 9216 | grid], tcnn layout) and returns it."""
    position = position.contiguous().float()
    d_mat = d_mat.contiguous().float()
    n = position.shape[0]
    dev = position.device
    if d_params is None:
        d_params = torch.zeros(N_MLP + tables.t["grid_f16"].numel(), device=dev)
    lib = C.lib()
    wb = lib.iris_field_backward_workspace_bytes(n)
    if workspace is None or workspace.numel() < wb:
        workspace = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    P = tables.c()
    with torch.cuda.device(dev):
        C.check(lib.iris_field_backward(ctypes.byref(P), C.ptr(position), C.ptr(d_mat), n, C.ptr(d_params), C.ptr(encoded), C.ptr(workspace), workspace.numel(),
                                        C.stream_ptr()))
    return d_params


def brdf_shading_forward(mat, diffuse, specular0, specular1):
    """train_brdf_crf.py:193-206: L = kd*diffuse + ks*lerp_specular(specular0, r) + lerp_specular(specular1, r).  mat (n,5)."""
    mat, diffuse = mat.contiguous().float(), diffuse.contiguous().float()
    specular0, specular1 = specular0.contiguous().float(), specular1.contiguous().float()
    n, R = mat.shape[0], specular0.shape[-2]
    if diffuse.shape != (n, 3) or specular0.shape != (n, R, 3) or specular1.shape != (n, R, 3):
        raise ValueError("brdf_shading: expected mat (n,5), diffuse (n,3), specular0/1 (n,R,3)")
    L = torch.empty(n, 3, device=mat.device)
    with torch.cuda.device(mat.device):
        C.check(C.lib().iris_brdf_shading_forward(C.ptr(mat), C.ptr(diffuse), C.ptr(specular0), C.ptr(specular1), R, n, C.ptr(L), C.stream_ptr()))
    return L


def brdf_shading_backward(mat, diffuse, specular0, specular1, dL, d_mat=None):
    """d_mat (n,5) += J^T dL; returns d_mat (zeros are allocated when none is passed)."""
    mat, diffuse, dL = mat.contiguous().float(), diffuse.contiguous().float(), dL.contiguous().float()
    specular0, specular1 = specular0.contiguous().float(), specular1.contiguous().float()
    n, R = mat.shape[0], specular0.shape[-2]
    if d_mat is None:
        d_mat = torch.zeros(n, 5, device=mat.device)
    with torch.cuda.device(mat.device):
        C.check(C.lib().iris_brdf_shading_backward(C.ptr(mat), C.ptr(diffuse), C.ptr(specular0), C.ptr(specular1), R, n, C.ptr(dL), C.ptr(d_mat),
                                                   C.stream_ptr()))
    return d_mat


def single_forward(scene, tables, rays, spp, sampler, want_record=False, workspace=None, want_encoded=None):
    """path_tracing_single forward.  Returns (L (B,3), record or None).  want_encoded (default: same as want_record): also keep the
    samples' encoded field inputs for the field adjoint; they travel with the record as `record.encoded`."""
    rays = rays.contiguous().float()
    B = rays.shape[0]
    dev = rays.device
    L = torch.empty(B, 3, device=dev)
    lib = C.lib()
    wb = lib.iris_single_workspace_bytes(B, int(spp))
    if workspace is None or workspace.numel() < wb:
        workspace = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    record = torch.empty(max(lib.iris_single_record_bytes(B, int(spp)), 16), dtype=torch.uint8, device=dev) if want_record else None
    if want_encoded is None:
        want_encoded = want_record
    enc = torch.empty(max(lib.iris_single_encoded_bytes(B, int(spp)), 16), dtype=torch.uint8, device=dev) if (want_record and want_encoded) else None
    P, S = tables.c(), sampler.c()
    with torch.cuda.device(dev):
        C.check(lib.iris_single_forward(scene.handle, ctypes.byref(P), C.ptr(rays), B, int(spp), ctypes.byref(S), C.ptr(L), C.ptr(record), C.ptr(enc),
                                        C.ptr(workspace), workspace.numel(), C.stream_ptr()))
    if record is not None:
        record.encoded = enc
    return L, record


def single_backward(tables, dL, spp, record, want_radiance=True, d_params=None, workspace=None, encoded=None):
    """Adjoint of path_tracing_single: returns d_radiance (K,3) (accumulated from zero).  encoded: the array the forward kept
    (default: the one travelling with the record as `record.encoded`)."""
    dL = dL.contiguous().float()
    B = dL.shape[0]
    dev = dL.device
    d_rad = torch.zeros(tables.K, 3, device=dev) if want_radiance else None
    lib = C.lib()
    wb = lib.iris_single_workspace_bytes(B, int(spp))
    if d_params is not None and (workspace is None or workspace.numel() < wb):
        workspace = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    P = tables.c()
    with torch.cuda.device(dev):
        enc = encoded if encoded is not None else getattr(record, "encoded", None)
        C.check(lib.iris_single_backward(ctypes.byref(P), C.ptr(dL), B, int(spp), C.ptr(record), C.ptr(enc), C.ptr(d_rad), C.ptr(d_params),
                                         C.ptr(workspace), 0 if workspace is None else workspace.numel(), C.stream_ptr()))
    return d_rad


def _wave_ws(n_lanes, dev, workspace):
    wb = C.lib().iris_wave_workspace_bytes(n_lanes)
    if workspace is None or workspace.numel() < wb:
        workspace = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    return workspace


def path_tracing(scene, tables, rays, spp, indir_depth, sampler, workspace=None):
    """path_tracing (utils/path_tracing.py:214-318), forward only: L (B,3)."""
    rays = rays.contiguous().float()
    B, dev = rays.shape[0], rays.device
    L = torch.empty(B, 3, device=dev)
    ws = _wave_ws(B * int(spp), dev, workspace)
    P, S = tables.c(), sampler.c()
    with torch.cuda.device(dev):
        C.check(C.lib().iris_path_tracing(scene.handle, ctypes.byref(P), C.ptr(rays), B, int(spp), int(indir_depth), ctypes.byref(S), C.ptr(L),
                                          C.ptr(ws), ws.numel(), C.stream_ptr()))
    return L


def path_tracing_det(scene, tables, mode, roughness_level, positions, wis, normals, prim, spp, indir_depth, sampler, workspace=None):
    """path_tracing_det_diff (mode 0) -> L ; path_tracing_det_spec (mode 1) -> (L0, L1)   (utils/path_tracing.py:50-212)."""
    positions, wis, normals = positions.contiguous().float(), wis.contiguous().float(), normals.contiguous().float()
    prim = prim.to(torch.int32).contiguous()
    B, dev = positions.shape[0], positions.device
    L0 = torch.empty(B, 3, device=dev)
    L1 = torch.empty(B, 3, device=dev) if mode == 1 else None
    ws = _wave_ws(B * int(spp), dev, workspace)
    P, S = tables.c(), sampler.c()
    with torch.cuda.device(dev):
        C.check(C.lib().iris_path_tracing_det(scene.handle, ctypes.byref(P), mode, float(roughness_level), C.ptr(positions), C.ptr(wis), C.ptr(normals),
                                              C.ptr(prim), B, int(spp), int(indir_depth), ctypes.byref(S), C.ptr(L0), C.ptr(L1), C.ptr(ws), ws.numel(),
                                              C.stream_ptr()))
    return L0 if mode == 0 else (L0, L1)


def trace_indirect(scene, tables, position, wo, normal, indir_depth, sampler, workspace=None):
    """trace_indirect (utils/path_tracing.py:409-502): L (n,3)."""
    position, wo, normal = position.contiguous().float(), wo.contiguous().float(), normal.contiguous().float()
    n, dev = position.shape[0], position.device
    L = torch.zeros(n, 3, device=dev)
    ws = _wave_ws(n, dev, workspace)
    P, S = tables.c(), sampler.c()
    with torch.cuda.device(dev):
        C.check(C.lib().iris_trace_indirect(scene.handle, ctypes.byref(P), C.ptr(position), C.ptr(wo), C.ptr(normal), n, int(indir_depth), ctypes.byref(S),
                                            C.ptr(L), C.ptr(ws), ws.numel(), C.stream_ptr()))
    return L
