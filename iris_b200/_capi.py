"""ctypes binding of libiris_b200.so (include/iris_b200.h).

torch is used only for device memory and streams: every call passes raw device pointers and the current CUDA stream
through the C ABI.  There is no CPU fallback: importing works anywhere (so CPU-only tests can check the exported symbols),
but every compute entry point raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IRIS_B200_LIB") or os.path.join(HERE, "_lib", "libiris_b200.so")   # override: A/B builds of the same ABI

c_i64, c_i32, c_f32, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p


class IrisSceneStats(ctypes.Structure):
    _fields_ = [("n_tris", c_i64), ("n_nodes", c_i64), ("node_bytes", c_i64), ("tri_bytes", c_i64), ("build_ms", c_f32),
                ("sah_cost", c_f32), ("max_depth", c_i32), ("bounds_lo", c_f32 * 3), ("bounds_hi", c_f32 * 3)]


class IrisShadeParams(ctypes.Structure):
    _fields_ = [("emitter_of_face", c_vp), ("face_of_emitter", c_vp), ("emitter_vertices", c_vp), ("emitter_area", c_vp),
                ("emitter_pdf", c_vp), ("emitter_cdf", c_vp), ("radiance", c_vp), ("n_emitters", c_i32), ("n_faces", c_i32),
                ("slf_inds", c_vp), ("slf_radiance", c_vp), ("slf_H", c_i32), ("slf_vmin", c_f32), ("slf_range", c_f32),
                ("grid_f16", c_vp), ("mlp_f16", c_vp), ("field_vmin", c_f32), ("field_range", c_f32)]


class IrisSampler(ctypes.Structure):
    _fields_ = [("U", c_vp), ("stride", c_i32), ("seed", ctypes.c_uint64), ("lane_offset", ctypes.c_uint64)]


# name -> (restype, argtypes); every symbol include/iris_b200.h declares
PROTOTYPES = {
    "iris_last_error": (ctypes.c_char_p, []),
    "iris_version": (ctypes.c_char_p, []),
    "iris_scene_create": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "iris_scene_destroy": (None, [c_vp]),
    "iris_scene_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(IrisSceneStats)]),
    "iris_intersect": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "iris_sampler_fill": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, c_i64, c_i32, c_vp, c_vp]),
    "iris_bake": (ctypes.c_int, [c_vp, ctypes.POINTER(IrisShadeParams), ctypes.c_int, c_f32, c_vp, c_vp, c_vp, c_i64, c_i32,
                                 ctypes.POINTER(IrisSampler), c_vp, c_vp, c_vp]),
    "iris_field_levels": (c_i64, [c_vp, c_vp, c_vp, c_vp]),
    "iris_field_forward": (ctypes.c_int, [ctypes.POINTER(IrisShadeParams), c_vp, c_i64, c_vp, c_vp, c_vp]),
    "iris_field_backward_workspace_bytes": (c_i64, [c_i64]),
    "iris_field_backward": (ctypes.c_int, [ctypes.POINTER(IrisShadeParams), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "iris_single_workspace_bytes": (c_i64, [c_i64, c_i32]),
    "iris_single_record_bytes": (c_i64, [c_i64, c_i32]),
    "iris_single_encoded_bytes": (c_i64, [c_i64, c_i32]),
    "iris_single_forward": (ctypes.c_int, [c_vp, ctypes.POINTER(IrisShadeParams), c_vp, c_i64, c_i32, ctypes.POINTER(IrisSampler),
                                           c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "iris_single_backward": (ctypes.c_int, [ctypes.POINTER(IrisShadeParams), c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "iris_wave_workspace_bytes": (c_i64, [c_i64]),
    "iris_path_tracing": (ctypes.c_int, [c_vp, ctypes.POINTER(IrisShadeParams), c_vp, c_i64, c_i32, c_i32, ctypes.POINTER(IrisSampler), c_vp, c_vp, c_i64, c_vp]),
    "iris_path_tracing_det": (ctypes.c_int, [c_vp, ctypes.POINTER(IrisShadeParams), ctypes.c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32,
                                             ctypes.POINTER(IrisSampler), c_vp, c_vp, c_vp, c_i64, c_vp]),
    "iris_trace_indirect": (ctypes.c_int, [c_vp, ctypes.POINTER(IrisShadeParams), c_vp, c_vp, c_vp, c_i64, c_i32, ctypes.POINTER(IrisSampler), c_vp, c_vp, c_i64, c_vp]),
    "iris_slf_bounds": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "iris_slf_mark": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_float, ctypes.c_float, c_i32, c_vp, c_vp]),
    "iris_slf_index_workspace_bytes": (c_i64, [c_i32]),
    "iris_slf_index": (ctypes.c_int, [c_vp, c_i32, c_vp, ctypes.POINTER(c_i64), c_vp, c_i64, c_vp]),
    "iris_slf_accumulate": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_float, ctypes.c_float, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "iris_slf_finalize": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "iris_tri_accumulate": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "iris_emitter_classify": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_float, c_vp, c_vp]),
    "iris_emitter_geometry": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "iris_brdf_shading_forward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_vp, c_vp]),
    "iris_brdf_shading_backward": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_vp, c_vp, c_vp]),
    "iris_denoise_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "iris_denoise_atrous": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_i64, c_vp]),
    "iris_crf_forward": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i64, c_vp, c_vp]),
    "iris_crf_backward": (ctypes.c_int, [c_vp, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "iris_bsdf_sample": (ctypes.c_int, [ctypes.c_int, c_vp, c_i32, c_vp, c_vp, c_vp, c_f32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "iris_abi_info": (c_i64, [ctypes.c_int]),
    "iris_launch_count": (c_i64, []),
    "iris_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "iris_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "iris_profile_name": (ctypes.c_char_p, [ctypes.c_int]),
    "iris_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_double), ctypes.c_int]),
}

_LIB = None
ABI_VERSION = 3          # include/iris_b200.h: IRIS_ABI_VERSION


def lib():
    """Load libiris_b200.so.  With nvcc on PATH the (mtime-checked) in-tree build runs first, so an edited source never meets a
    stale binary; without nvcc a missing library is an error.  The library's ABI version and struct sizes are checked against the
    ctypes mirrors above: a drifted layout raises instead of corrupting device pointers."""
    global _LIB
    if _LIB is None:
        import shutil
        if "IRIS_B200_LIB" not in os.environ and (shutil.which("nvcc") or not os.path.exists(LIB_PATH)):
            from . import build as _build
            _build.build()
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        got = [L.iris_abi_info(k) for k in range(4)]
        want = [ABI_VERSION, ctypes.sizeof(IrisShadeParams), ctypes.sizeof(IrisSampler), ctypes.sizeof(IrisSceneStats)]
        if got != want:
            raise RuntimeError("iris_b200: %s does not match this binding (abi/struct sizes %r, expected %r) -- rebuild with python iris_b200/build.py --force"
                               % (LIB_PATH, got, want))
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise RuntimeError("iris_b200: %s (status %d)" % (lib().iris_last_error().decode(), rc))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("iris_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())
