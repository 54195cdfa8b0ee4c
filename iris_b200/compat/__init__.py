"""Compatibility shims that let the reference's stage drivers (train_emitter.py, bake_shading.py, render.py, ...) run on this
path unchanged: stand-ins for the two absent third-party engines and a patcher for the reference's own modules.

    import iris_b200.compat as compat
    compat.install("/path/to/iris")      # BEFORE importing any reference module
    import train_emitter                  # `import mitsuba`, `import tinycudann`, `from utils.path_tracing import ...` now resolve here

What `install` does:
  * sys.modules['mitsuba']     = iris_b200.compat.mitsuba      (set_variant, load_dict -> BVH scene, math.RayEpsilon, OptixDenoiser, TensorXf)
  * sys.modules['tinycudann']  = iris_b200.compat.tinycudann   (NetworkWithInputEncoding: parameter container in tcnn's layout)
  * with a reference root: imports the reference's utils.path_tracing / model.brdf and replaces the six estimator functions and
    NGPBRDF by the CUDA-backed ones (the reference's SLFEmitter / VoxelSLF classes are used as they are: the kernels read their buffers).
"""
from __future__ import annotations

import os
import sys


def install(reference_root=None):
    from . import mitsuba as mi_shim
    from . import tinycudann as tcnn_shim
    sys.modules["mitsuba"] = mi_shim
    sys.modules["tinycudann"] = tcnn_shim
    if reference_root is None:
        return
    root = os.path.abspath(reference_root)
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    pt = importlib.import_module("utils.path_tracing")
    brdf = importlib.import_module("model.brdf")
    from ..model.brdf import NGPBRDF
    from ..utils import path_tracing as ours
    for name in ("ray_intersect", "path_tracing", "path_tracing_single", "path_tracing_det_diff", "path_tracing_det_spec", "trace_indirect"):
        setattr(pt, name, getattr(ours, name))
    brdf.NGPBRDF = NGPBRDF
