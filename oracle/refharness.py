"""oracle/refharness.py -- TEST INFRASTRUCTURE, not product code.  Usable ONLY where
/root/reference exists (this build container); nothing under tests/ -m gpu, smoke() or bench.py
imports it.

Runs the REFERENCE's own estimators (utils/path_tracing.py, model/brdf.py, model/emitter.py,
model/slf.py -- imported unmodified from /root/reference) on CPU, following the recipe verified
in SURVEY.md Appendix C:

  * `mitsuba` and `tinycudann` (absent third-party engines) are stubbed in sys.modules;
  * utils.path_tracing.ray_intersect is replaced by the CPU closest-hit of oracle/intersect.c;
  * tcnn.NetworkWithInputEncoding is replaced by oracle/field.CpuHashGridMLP;
  * torch.rand is served from an explicit full-lane sample buffer U: the reference draws on
    COMPACTED lane sets (utils/path_tracing.py:347-352,492-501), so the harness tracks the
    compaction (boolean masks keep lane order) by observing the valid_next masks returned by
    emitter_net.eval_emitter, and hands each draw the rows of U that belong to the surviving lanes.

The outputs are the golden vectors committed under tests/golden/ (tests/golden/make_golden.py).
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
_MODS = {}


def available():
    return os.path.isdir(os.path.join(REF, "utils"))


def load_reference():
    """Import the reference modules with stubs; returns dict(pt=, brdf=, emitter=, slf=)."""
    if _MODS:
        return _MODS
    from . import field as ofield
    sys.dont_write_bytecode = True                       # the reference tree is read-only
    mi = types.ModuleType("mitsuba")
    mi.set_variant = lambda *a, **k: None
    mi.math = types.SimpleNamespace(RayEpsilon=1500.0 * 2.0 ** -24)
    mi.BSDF = object
    tc = types.ModuleType("tinycudann")
    tc.NetworkWithInputEncoding = ofield.CpuHashGridMLP
    saved = {k: sys.modules.get(k) for k in ("mitsuba", "tinycudann", "utils", "model", "const")}
    sys.modules["mitsuba"] = mi
    sys.modules["tinycudann"] = tc
    for k in ("utils", "model", "const"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    try:
        os.chdir(REF)
        import utils.path_tracing as pt
        import model.brdf as brdf
        import model.emitter as emitter
        import model.slf as slf
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    pt.ray_intersect = lambda scene, xs, ds: scene.ray_intersect(xs, ds)
    _MODS.update(pt=pt, brdf=brdf, emitter=emitter, slf=slf, mitsuba=mi)
    # keep the reference's top-level names out of the way of this repo's own packages
    for k in ("utils", "model", "const"):
        m = sys.modules.pop(k, None)
        if m is not None:
            _MODS["_" + k] = m
        if saved[k] is not None:
            sys.modules[k] = saved[k]
    for k in list(sys.modules):
        if k.startswith(("utils.", "model.")) and getattr(sys.modules[k], "__file__", "").startswith(REF):
            _MODS["_" + k] = sys.modules.pop(k)
    return _MODS


def load_reference_crf():
    """Import the reference's crf/model_crf.py (EmorCRF on the real EMoR tables crf/emor.txt) with its absent third-party imports
    stubbed: torch_interpolations -> oracle/crf.RegularGridInterpolator (the package's published 1-D algorithm), matplotlib -> empty."""
    if "crf" in _MODS:
        return _MODS["crf"]
    from . import crf as ocrf
    sys.dont_write_bytecode = True
    ti = types.ModuleType("torch_interpolations")
    ti.RegularGridInterpolator = ocrf.RegularGridInterpolator
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    saved = {k: sys.modules.get(k) for k in ("torch_interpolations", "matplotlib", "matplotlib.pyplot", "crf", "const")}
    sys.modules.update({"torch_interpolations": ti, "matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})
    for k in ("crf", "const"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    try:
        os.chdir(REF)                                     # crf/emor.py:14-17 resolves emor.txt from the working directory
        import crf.model_crf as mc
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    for k in list(sys.modules):
        if k == "crf" or k.startswith("crf.") or k == "const":
            if getattr(sys.modules[k], "__file__", "").startswith(REF):
                _MODS["_" + k] = sys.modules.pop(k)
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
        elif k in ("torch_interpolations", "matplotlib", "matplotlib.pyplot"):
            sys.modules.pop(k, None)
    _MODS["crf"] = mc
    return mc


def run_extract_emitter_script(scene, views, threshold):
    """Execute the reference's extract_emitter_ldr.py (its `main()`, mode 'export', lines 72-115) UNMODIFIED on CPU and return the
    emitter.pth dict it writes.  The script's absent imports are stubbed for the duration of the run: mitsuba (load_dict -> the CPU
    closest-hit scene), utils.dataset (a list of {'rays','rgbs'} batches = `views`), utils.path_tracing.ray_intersect (oracle/intersect.c),
    trimesh.load_mesh (the scene's own arrays), torch_scatter.scatter(src, index, 0, out, reduce='sum') -> out.index_add (the
    package's documented semantics), and torch.device(0) -> cpu."""
    import runpy
    from .intersect import OracleScene
    osc = OracleScene(scene.vertices, scene.faces)
    sys.dont_write_bytecode = True

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    class _Dataset(list):
        img_hw = (0, 0)

    def make_dataset(*a, **k):
        return _Dataset([{"rays": torch.as_tensor(r), "rgbs": torch.as_tensor(c)} for r, c in views])

    def scatter(src, index, dim, out, reduce="sum"):
        assert dim == 0 and reduce == "sum"
        return out.index_add(0, index, src)

    stubs = {
        "mitsuba": mod("mitsuba", set_variant=lambda *a, **k: None, load_dict=lambda d: osc),
        "utils": mod("utils", __path__=[]),
        "utils.dataset": mod("utils.dataset", __path__=[], InvRealDatasetLDR=make_dataset, InvSyntheticDatasetLDR=make_dataset),
        "utils.dataset.scannetpp": mod("utils.dataset.scannetpp", __path__=[]),
        "utils.dataset.scannetpp.dataset": mod("utils.dataset.scannetpp.dataset", InvScannetpp=make_dataset),
        "utils.path_tracing": mod("utils.path_tracing", ray_intersect=lambda sc, xs, ds: sc.ray_intersect(xs, ds)),
        "trimesh": mod("trimesh", load_mesh=lambda p: types.SimpleNamespace(vertices=np.asarray(scene.vertices), faces=np.asarray(scene.faces).astype(np.int64))),
        "torch_scatter": mod("torch_scatter", scatter=scatter),
    }
    saved = {k: sys.modules.get(k) for k in list(stubs) + ["const"]}
    argv, real_device = sys.argv, torch.device
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "scene.obj"), "w").write("# geometry comes from the trimesh stub\n")
        try:
            sys.modules.update(stubs)
            sys.modules.pop("const", None)
            sys.path.insert(0, REF)
            sys.argv = ["extract_emitter_ldr.py", "--scene", td, "--output", td, "--dataset", "synthetic", "--threshold", repr(float(threshold))]
            torch.device = lambda *a, **k: real_device("cpu")
            runpy.run_path(os.path.join(REF, "extract_emitter_ldr.py"), run_name="__main__")
        finally:
            torch.device = real_device
            sys.argv = argv
            sys.path.remove(REF)
            _MODS["_const_extract"] = sys.modules.pop("const", None)
            for k, v in saved.items():
                if v is not None:
                    sys.modules[k] = v
                else:
                    sys.modules.pop(k, None)
        return torch.load(os.path.join(td, "emitter.pth"))


def make_emitter(scene, H, learn=True):
    """Write the scene's emitter.pth / vslf.npz to a temp dir in the reference's on-disk format
    (extract_emitter_ldr.py:109-115, slf_bake.py:140-145) and load them with the reference classes."""
    m = load_reference()
    ed = scene.emitter_dict()
    sd = scene.slf_dict(H)
    with tempfile.TemporaryDirectory() as td:
        ep, sp = os.path.join(td, "emitter.pth"), os.path.join(td, "vslf.npz")
        torch.save({k: torch.as_tensor(v) for k, v in ed.items()}, ep)
        torch.save({"mask": torch.as_tensor(sd["mask"]), "voxel_min": sd["voxel_min"], "voxel_max": sd["voxel_max"],
                    "weight": {k: torch.as_tensor(v) for k, v in sd["weight"].items()}}, sp)
        cls = m["emitter"].SLFEmitterLearn if learn else m["emitter"].SLFEmitter
        em = cls(ep, sp)
    return em


def make_material(scene, params=None):
    m = load_reference()
    vmin, vmax = scene.voxel_bounds()
    mat = m["brdf"].NGPBRDF(vmin, vmax)
    if params is not None:
        with torch.no_grad():
            mat.mlp.params.copy_(params)
    return mat


class _Injector:
    """Serves torch.rand from U and follows the reference's lane compaction."""

    def __init__(self, U, n_lanes, first_eval_compacts):
        self.U = U
        self.lanes = torch.arange(n_lanes)
        self.col = 0
        self.first_eval_compacts = first_eval_compacts
        self.n_eval = 0

    def rand(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        if len(shape) == 4:                                # torch.rand(2,B,spp,1): camera jitter
            assert shape[0] == 2 and shape[3] == 1 and shape[1] * shape[2] == len(self.lanes)
            out = self.U[:, 0:2].t().reshape(2, shape[1], shape[2], 1).clone()
            self.col = 2
            return out
        n = shape[0]
        assert n == len(self.lanes), (shape, len(self.lanes))
        w = 1 if len(shape) == 1 else shape[1]
        out = self.U[self.lanes, self.col:self.col + w].clone()
        self.col += w
        return out.reshape(shape)

    def wrap_eval(self, fn):
        def eval_emitter(position, light_dir, triangle_idx, roughness=None, *a, **k):
            out = fn(position, light_dir, triangle_idx, roughness, *a, **k)
            first = self.n_eval == 0
            self.n_eval += 1
            if roughness is not None or (first and self.first_eval_compacts):
                self.lanes = self.lanes[out[2]]
            return out
        return eval_emitter


def _nan_guard(v):
    v2 = v.clone()
    if v2.requires_grad:
        v2.register_hook(lambda g: torch.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0))
    return v2


@contextlib.contextmanager
def injected(emitter, material, U, n_lanes, first_eval_compacts, lanes0=None, guard_nan_grad=False):
    """Patch torch.rand + emitter.eval_emitter for the duration of one reference estimator call.
    guard_nan_grad: zero the NaN gradient the reference produces on pdf==0 lanes of sample_brdf
    (model/brdf.py:208) by hooking a clone of `mat` -- the reference code itself is not modified."""
    inj = _Injector(U, n_lanes, first_eval_compacts)
    if lanes0 is not None:
        inj.lanes = lanes0
    orig_rand = torch.rand
    orig_eval = emitter.eval_emitter
    orig_sample = getattr(material, "sample_brdf", None)
    torch.rand = inj.rand
    emitter.eval_emitter = inj.wrap_eval(orig_eval)
    if guard_nan_grad and orig_sample is not None:
        material.sample_brdf = lambda s1, s2, wo, n, mat: orig_sample(s1, s2, wo, n, {k: _nan_guard(v) for k, v in mat.items()})
    try:
        yield inj
    finally:
        torch.rand = orig_rand
        del emitter.eval_emitter
        if guard_nan_grad and orig_sample is not None:
            del material.sample_brdf


# ----------------------------------------------------------------------------- estimator runs
def run_single(oscene, emitter, material, rays, spp, U, guard_nan_grad=True):
    m = load_reference()
    r = torch.as_tensor(rays)
    with injected(emitter, material, U, len(r) * spp, True, guard_nan_grad=guard_nan_grad):
        return m["pt"].path_tracing_single(oscene, emitter, material, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp)


def run_full(oscene, emitter, material, rays, spp, depth, U):
    m = load_reference()
    r = torch.as_tensor(rays)
    with torch.no_grad(), injected(emitter, material, U, len(r) * spp, True):
        return m["pt"].path_tracing(oscene, emitter, material, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp, depth)


def _det_lanes(tri, spp):
    ok = (tri != -1)
    return torch.arange(len(tri) * spp)[ok.repeat_interleave(spp, 0)]


def run_det_diff(oscene, emitter, material, positions, wis, normals, tri, spp, depth, U):
    m = load_reference()
    with torch.no_grad(), injected(emitter, material, U, len(tri) * spp, False, lanes0=_det_lanes(tri, spp)):
        return m["pt"].path_tracing_det_diff(oscene, emitter, material, positions, wis, normals, None, tri, spp, depth)


def run_det_spec(oscene, emitter, material, level, positions, wis, normals, tri, spp, depth, U):
    m = load_reference()
    with torch.no_grad(), injected(emitter, material, U, len(tri) * spp, False, lanes0=_det_lanes(tri, spp)):
        return m["pt"].path_tracing_det_spec(oscene, emitter, material, level, positions, wis, normals, None, tri, spp, depth)


def run_bake(oscene, emitter, position, normal, wo, spp, U, level=None):
    """The inner loop of bake_shading.py (108-123 diffuse, 168-188 specular) executed with the reference's
    own BaseBRDF / SLFEmitter methods, one chunk (the loop lives in the script's __main__)."""
    m = load_reference()
    mi = m["mitsuba"]
    pt = m["pt"]
    material_net = m["brdf"].BaseBRDF()
    n = len(position)
    with torch.no_grad(), injected(emitter, material_net, U, n * spp, False):
        if level is None:
            wi, _, _ = material_net.sample_diffuse(torch.rand(n * spp, 2), normal.repeat_interleave(spp, 0))
        else:
            wi, _, g0, g1 = material_net.sample_specular(torch.rand(n * spp, 2), wo.repeat_interleave(spp, 0),
                                                         normal.repeat_interleave(spp, 0), level)
        p_next, _, _, tri_next, _ = pt.ray_intersect(oscene, position.repeat_interleave(spp, 0) + mi.math.RayEpsilon * wi,
                                                     wi.reshape(-1, 3))
        roughness_one = torch.ones_like(tri_next)[:, None]
        Le, _, _ = emitter.eval_emitter(p_next, wi, tri_next, roughness_one, trace_roughness=0.0)
        if level is None:
            return Le.reshape(n, spp, 3).mean(1)
        return (Le * g0).reshape(n, spp, 3).mean(1), (Le * g1).reshape(n, spp, 3).mean(1)
