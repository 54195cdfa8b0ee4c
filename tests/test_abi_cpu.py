"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/iris_b200.h declares
(no compute calls without a GPU), and its hash-grid level table is the oracle's."""
import os
import re

import numpy as np

from iris_b200 import _capi as C
from oracle import field as OF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "iris_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(iris_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(C.PROTOTYPES), declared ^ set(C.PROTOTYPES)
    L = C.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.iris_version()


def test_level_table_matches_oracle():
    from iris_b200.core import field_levels
    sc, res, size, off, n = field_levels()
    assert n == OF.N_ENTRIES
    for l, (scale, r, s, o) in enumerate(OF.LEVELS):
        assert np.float32(scale) == sc[l] and r == res[l] and s == size[l] and o == off[l], l


def test_no_compute_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from iris_b200.core import Scene
    with pytest.raises(RuntimeError):
        Scene(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.int32))


def test_product_does_not_import_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "iris_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


_COMPAT_SCRIPT = r'''
import os, sys, tempfile
import numpy as np, torch
sys.dont_write_bytecode = True
sys.path.insert(0, %(root)r)
import iris_b200.compat as compat
from iris_b200 import ops, scenes, core
ref = "/root/reference"
os.chdir(ref)                                      # the reference resolves crf/emor.txt and `const` from its own root
compat.install(ref)                                # BEFORE any reference module is imported
import utils.path_tracing as pt                    # the REFERENCE's module objects ...
import model.brdf as rbrdf
import model.emitter as remitter
from iris_b200.utils import path_tracing as ours
from iris_b200.model.brdf import NGPBRDF
assert pt.__file__.startswith(ref) and remitter.__file__.startswith(ref)
for name in ("ray_intersect", "path_tracing", "path_tracing_single", "path_tracing_det_diff", "path_tracing_det_spec", "trace_indirect"):
    assert getattr(pt, name) is getattr(ours, name), name          # ... now carry the CUDA-backed estimators
assert rbrdf.NGPBRDF is NGPBRDF
import mitsuba, tinycudann
assert mitsuba.math.RayEpsilon == 1500.0 * 2.0 ** -24 and hasattr(tinycudann, "NetworkWithInputEncoding")
# the reference's OWN emitter class, loaded from files in its on-disk format, feeds the kernels' tables unchanged (duck typing)
sc = scenes.cornell()
ed, sd = sc.emitter_dict(), sc.slf_dict(32)
with tempfile.TemporaryDirectory() as td:
    torch.save({k: torch.as_tensor(v) for k, v in ed.items()}, os.path.join(td, "emitter.pth"))
    torch.save({"mask": torch.as_tensor(sd["mask"]), "voxel_min": sd["voxel_min"], "voxel_max": sd["voxel_max"],
                "weight": {k: torch.as_tensor(v) for k, v in sd["weight"].items()}}, os.path.join(td, "vslf.npz"))
    em = remitter.SLFEmitterLearn(os.path.join(td, "emitter.pth"), os.path.join(td, "vslf.npz"))
mat = rbrdf.NGPBRDF(*sc.voxel_bounds())
T = ops.tables_for(em, mat, "cpu")
T2 = core.ShadingTables.from_dicts("cpu", ed, sd, mat.mlp.params, sc.voxel_bounds())
for k in T2.t:
    assert torch.equal(T.t[k], T2.t[k]), k
assert (T.K, T.F, T.H, T.slf_vmin, T.slf_range, T.field_vmin, T.field_range) == (T2.K, T2.F, T2.H, T2.slf_vmin, T2.slf_range, T2.field_vmin, T2.field_range)
# the patched estimator is the product path: without a GPU it must refuse loudly (no CPU fallback), with one it runs
rays = torch.as_tensor(sc.camera_rays(4, 4))
class FakeScene: iris_scene = None
try:
    if torch.cuda.is_available():
        dev = torch.device("cuda", 0)
        scene = core.Scene(sc.vertices, sc.faces, 0)
        em, mat, r = em.to(dev), mat.to(dev), rays.to(dev)
        L = pt.path_tracing_single(scene, em, mat, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], 4)
        L.sum().backward()
        assert L.shape == (16, 3) and torch.isfinite(L).all() and em.radiance.grad is not None and mat.mlp.params.grad is not None
        print("COMPAT-OK gpu")
    else:
        try:
            core.Scene(sc.vertices, sc.faces, 0)
        except RuntimeError as e:
            assert "CUDA" in str(e)
            print("COMPAT-OK cpu")
except Exception:
    raise
'''


def test_compat_install_patches_the_reference_tree():
    """compat.install('/root/reference') -- the zero-edit route: the reference's own utils.path_tracing / model.brdf modules get the
    CUDA-backed estimators and NGPBRDF, `import mitsuba` / `import tinycudann` resolve to the shims, and the reference's own
    SLFEmitterLearn (loaded from emitter.pth / vslf.npz) produces exactly the device tables the dict route produces.  Runs in a
    subprocess (the reference's top-level module names must not leak into this session); skipped where the tree is absent."""
    import subprocess
    import sys
    import pytest
    if not os.path.isdir("/root/reference/utils"):
        pytest.skip("reference tree not present")
    out = subprocess.run([sys.executable, "-c", _COMPAT_SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "COMPAT-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_torch_library_ops_have_schemas_and_fake_impls():
    """The entry points are torch.library custom ops (iris_b200/library.py): schema'd, and traceable with fake tensors on a box
    without a GPU (output shapes / dtypes of every op; the registered adjoints themselves run in the GPU suite)."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    import iris_b200.ops  # noqa: F401  (registers the ops)
    ns = torch.ops.iris_b200
    for name in ("intersect", "single_forward", "single_backward", "field_forward", "field_backward", "bake", "path_tracing"):
        assert hasattr(ns, name), name
    assert "Tensor? params" in str(ns.single_forward.default._schema) and "Tensor? U" in str(ns.single_forward.default._schema)
    with FakeTensorMode():
        rays = torch.empty(100, 12, device="cuda")
        rad = torch.empty(50, 3, device="cuda", requires_grad=True)
        par = torch.empty(9216 + 27954112, device="cuda", requires_grad=True)
        L, rec, enc = ns.single_forward(rad, par, rays, 1, 2, 8, 0, 0, None, 2)
        assert L.shape == (100, 3) and rec.numel() == 96 * 800 and enc.numel() == 128 * 800 and L.requires_grad and L.device.type == "cuda"
        L, rec, enc = ns.single_forward(rad, None, rays, 1, 2, 8, 0, 0, None, 0)
        assert rec.numel() == 0 and enc.numel() == 0
        a, b = ns.single_backward(torch.empty(100, 3, device="cuda"), rec, enc, 2, 8, 50, 0)
        assert a.shape == (50, 3) and b.numel() == 0
        m, e = ns.field_forward(par, torch.empty(77, 3, device="cuda"), 2, True)
        assert m.shape == (77, 5) and e.shape == (77, 64) and e.dtype == torch.float16 and m.requires_grad
        assert ns.field_backward(torch.empty(77, 5, device="cuda"), torch.empty(77, 3, device="cuda"), e, 2, par.numel()).shape == par.shape
        t, prim, uv, p, n = ns.intersect(1, rays[:, :3], rays[:, 3:6])
        assert prim.dtype == torch.int32 and uv.shape == (100, 2)
        o0, o1 = ns.bake(rays[:, :3], rays[:, 3:6], rays[:, 6:9], 1, 2, 1, 0.5, 16, 0, None)
        assert o0.shape == (100, 3) and o1.shape == (100, 3)
        assert ns.path_tracing(rays, 1, 2, 16, 5, 0, None).shape == (100, 3)
