"""Generate tests/golden/*.npz by running the REFERENCE's own estimators (imported from
/root/reference with the two absent third-party engines substituted, see oracle/refharness.py)
on the seeded inputs of tests/golden/cases.py.  Runs only in the build container:

    PYTHONPATH=. python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refharness as RH                    # noqa: E402
from oracle.intersect import OracleScene               # noqa: E402
from tests.golden import cases                         # noqa: E402


def single_with_grads(c, osc):
    em = RH.make_emitter(c["sc"], c["H"], learn=True)
    mat = RH.make_material(c["sc"], c["params"])
    U = torch.as_tensor(c["U"][:, :8])
    L = RH.run_single(osc, em, mat, c["rays"], c["spp"], U)
    (L * torch.as_tensor(c["Gw"])).sum().backward()
    K = c["sc"].n_emitters
    gp = mat.mlp.params.grad.numpy()
    assert np.isfinite(gp).all()
    lv_sum, lv_abs, top_i, top_v = cases.grid_fingerprint(gp[9216:])
    return dict(L=L.detach().numpy(), d_radiance=em.radiance.grad[:K].numpy(),
                d_radiance_rest_absmax=np.float32(em.radiance.grad[K:].abs().max().item()),
                d_mlp=gp[:9216], d_grid_level_sum=lv_sum, d_grid_level_abs=lv_abs, d_grid_top_idx=top_i, d_grid_top_val=top_v)


def shading_golden():
    """tests/golden/shading.npz: the reference's own lerp_specular (utils/ops.py:99-119, imported from /root/reference) inside the
    six restated lines of train_brdf_crf.py:197-206, forward and autograd gradients."""
    import importlib.util
    from oracle import shading as OS
    spec = importlib.util.spec_from_file_location("_ref_utils_ops", "/root/reference/utils/ops.py")
    ref_ops = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(ref_ops)
    x = cases.shading_inputs()
    a, r, m = (x[k].clone().requires_grad_(True) for k in ("albedo", "roughness", "metallic"))
    L = OS.brdf_shading(a, r, m, x["diffuse"], x["specular0"], x["specular1"], lerp=ref_ops.lerp_specular)
    L.backward(x["dL"])
    with torch.no_grad():
        assert torch.equal(ref_ops.lerp_specular(x["specular0"], x["roughness"]), OS.lerp_specular(x["specular0"], x["roughness"]))
    g = dict(L=L.detach().numpy(), d_albedo=a.grad.numpy(), d_roughness=r.grad.numpy(), d_metallic=m.grad.numpy(),
             lerp0=ref_ops.lerp_specular(x["specular0"], x["roughness"]).numpy())
    np.savez_compressed(os.path.join(HERE, "shading.npz"), **g)
    print("shading.npz", {k: v.shape for k, v in g.items()})


def slf_golden():
    """tests/golden/slf.npz: the SLF bake of slf_bake.py:70-145 with the reference's OWN VoxelSLF (model/slf.py, imported from
    /root/reference through the harness) on the seeded points of cases.slf_inputs()."""
    from oracle import slf as OSLF
    ref_slf = RH.load_reference()["slf"].VoxelSLF

    class RefSLF:                                          # adapter: nn.Module buffers -> the attributes oracle.slf.bake reads
        def __init__(self, mask, vmin, vmax):
            self.m = ref_slf(mask, vmin, vmax)

        def scatter_add(self, x, r):
            self.m.scatter_add(x, r)

        inds = property(lambda s: s.m.inds)
        count = property(lambda s: s.m.count)
        radiance = property(lambda s: s.m.radiance, lambda s, v: setattr(s.m, "radiance", v))

    H = 32
    views, rads = cases.slf_inputs()
    out = OSLF.bake(views, rads, H, "synthetic", slf_cls=RefSLF)
    own = OSLF.bake(views, rads, H, "synthetic")
    assert torch.equal(out["weight"]["inds"], own["weight"]["inds"]) and torch.equal(out["weight"]["count"], own["weight"]["count"])
    assert torch.equal(out["weight"]["radiance"], own["weight"]["radiance"]) and torch.equal(out["mask"], own["mask"])
    g = dict(voxel_min=np.float64(out["voxel_min"]), voxel_max=np.float64(out["voxel_max"]), mask=np.packbits(out["mask"].numpy().reshape(-1)),
             inds=out["weight"]["inds"].numpy().astype(np.int32), radiance=out["weight"]["radiance"].numpy(), count=out["weight"]["count"].numpy().astype(np.int32))
    np.savez_compressed(os.path.join(HERE, "slf.npz"), **g)
    print("slf.npz", {k: getattr(v, "shape", v) for k, v in g.items()}, "cells", len(g["count"]))


def small_case(shared_trig):
    """Every estimator on the `small` case.  shared_trig: run the reference with torch.sin / cos / asin / acos replaced by the
    fixed polynomial definitions of oracle/trig.c (the ones the CUDA kernels use) -- "the reference on this libm"."""
    import contextlib
    from oracle import trig as T
    ctx = T.patched_torch() if shared_trig else contextlib.nullcontext()
    with ctx:
        c = cases.build("small")
        osc = OracleScene(c["sc"].vertices, c["sc"].faces)
        r = torch.as_tensor(c["rays"])
        U = torch.as_tensor(c["U"])
        g = single_with_grads(c, osc)
        em = RH.make_emitter(c["sc"], c["H"], learn=False)
        mat = RH.make_material(c["sc"], c["params"])
        pos, nrm, uv, tri, valid = osc.ray_intersect(r[:, 0:3], r[:, 3:6])
        raw = osc.intersect_raw(c["rays"][:, 0:3], c["rays"][:, 3:6], "brute")
        g.update(prim=raw["prim"], t=raw["t"], uv=raw["uv"], p=raw["p"], n=raw["n"])
        g["L_full"] = RH.run_full(osc, em, mat, c["rays"], c["spp"], c["depth"], U).numpy()
        tri2 = tri.clone()
        tri2[5] = -1
        tri2[100] = -1
        Ud = U[:, :2 + 6 * c["depth"]]
        g["det_diff"] = RH.run_det_diff(osc, em, mat, pos, r[:, 3:6], nrm, tri2, c["spp"], c["depth"], Ud).numpy()
        levels = torch.linspace(0.02, 1.0, 6)
        for i in (0, 2, 5):
            a0, a1 = RH.run_det_spec(osc, em, mat, levels[i], pos, r[:, 3:6], nrm, tri2, c["spp"], c["depth"], Ud)
            g["det_spec0_%d" % i], g["det_spec1_%d" % i] = a0.numpy(), a1.numpy()
        g["bake_diff"] = RH.run_bake(osc, em, pos, nrm, -r[:, 3:6], c["spp"], U[:, :2]).numpy()
        for i in range(6):
            a0, a1 = RH.run_bake(osc, em, pos, nrm, -r[:, 3:6], c["spp"], U[:, :2], level=levels[i])
            g["bake_spec0_%d" % i], g["bake_spec1_%d" % i] = a0.numpy(), a1.numpy()
        # the sampled directions themselves (BaseBRDF.sample_diffuse / sample_specular / sample_brdf on the pixel-centre hits)
        n = len(pos)
        brdf = RH.load_reference()["brdf"].BaseBRDF()
        u3 = U[:n, 5:8].contiguous()
        wo = -r[:, 3:6]
        g["dir_diffuse"] = brdf.sample_diffuse(u3[:, 1:3], nrm)[0].numpy()
        g["dir_specular"] = brdf.sample_specular(u3[:, 1:3], wo, nrm, levels[2])[0].numpy()
        with torch.no_grad():
            m = mat(pos)
            wi, pdf, w = brdf.sample_brdf(u3[:, 0], u3[:, 1:3], wo, nrm, m)
        g["dir_brdf"], g["dir_brdf_pdf"], g["dir_brdf_w"] = wi.numpy(), pdf.numpy(), w.numpy()
    return g


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "shading":
        return shading_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "slf":
        return slf_golden()
    # ---------------- small: every estimator, with the reference on torch's own libm and on the shared polynomial definitions
    for tag, st in (("small", False), ("small_st", True)):
        if len(sys.argv) > 1 and sys.argv[1] == "c1":
            break
        g = small_case(st)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **g)
        print(tag + ".npz", {k: v.shape for k, v in g.items()})
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        return

    # ---------------- C1: Cornell 64x64 spp 16, path_tracing_single fwd + bwd (BASELINE.json configs[0])
    from oracle import trig as T
    c = cases.build("c1")
    osc = OracleScene(c["sc"].vertices, c["sc"].faces)
    g = single_with_grads(c, osc)
    np.savez_compressed(os.path.join(HERE, "c1.npz"), **g)
    print("c1.npz", {k: v.shape for k, v in g.items()})
    with T.patched_torch():
        g = single_with_grads(c, osc)
    np.savez_compressed(os.path.join(HERE, "c1_st.npz"), **g)
    print("c1_st.npz", {k: v.shape for k, v in g.items()})
    if len(sys.argv) > 1 and sys.argv[1] == "c1":
        return
    shading_golden()
    slf_golden()


if __name__ == "__main__":
    main()
