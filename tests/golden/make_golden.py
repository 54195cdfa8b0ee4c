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
    # the refine pass, slf_refine.py:85-108: the reference's VoxelSLF rebuilt from the saved mask / bounds, radiance re-accumulated from
    # LDR colours through the reference's EmorCRF.inverse (real EMoR tables, the weights of the CRF case), mean pooling
    mc = RH.load_reference_crf()
    crf = mc.EmorCRF(dim=11)
    with torch.no_grad():
        crf.weight.copy_(cases.crf_inputs()["weight"])
    ldr, exposure = cases.slf_refine_inputs(views)
    ref2 = ref_slf(out["mask"], out["voxel_min"], out["voxel_max"])
    with torch.no_grad():
        for (pos, valid), c, e in zip(views, ldr, exposure):
            rad = crf.inverse(c, e)
            if not valid.any():
                continue
            ref2.scatter_add(pos[valid], rad[valid])
        ref2.radiance = ref2.radiance / ref2.count[..., None].float().clamp_min(1)
    assert torch.equal(ref2.inds, out["weight"]["inds"])
    g["refine_radiance"], g["refine_count"] = ref2.radiance.numpy(), ref2.count.numpy().astype(np.int32)
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


def crf_golden():
    """tests/golden/crf.npz: the reference's own EmorCRF (crf/model_crf.py, imported from /root/reference with the absent
    torch_interpolations stubbed by its published 1-D algorithm) on the REAL EMoR tables (crf/emor.txt): forward + autograd
    gradients, the inverse table, inverse(), and the weight fit.  f0 and the 11 basis curves are stored with the outputs (12 x 1024
    floats of the public EMoR database) because /root/reference does not exist where the GPU tests run."""
    mc = RH.load_reference_crf()
    x = cases.crf_inputs()
    m = mc.EmorCRF(dim=11)
    with torch.no_grad():
        m.weight.copy_(x["weight"])
    hdr = x["hdr"].clone().requires_grad_(True)
    ldr = m(hdr, x["exposure"])
    (ldr * x["d_ldr"]).sum().backward()
    g = dict(f0=m.f0.numpy()[0], basis=m.basis.numpy(), ldr=ldr.detach().numpy(), d_hdr=hdr.grad.numpy(), d_weight=m.weight.grad.numpy(),
             crf=m.get_crf().detach().numpy(), inv_crf=m.get_inv_crf().detach().numpy(),
             hdr_inv=m.inverse(x["ldr"], x["exposure"]).detach().numpy(),
             ldr_scalar_exposure=m(x["hdr"], torch.tensor(0.7)).detach().numpy())
    # a monotone weight set (gamma-like camera): forward -> inverse is a round trip; and the weight fit of that response
    m2 = mc.EmorCRF(dim=11)
    target = torch.linspace(0, 1, 1024).pow(1 / 2.2)[None].expand(3, 1024).numpy() * np.array([[1.0], [0.97], [1.02]], np.float32)
    g["fit_target"] = target.astype(np.float32)
    g["fit_weight"] = m2.cal_weight_fitting_crf(target).astype(np.float32)
    m2.initialize_weight(target)
    g["fit_crf"] = m2.get_crf().detach().numpy()
    g["fit_inv_crf"] = m2.get_inv_crf().detach().numpy()
    g["fit_roundtrip"] = m2.inverse(m2(x["hdr"], x["exposure"]).detach(), x["exposure"]).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "crf.npz"), **g)
    print("crf.npz", {k: v.shape for k, v in g.items()})


def emitter_extract_golden():
    """tests/golden/emitter_extract.npz: the reference's own extract_emitter_ldr.py (mode 'export') executed unmodified through the
    harness on the seeded views of cases.emitter_extract_inputs()."""
    sc, views = cases.emitter_extract_inputs()
    out = RH.run_extract_emitter_script(sc, views, 0.99)
    g = {k: np.asarray(v.numpy()) for k, v in out.items()}
    g["is_emitter"] = np.packbits(g["is_emitter"])
    g["emitter_radiance_shape"] = np.array(out["emitter_radiance"].shape)
    g["emitter_radiance_absmax"] = np.float32(out["emitter_radiance"].abs().max().item())
    del g["emitter_radiance"]
    np.savez_compressed(os.path.join(HERE, "emitter_extract.npz"), **g)
    print("emitter_extract.npz", {k: v.shape for k, v in g.items()}, "K =", int(out["is_emitter"].sum()))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "emitter_extract":
        return emitter_extract_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "crf":
        return crf_golden()
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
    crf_golden()
    emitter_extract_golden()


if __name__ == "__main__":
    main()
