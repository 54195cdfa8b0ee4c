"""Scratch timing of the individual kernels on one GPU (not the bench contract; see bench.py)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes

def ev_time(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def main():
    n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = torch.device("cuda", 0)
    out = {}
    t0 = time.time(); sc = scenes.room(n_tris, 16, seed=0); out["gen_s"] = time.time() - t0
    t0 = time.time(); scene = core.Scene(sc.vertices, sc.faces, 0); out["scene_s"] = time.time() - t0
    out["stats"] = scene.stats()
    if len(sys.argv) > 2 and sys.argv[2] == "lbvh":
        t0 = time.time(); scene = core.Scene(sc.vertices, sc.faces, 0, builder=1); out["scene_lbvh_first_s"] = time.time() - t0
        t0 = time.time(); scene = core.Scene(sc.vertices, sc.faces, 0, builder=1); out["scene_lbvh_second_s"] = time.time() - t0
        out["stats_lbvh"] = scene.stats()
    H = 256
    t0 = time.time(); slf = sc.slf_dict(H); out["slf_s"] = time.time() - t0
    params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
    params[:9216].uniform_(-0.2, 0.2)
    tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), slf, params, sc.voxel_bounds())
    W, Hh = 640, 480
    rays = torch.as_tensor(sc.camera_rays(W, Hh, view=1)).to(dev)
    o, d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
    ms = ev_time(lambda: scene.intersect_raw(o, d)); out["primary_Mrays_s"] = W * Hh / ms / 1e3
    t, prim, uv, p, n = scene.intersect_raw(o, d)
    valid = prim >= 0
    out["primary_hit_frac"] = float(valid.float().mean())
    pos, nrm, wo = p[valid].contiguous(), n[valid].contiguous(), (-d[valid]).contiguous()
    for spp in (64,):
        smp = core.Sampler(seed=1)
        ms = ev_time(lambda: core.bake(scene, tables, 0, 1.0, pos, nrm, None, spp, smp), 3, 1)
        out["bake_diffuse_Mrays_s_spp%d" % spp] = pos.shape[0] * spp / ms / 1e3
        for r in (0.02, 0.412, 1.0):
            ms = ev_time(lambda: core.bake(scene, tables, 1, r, pos, nrm, wo, spp, smp), 3, 1)
            out["bake_spec_r%.2f_Mrays_s" % r] = pos.shape[0] * spp / ms / 1e3
    # incoherent random rays
    g = torch.Generator(device=dev).manual_seed(0)
    lo = torch.tensor(sc.vertices.min(0), device=dev); hi = torch.tensor(sc.vertices.max(0), device=dev)
    N = 8_000_000
    ro = lo + (hi - lo) * (0.1 + 0.8 * torch.rand(N, 3, device=dev, generator=g))
    rd = torch.nn.functional.normalize(torch.randn(N, 3, device=dev, generator=g), dim=-1)
    ms = ev_time(lambda: scene.intersect_raw(ro, rd), 3, 1); out["random_Mrays_s"] = N / ms / 1e3
    x = ro[:4_000_000].contiguous()
    ms = ev_time(lambda: core.field_forward(tables, x), 3, 1); out["field_fwd_Msamples_s"] = x.shape[0] / ms / 1e3
    spp = 32
    ws = torch.empty(core.C.lib().iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
    smp = core.Sampler(seed=3)
    ms = ev_time(lambda: core.single_forward(scene, tables, rays, spp, smp, False, ws), 3, 1); out["single_fwd_Msamples_s"] = rays.shape[0] * spp / ms / 1e3
    ms = ev_time(lambda: core.single_forward(scene, tables, rays, spp, smp, True, ws), 3, 1); out["single_fwd_rec_Msamples_s"] = rays.shape[0] * spp / ms / 1e3
    L, rec = core.single_forward(scene, tables, rays, spp, smp, True, ws)
    dL = torch.randn_like(L)
    ms = ev_time(lambda: core.single_backward(tables, dL, spp, rec), 3, 1); out["single_bwd_emitter_Msamples_s"] = rays.shape[0] * spp / ms / 1e3
    dp = torch.zeros(9216 + 27954112, device=dev)
    ws2 = torch.empty(core.C.lib().iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
    ms = ev_time(lambda: core.single_backward(tables, dL, spp, rec, True, dp, ws2), 3, 1); out["single_bwd_brdf_Msamples_s"] = rays.shape[0] * spp / ms / 1e3
    dm = torch.randn(x.shape[0], 5, device=dev) * 1e-4
    wsf = torch.empty(core.C.lib().iris_field_backward_workspace_bytes(x.shape[0]), dtype=torch.uint8, device=dev)
    ms = ev_time(lambda: core.field_backward(tables, x, dm, dp, wsf), 3, 1); out["field_bwd_Msamples_s"] = x.shape[0] / ms / 1e3
    out["L_mean"] = float(L.mean()); out["L_finite"] = bool(torch.isfinite(L).all())
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
