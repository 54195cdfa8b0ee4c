"""Rates of the forward-only wavefront estimators (path_tracing, path_tracing_det_diff/_spec) with their per-kernel profile."""
import sys, json, ctypes
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from iris_b200 import core, scenes
from quick_perf import ev_time

def main():
    dev = torch.device("cuda", 0)
    lib = core.C.lib()
    sc = scenes.room(1_000_000, 16, seed=0)
    scene = core.Scene(sc.vertices, sc.faces, 0)
    params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
    params[:9216].uniform_(-0.2, 0.2)
    tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
    rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
    B = rays.shape[0]
    out = {}
    def prof(fn, n_samples, tag):
        for k in range(64):
            if not lib.iris_profile_name(k): break
            lib.iris_profile_read(k, None, None, 1)
        ms = ev_time(fn, 3, 1)
        out[tag + "_Msamples_s"] = round(n_samples / ms / 1e3, 1)
        lib.iris_profile_enable(1)
        fn(); torch.cuda.synchronize()
        p = {}
        for k in range(64):
            nm = lib.iris_profile_name(k)
            if not nm: break
            cnt, t = ctypes.c_int64(), ctypes.c_double()
            lib.iris_profile_read(k, ctypes.byref(cnt), ctypes.byref(t), 1)
            if cnt.value: p[nm.decode()] = [round(t.value, 2), cnt.value]
        lib.iris_profile_enable(0)
        out[tag + "_profile_ms"] = p
    spp, depth = 16, 5
    smp = core.Sampler(seed=5)
    prof(lambda: core.path_tracing(scene, tables, rays, spp, depth, smp), B * spp, "path_tracing_spp16_depth5")
    t, prim, uv, p, n = scene.intersect_raw(rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous())
    v = prim >= 0
    pos, nrm, wo, tri = p[v].contiguous(), n[v].contiguous(), (-rays[:, 3:6])[v].contiguous(), prim[v].contiguous()
    prof(lambda: core.path_tracing_det(scene, tables, 0, 0.0, pos, wo, nrm, tri, spp, depth, smp), pos.shape[0] * spp, "det_diff_spp16_depth5")
    prof(lambda: core.path_tracing_det(scene, tables, 1, 0.3, pos, wo, nrm, tri, spp, depth, smp), pos.shape[0] * spp, "det_spec_spp16_depth5")
    print(json.dumps(out))
main()
