"""A/B of path_tracing_single's secondary bounce: fused kernel vs wavefront (gen -> ray-queue trace -> shade), c3-like launch."""
import sys, json
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
    spp = 32
    n = rays.shape[0] * spp
    smp = core.Sampler(seed=3)
    out = {}
    def run(tag, rec):
        ws = torch.empty(lib.iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
        ms = ev_time(lambda: core.single_forward(scene, tables, rays, spp, smp, rec, ws), 3, 1)
        out[tag] = round(n / ms / 1e3, 1)
    import os
    quick = os.environ.get("IRIS_PERF_QUICK") == "1"
    core.C.check(lib.iris_set_option(b"single_impl", 0))
    run("fused_rec", True)
    core.C.check(lib.iris_set_option(b"single_impl", 1))
    for log2 in ((23,) if quick else (19, 20, 21, 22, 23, 26)):
        core.C.check(lib.iris_set_option(b"single_chunk_log2", log2))
        run("wave_c%d" % log2, False); run("wave_rec_c%d" % log2, True)
    core.C.check(lib.iris_set_option(b"single_chunk_log2", 23))
    for ctas in (() if quick else (6, 10)):
        core.C.check(lib.iris_set_option(b"persist_ctas_per_sm", ctas))
        run("wave_rec_c21_ctas%d" % ctas, True)
    core.C.check(lib.iris_set_option(b"persist_ctas_per_sm", 8))
    lib.iris_profile_enable(1)
    ws = torch.empty(lib.iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
    for _ in range(3): core.single_forward(scene, tables, rays, spp, smp, True, ws)
    torch.cuda.synchronize()
    prof = {}
    import ctypes
    for k in range(32):
        nm = lib.iris_profile_name(k)
        if not nm: break
        ms, cnt = ctypes.c_double(), ctypes.c_int64()
        lib.iris_profile_read(k, ctypes.byref(cnt), ctypes.byref(ms), 0)
        if cnt.value: prof[nm.decode()] = [round(ms.value / 3, 3), cnt.value // 3]
    out["profile_ms_per_call"] = prof
    print(json.dumps(out))
main()
