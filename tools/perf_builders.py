"""Ray-cast rates of path_tracing_single on the regular and on the scan-like irregular 1M-triangle room, host binned-SAH builder vs
device LBVH: per-kernel rates from the library's events (k_primary: camera rays, k_trace_queue: shadow + BSDF rays)."""
import sys, ctypes, time
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
lib = core.C.lib()
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
spp = 32
for irregular in (False, True):
    sc = scenes.room(1_000_000, 16, seed=0, irregular=irregular)
    tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
    views = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1]
    rays = torch.cat([torch.as_tensor(sc.camera_rays(1280, 960, view=v)) for v in views])
    if len(views) == 1:
        rays = rays[:262144 * 2]
    rays = rays.to(dev)
    n = rays.shape[0] * spp
    ws = torch.empty(lib.iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
    for builder, treelets, top, label in ((0, 1, 1, "host binned SAH"), (1, 0, 0, "device: Morton LBVH"), (1, 1, 0, "device: LBVH top, SAH treelets"),
                                          (1, 1, 1, "device: SAH top + SAH treelets")):
        core.C.check(lib.iris_set_option(b"lbvh_sah_treelets", treelets))
        core.C.check(lib.iris_set_option(b"lbvh_sah_top", top))
        t0 = time.time()
        scene = core.Scene(sc.vertices, sc.faces, 0, builder=builder)
        tb = time.time() - t0
        st = scene.stats()
        for _ in range(2):
            core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), False, ws)
        torch.cuda.synchronize()
        lib.iris_profile_enable(1)
        for k in range(64):
            if not lib.iris_profile_name(k): break
            lib.iris_profile_read(k, None, None, 1)
        for _ in range(3):
            core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), False, ws)
        torch.cuda.synchronize()
        out = {}
        for k in range(64):
            nm = lib.iris_profile_name(k)
            if not nm: break
            c, t = ctypes.c_int64(), ctypes.c_double()
            lib.iris_profile_read(k, ctypes.byref(c), ctypes.byref(t), 1)
            if c.value: out[nm.decode()] = t.value / 3
        lib.iris_profile_enable(0)
        print("%-9s %-32s build %.2f s (library %.0f ms)  nodes %d depth %d | primary %.2f G rays/s  trace_queue %.2f G rays/s  forward %.1f M samples/s"
              % ("irregular" if irregular else "regular", label, tb, st["build_ms"], st["n_nodes"], st["max_depth"],
                 n / out["k_primary"] / 1e6, 2 * n / out["k_trace_queue"] / 1e6, n / sum(out.values()) / 1e3))
        del scene
