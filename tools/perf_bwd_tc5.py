"""Adjoint rate of path_tracing_single (field + emitter gradients) with the default two-kernel field adjoint vs the fused tcgen05 one."""
import sys, json, ctypes
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from iris_b200 import core, scenes
from quick_perf import ev_time
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(200_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(64), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
spp = 8
L, rec = core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), True)
dL = torch.randn_like(L)
dp = torch.zeros(9216 + 27954112, device=dev)
ws = torch.empty(lib.iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
out = {}
for impl in (0, 1):
    core.C.check(lib.iris_set_option(b"field_backward_impl", impl))
    ms = ev_time(lambda: core.single_backward(tables, dL, spp, rec, True, dp, ws), 3, 1)
    out["impl%d_Msamples_s" % impl] = round(rays.shape[0] * spp / ms / 1e3, 1)
    lib.iris_profile_enable(1)
    for k in range(64):
        if not lib.iris_profile_name(k): break
        lib.iris_profile_read(k, None, None, 1)
    core.single_backward(tables, dL, spp, rec, True, dp, ws); torch.cuda.synchronize()
    p = {}
    for k in range(64):
        nm = lib.iris_profile_name(k)
        if not nm: break
        c, t = ctypes.c_int64(), ctypes.c_double()
        lib.iris_profile_read(k, ctypes.byref(c), ctypes.byref(t), 1)
        if c.value: p[nm.decode()] = round(t.value, 3)
    lib.iris_profile_enable(0)
    out["impl%d_ms" % impl] = p
core.C.check(lib.iris_set_option(b"field_backward_impl", 0))
print(json.dumps(out))
