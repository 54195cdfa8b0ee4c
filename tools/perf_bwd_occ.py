"""Fused field adjoint (k_field_backward_tc5v2) with one and with two CTAs per SM: how does its rate scale with the tiles in flight?"""
import sys, json, ctypes
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(200_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(64), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1))[:262144].to(dev)
spp = 32
L, rec = core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), True)
dL = torch.randn_like(L)
dp = torch.zeros(9216 + 27954112, device=dev)
ws = torch.empty(lib.iris_single_workspace_bytes(rays.shape[0], spp), dtype=torch.uint8, device=dev)
n = rays.shape[0] * spp
for ctas in (2, 1):
    core.C.check(lib.iris_set_option(b"tc5_bwd_ctas_per_sm", ctas))
    for _ in range(2):
        core.single_backward(tables, dL, spp, rec, True, dp, ws)
    torch.cuda.synchronize()
    lib.iris_profile_enable(1)
    for k in range(64):
        if not lib.iris_profile_name(k): break
        lib.iris_profile_read(k, None, None, 1)
    for _ in range(3):
        core.single_backward(tables, dL, spp, rec, True, dp, ws)
    torch.cuda.synchronize()
    for k in range(64):
        nm = lib.iris_profile_name(k)
        if not nm: break
        c, t = ctypes.c_int64(), ctypes.c_double()
        lib.iris_profile_read(k, ctypes.byref(c), ctypes.byref(t), 1)
        if c.value and b"tc5" in nm: print("ctas/SM %d: %s %.3f ms per launch = %.2f G samples/s" % (ctas, nm.decode(), t.value / c.value, n / (t.value / c.value) / 1e6))
    lib.iris_profile_enable(0)
core.C.check(lib.iris_set_option(b"tc5_bwd_ctas_per_sm", 2))
