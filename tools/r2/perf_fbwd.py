"""Time the fused field adjoint kernels (impl 1 vs 2) on 8.4M coherent samples, events from the library's own profiler."""
import sys, ctypes
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
spp = 8
dp = torch.zeros(9216 + 27954112, device=dev)
L, rec = core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), True)
g = torch.randn_like(L)
names = {}
k = 0
while lib.iris_profile_name(k):
    names[lib.iris_profile_name(k).decode()] = k
    k += 1
for impl in (1, 2, 1, 2):
    core.C.check(lib.iris_set_option(b"field_backward_impl", impl))
    for _ in range(2):
        core.single_backward(tables, g, spp, rec, True, dp)
    torch.cuda.synchronize()
    lib.iris_profile_enable(1)
    for kk in names.values():
        lib.iris_profile_read(kk, None, None, 1)
    for _ in range(5):
        core.single_backward(tables, g, spp, rec, True, dp)
    torch.cuda.synchronize()
    out = {}
    for nm in ("k_field_backward_tc5", "k_field_backward_scatter", "k_single_backward"):
        n_, t_ = ctypes.c_int64(), ctypes.c_double()
        lib.iris_profile_read(names[nm], ctypes.byref(n_), ctypes.byref(t_), 1)
        out[nm] = (n_.value, t_.value / max(n_.value, 1))
    lib.iris_profile_enable(0)
    n = rays.shape[0] * spp
    print("impl", impl, {k: "%d x %.3f ms" % v for k, v in out.items()}, "tc5: %.2f G samples/s" % (n / (out["k_field_backward_tc5"][1] * out["k_field_backward_tc5"][0] / (out["k_field_backward_tc5"][0] / 1) * 1e-3) / 1e9 * 1 if out["k_field_backward_tc5"][0] else 0))
print("d_params L1", float(dp.abs().sum()))
