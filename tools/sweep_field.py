"""Sweep CTAs/SM and the shared-memory carveout of the tcgen05 field forward on camera-coherent and random positions."""
import sys, json
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
t, prim, uv, p, n = scene.intersect_raw(rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous())
g = torch.Generator(device=dev).manual_seed(0)
coh = (p.repeat_interleave(8, 0) + 1e-3 * torch.randn(p.shape[0] * 8, 3, device=dev, generator=g)).contiguous()     # 8 jittered samples per pixel
lo, hi = sc.voxel_bounds()
rnd = (lo + (hi - lo) * torch.rand(8_000_000, 3, device=dev, generator=g)).contiguous()
out = {}
for ctas in (3, 4, 5, 6):
    for pct in (50, 60, 70, 80, 90, 100):
        core.C.check(lib.iris_set_option(b"tc5_ctas_per_sm", ctas)); core.C.check(lib.iris_set_option(b"field_smem_carveout_pct", pct))
        a = ev_time(lambda: core.field_forward(tables, coh), 3, 1); b = ev_time(lambda: core.field_forward(tables, rnd), 3, 1)
        out["ctas%d_pct%d" % (ctas, pct)] = [round(coh.shape[0] / a / 1e3), round(rnd.shape[0] / b / 1e3)]
print(json.dumps(out))
