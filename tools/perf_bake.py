"""bake_impl 0 (fused, block-sorted), 1 (ray queue), 2 (persistent in-kernel) on a 640x480 view of the 1M-triangle room, spp 64."""
import sys, json
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from iris_b200 import core, scenes
from quick_perf import ev_time
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(640, 480, view=1)).to(dev)
t, prim, uv, p, n = scene.intersect_raw(rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous())
v = prim >= 0
pos, nrm, wo = p[v].contiguous(), n[v].contiguous(), (-rays[:, 3:6])[v].contiguous()
out = {}
spp = 64
smp = core.Sampler(seed=1)
for impl in (0, 1, 2):
    core.C.check(lib.iris_set_option(b"bake_impl", impl))
    ms = ev_time(lambda: core.bake(scene, tables, 0, 1.0, pos, nrm, None, spp, smp), 3, 1)
    out["impl%d_diffuse" % impl] = round(pos.shape[0] * spp / ms / 1e3, 1)
    for r in (0.02, 0.412, 1.0):
        ms = ev_time(lambda: core.bake(scene, tables, 1, r, pos, nrm, wo, spp, smp), 3, 1)
        out["impl%d_spec%.2f" % (impl, r)] = round(pos.shape[0] * spp / ms / 1e3, 1)
core.C.check(lib.iris_set_option(b"bake_impl", 2))
print(json.dumps(out))
