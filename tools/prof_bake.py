"""One diffuse and one specular (roughness 0.41) bake launch of a 640x480 view, spp 64, for ncu captures of k_bake_persistent."""
import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(640, 480, view=1)).to(dev)
t, prim, uv, p, n = scene.intersect_raw(rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous())
v = prim >= 0
pos, nrm, wo = p[v].contiguous(), n[v].contiguous(), (-rays[:, 3:6])[v].contiguous()
smp = core.Sampler(seed=1)
for _ in range(2):
    a = core.bake(scene, tables, 0, 1.0, pos, nrm, None, 64, smp)
    b = core.bake(scene, tables, 1, 0.412, pos, nrm, wo, 64, smp)
torch.cuda.synchronize()
print(float(a.mean()), float(b[0].mean()))
