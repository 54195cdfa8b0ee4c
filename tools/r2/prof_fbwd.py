"""One backward over 8.4M coherent samples for ncu captures of the fused field adjoint + scatter."""
import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)[: (1 << 20)]
spp = 8
dp = torch.zeros(9216 + 27954112, device=dev)
L, rec = core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), True)
g = torch.randn_like(L)
for _ in range(2):
    core.single_backward(tables, g, spp, rec, True, dp)
torch.cuda.synchronize()
print(float(dp.abs().sum()))
