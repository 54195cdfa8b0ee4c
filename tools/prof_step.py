"""One path_tracing_single forward + adjoint (field + emitter gradients) at 1280x960, spp 8, for ncu captures: the first
k_trace_queue launch traces one full 2^23-sample chunk (16.8 M rays), the launch size bench.py times."""
import os
import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
sc = scenes.room(int(os.environ.get("IRIS_PROF_TRIS", "1000000")), 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
dp = torch.zeros(9216 + 27954112, device=dev)
for _ in range(2):
    L, rec = core.single_forward(scene, tables, rays, 8, core.Sampler(seed=3), True)
    core.single_backward(tables, torch.randn_like(L), 8, rec, True, dp)
torch.cuda.synchronize()
print(float(L.mean()), float(dp.abs().sum()))
