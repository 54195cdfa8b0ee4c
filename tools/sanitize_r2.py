"""Small workload for compute-sanitizer over the kernels added in round 2: the SAH-treelet device builder (all three forms), the denoiser,
and one path_tracing_single forward + adjoint on the resulting scene."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes, denoise
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(30_000, 16, seed=2, irregular=True)
for treelets, top in ((1, 0), (1, 1), (0, 0)):
    core.C.check(lib.iris_set_option(b"lbvh_sah_treelets", treelets))
    core.C.check(lib.iris_set_option(b"lbvh_sah_top", top))
    scene = core.Scene(sc.vertices, sc.faces, 0, builder=1)
    print("built", treelets, top, scene.stats()["n_nodes"], scene.stats()["max_depth"])
core.C.check(lib.iris_set_option(b"lbvh_sah_treelets", 1))
core.C.check(lib.iris_set_option(b"lbvh_sah_top", 0))
scene = core.Scene(sc.vertices, sc.faces, 0)
if len(sys.argv) > 1:      # e.g. 1: the first-generation fused adjoint (per-thread stores of dx^ instead of bulk tensor stores, which initcheck does not track)
    core.C.check(lib.iris_set_option(b"field_backward_impl", int(sys.argv[1])))
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(32), params, sc.voxel_bounds())
rays = torch.as_tensor(sc.camera_rays(48, 36, view=1)).to(dev)
L, rec = core.single_forward(scene, tables, rays, 4, core.Sampler(seed=1), True)
dp = torch.zeros(9216 + 27954112, device=dev)
core.single_backward(tables, torch.randn_like(L), 4, rec, True, dp)
img = torch.rand(37, 53, 3, device=dev)
nrm = torch.nn.functional.normalize(torch.randn(37, 53, 3, device=dev), dim=-1)
out = denoise.atrous(img, nrm, torch.rand(37, 53, 3, device=dev), iterations=3, sigma_c=0.5)
torch.cuda.synchronize()
print("ok", float(L.mean()), float(dp.abs().sum()), float(out.mean()))
