import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
from tools.quick_perf import ev_time
dev = torch.device("cuda", 0)
sc = scenes.room(200_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4); params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables(dev).set_field(params, *sc.voxel_bounds())
# positions: primary hits of a camera with 32 jittered samples per pixel (the pattern the estimator produces)
rays = torch.as_tensor(sc.camera_rays(512, 256, view=1)).to(dev)
spp = 32
o = rays[:, 0:3].repeat_interleave(spp, 0)
d = torch.nn.functional.normalize(rays[:, 3:6].repeat_interleave(spp, 0) + 0.0005 * torch.randn(len(o), 3, device=dev), dim=-1)
t, prim, uv, p, n = scene.intersect_raw(o, d)
x = p[prim >= 0].contiguous()
dm = torch.randn(x.shape[0], 5, device=dev) * 1e-4
dp = torch.zeros(9216 + 27954112, device=dev)
ws = torch.empty(core.C.lib().iris_field_backward_workspace_bytes(x.shape[0]), dtype=torch.uint8, device=dev)
ms = ev_time(lambda: core.field_backward(tables, x, dm, dp, ws), 3, 1)
print(json.dumps({"n": x.shape[0], "field_bwd_coherent_Msamples_s": x.shape[0] / ms / 1e3}))
ms = ev_time(lambda: core.field_forward(tables, x), 3, 1)
print(json.dumps({"field_fwd_coherent_Msamples_s": x.shape[0] / ms / 1e3}))
lib = core.C.lib()
for impl in (0, 1):
    for ctas in (2, 4, 6):
        for pct in (35, 50, 70):
            lib.iris_set_option(b"field_forward_impl", impl); lib.iris_set_option(b"tc5_ctas_per_sm", ctas); lib.iris_set_option(b"field_smem_carveout_pct", pct)
            ms = ev_time(lambda: core.field_forward(tables, x), 3, 1)
            print("impl", impl, "ctas", ctas, "carveout", pct, "coherent fwd Msamples/s %.0f" % (x.shape[0] / ms / 1e3))
        if impl == 0: break
