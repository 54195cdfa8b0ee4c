"""path_tracing depth 5 at 1280x960 x spp 16: live-lane lists on/off, random-init field (few terminations) and a varied field."""
import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
spp, depth = 16, 5
ws = torch.empty(lib.iris_wave_workspace_bytes(rays.shape[0] * spp), dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
for name, amp, wamp in (("random-init field (roughness ~0.51 everywhere: lanes die only on emitters / misses)", 1e-4, 0.2),
                        ("varied field (grid U(-0.5,0.5))", 0.5, 0.2),
                        ("strongly varied field (grid U(-0.5,0.5), MLP weights U(-1.5,1.5): roughness spread over (0.02,1), SLF terminations above 0.6)", 0.5, 1.5)):
    params = torch.empty(9216 + 27954112)
    params[:9216] = (torch.rand(9216, generator=g) * 2 - 1) * wamp
    params[9216:] = (torch.rand(27954112, generator=g) * 2 - 1) * amp
    tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
    out = {}
    for compact in (0, 1):
        core.C.check(lib.iris_set_option(b"wave_compact", compact))
        for s in range(2):
            L = core.path_tracing(scene, tables, rays, spp, depth, core.Sampler(seed=5), ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(3):
            L = core.path_tracing(scene, tables, rays, spp, depth, core.Sampler(seed=5), ws)
        e1.record()
        torch.cuda.synchronize()
        out[compact] = (e0.elapsed_time(e1) / 3, L.clone())
        if compact == 1:
            W_cnt = ws[22 * 16 * rays.shape[0] * spp + 64: 22 * 16 * rays.shape[0] * spp + 64 + 8 * 8].view(torch.int64).tolist()
    n = rays.shape[0] * spp
    print(name)
    print("  all lanes: %.1f ms (%.0f M samples/s)   live-lane lists: %.1f ms (%.0f M samples/s)   speed-up %.2fx   max |diff| %.2e" %
          (out[0][0], n / out[0][0] / 1e3, out[1][0], n / out[1][0] / 1e3, out[0][0] / out[1][0], float((out[0][1] - out[1][1]).abs().max())))
    print("  live lanes entering bounce 0..6 (of %d):" % n, W_cnt[:7])
