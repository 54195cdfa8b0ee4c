"""Does ordering the ray queue by direction inside blocks help the persistent trace kernel?  Secondary rays of a 1280x960 view
(8 per pixel, consecutive), traced in generation order vs sorted by a 24-bin direction key inside blocks of 256 .. 16384 rays."""
import sys, json
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from iris_b200 import core, scenes
from quick_perf import ev_time
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
o, d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
t, prim, uv, p, n = scene.intersect_raw(o, d)
v = prim >= 0
g = torch.Generator(device=dev).manual_seed(0)
S = 32
pos = p[v].repeat_interleave(S, 0); nn = n[v].repeat_interleave(S, 0)
pos = pos[: (pos.shape[0] // 16384) * 16384]; nn = nn[: pos.shape[0]]
r = torch.nn.functional.normalize(torch.randn(pos.shape[0], 3, device=dev, generator=g), dim=-1)
sd = torch.nn.functional.normalize(nn + 0.999 * r, dim=-1).contiguous()
so = (pos + 1e-4 * nn).contiguous()
core.C.check(lib.iris_set_option(b"intersect_impl", 1))
out = {}
def rate(a, b): return round(a.shape[0] / ev_time(lambda: scene.intersect_raw(a, b), 3, 1) / 1e3, 1)
out["unsorted"] = rate(so, sd)
oct_ = ((sd[:, 0] < 0).long() << 2) | ((sd[:, 1] < 0).long() << 1) | (sd[:, 2] < 0).long()
dom = sd.abs().argmax(-1)
for nbins, key in (("oct8", oct_), ("oct_dom24", oct_ * 3 + dom)):
    for B in (256, 1024, 16384):
        k2 = key.view(-1, B)
        perm = torch.argsort(k2, dim=1, stable=True) + (torch.arange(k2.shape[0], device=dev) * B)[:, None]
        perm = perm.reshape(-1)
        out["%s_block%d" % (nbins, B)] = rate(so[perm].contiguous(), sd[perm].contiguous())
core.C.check(lib.iris_set_option(b"intersect_impl", 0))
out["static_unsorted"] = rate(so, sd)
print(json.dumps(out))
