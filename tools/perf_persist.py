import os
"""A/B: k_intersect (one ray per lane) vs k_intersect_persistent (dynamic ray fetch) on three ray populations."""
import sys, json
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from iris_b200 import core, scenes
from quick_perf import ev_time

def main():
    dev = torch.device("cuda", 0)
    lib = core.C.lib()
    sc = scenes.room(1_000_000, 16, seed=0)
    scene = core.Scene(sc.vertices, sc.faces, 0, builder=int(os.environ.get('IRIS_BUILDER', '1')))
    print(scene.stats())
    rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
    o, d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
    t, prim, uv, p, n = scene.intersect_raw(o, d)
    v = prim >= 0
    g = torch.Generator(device=dev).manual_seed(0)
    # secondary rays: cosine-ish hemisphere directions from the primary hits, 8 per hit (bake-like population)
    pos = p[v].repeat_interleave(8, 0); nn = n[v].repeat_interleave(8, 0)
    r = torch.nn.functional.normalize(torch.randn(pos.shape[0], 3, device=dev, generator=g), dim=-1)
    sd = torch.nn.functional.normalize(nn + 0.999 * r, dim=-1)
    so = pos + 1e-4 * nn
    lo = torch.tensor(sc.vertices.min(0), device=dev); hi = torch.tensor(sc.vertices.max(0), device=dev)
    N = 8_000_000
    ro = lo + (hi - lo) * (0.1 + 0.8 * torch.rand(N, 3, device=dev, generator=g))
    rd = torch.nn.functional.normalize(torch.randn(N, 3, device=dev, generator=g), dim=-1)
    pops = {"primary": (o, d), "secondary": (so.contiguous(), sd.contiguous()), "random": (ro, rd)}
    out = {}

    if os.environ.get("IRIS_CARVEOUT"):
        core.C.check(lib.iris_set_option(b"trace_smem_carveout_pct", int(os.environ["IRIS_CARVEOUT"])))
    for name, (a, b) in pops.items():
        core.C.check(lib.iris_set_option(b"intersect_impl", 0))
        ref = scene.intersect_raw(a, b)
        ms0 = ev_time(lambda: scene.intersect_raw(a, b), 5, 2)
        out[name + "_base"] = round(a.shape[0] / ms0 / 1e3, 1)
        core.C.check(lib.iris_set_option(b"intersect_impl", 1))
        for ctas in (8,):
            core.C.check(lib.iris_set_option(b"persist_ctas_per_sm", ctas))
            got = scene.intersect_raw(a, b)
            same = all(bool(torch.equal(x, y)) for x, y in zip(ref, got))
            ms1 = ev_time(lambda: scene.intersect_raw(a, b), 5, 2)
            out[name + "_persist%d" % ctas] = round(a.shape[0] / ms1 / 1e3, 1)
            out[name + "_same%d" % ctas] = same
    print(json.dumps(out))

main()
