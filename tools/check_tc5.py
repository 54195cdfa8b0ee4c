import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
from tools.quick_perf import ev_time
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.cornell()
g = torch.Generator().manual_seed(1)
params = torch.empty(9216 + 27954112).uniform_(-0.5, 0.5, generator=g); params[:9216].uniform_(-0.2, 0.2, generator=g)
tables = core.ShadingTables(dev).set_field(params, *sc.voxel_bounds())
for n in (100, 128, 5000, 1_000_000):
    x = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
    core.C.check(lib.iris_set_option(b"field_forward_impl", 0))
    a = core.field_forward(tables, x)
    core.C.check(lib.iris_set_option(b"field_forward_impl", 1))
    b = core.field_forward(tables, x)
    torch.cuda.synchronize()
    d = (a - b).abs()
    print(n, "max abs diff", float(d.max()), "frac > 1e-3", float((d > 1e-3).float().mean()), "finite", bool(torch.isfinite(b).all()), "mean", float(b.mean()))
x = (torch.rand(4_000_000, 3, generator=g) * 2 - 1).to(dev)
for impl in (0, 1):
    core.C.check(lib.iris_set_option(b"field_forward_impl", impl))
    ms = ev_time(lambda: core.field_forward(tables, x), 5, 2)
    print("impl", impl, "random Msamples/s", x.shape[0] / ms / 1e3)

core.C.check(lib.iris_set_option(b"field_forward_impl", 1))
for dbg in (1, 2, 0):
    core.C.check(lib.iris_set_option(b"tc5_debug", dbg))
    ms = ev_time(lambda: core.field_forward(tables, x), 5, 2)
    print("tc5 debug", dbg, "(1 = no encode, 2 = no MLP)", "Msamples/s", x.shape[0] / ms / 1e3)

core.C.check(lib.iris_set_option(b"tc5_debug", 0))
for ctas in (2, 4, 6, 8):
    core.C.check(lib.iris_set_option(b"tc5_ctas_per_sm", ctas))
    ms = ev_time(lambda: core.field_forward(tables, x), 5, 2)
    print("tc5 ctas/SM", ctas, "Msamples/s", x.shape[0] / ms / 1e3)

core.C.check(lib.iris_set_option(b"tc5_ctas_per_sm", 4))
for pct in (100, 70, 50, 35, 20):
    core.C.check(lib.iris_set_option(b"field_smem_carveout_pct", pct))
    r = []
    for impl in (0, 1):
        core.C.check(lib.iris_set_option(b"field_forward_impl", impl))
        ms = ev_time(lambda: core.field_forward(tables, x), 5, 2)
        r.append(x.shape[0] / ms / 1e3)
    print("carveout %d%%: mma.sync %.0f  tcgen05 %.0f Msamples/s" % (pct, r[0], r[1]))
