import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
from tests.golden import cases
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.cornell()
params = cases.golden_params()
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(64), params, sc.voxel_bounds())
lo, hi = sc.voxel_bounds()
g = torch.Generator().manual_seed(12)
for n in (129, 5000, 40000, 70001, 300000):
    x = (lo + (hi - lo) * torch.rand(n, 3, generator=g)).to(dev)
    dmat = (torch.randn(n, 5, generator=g) * torch.exp(torch.rand(n, 1, generator=g) * -14)).to(dev)
    dmat[::5] = 0
    mat, enc = core.field_forward(tables, x, want_encoded=True)
    out = {}
    for impl in (0, 1, 2):
        core.C.check(lib.iris_set_option(b"field_backward_impl", impl))
        out[impl] = core.field_backward(tables, x, dmat, encoded=enc)
    core.C.check(lib.iris_set_option(b"field_backward_impl", 2))
    for name, sl in (("W1", slice(0, 4096)), ("W2", slice(4096, 8192)), ("W3", slice(8192, 9216)), ("grid", slice(9216, None))):
        b = out[0][sl]
        print(n, name, "impl1 vs 0: %.2e" % float((out[1][sl] - b).abs().max() / b.abs().max()), "impl2 vs 0: %.2e" % float((out[2][sl] - b).abs().max() / b.abs().max()),
              "impl2 vs 1: %.2e" % float((out[2][sl] - out[1][sl]).abs().max() / b.abs().max()))
