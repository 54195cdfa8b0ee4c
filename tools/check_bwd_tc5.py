"""EXPERIMENTAL fused tcgen05 field adjoint (field_backward_impl 1) against the default two-kernel adjoint on random positions."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.cornell()
params = torch.empty(9216 + 27954112).uniform_(-1e-2, 1e-2, generator=torch.Generator().manual_seed(0))
params[:9216].uniform_(-0.3, 0.3, generator=torch.Generator().manual_seed(1))
tables = core.ShadingTables(dev).set_field(params, *sc.voxel_bounds())
g = torch.Generator().manual_seed(2)
for n in (128, 3000, 70001):
    lo, hi = sc.voxel_bounds()
    x = (lo + (hi - lo) * torch.rand(n, 3, generator=g)).to(dev)
    dmat = (torch.randn(n, 5, generator=g) * torch.exp(torch.rand(n, 1, generator=g) * -12)).to(dev)
    dmat[::7] = 0
    mat, enc = core.field_forward(tables, x, want_encoded=True)
    core.C.check(lib.iris_set_option(b"field_backward_impl", 0))
    ref = core.field_backward(tables, x, dmat, encoded=enc)
    core.C.check(lib.iris_set_option(b"field_backward_impl", 1))
    got = core.field_backward(tables, x, dmat, encoded=enc)
    core.C.check(lib.iris_set_option(b"field_backward_impl", 0))
    torch.cuda.synchronize()
    for name, sl in (("W1", slice(0, 4096)), ("W2", slice(4096, 8192)), ("W3", slice(8192, 9216)), ("grid", slice(9216, None))):
        a, b = got[sl], ref[sl]
        print(n, name, "max|ref| %.3e  max err / max|ref| %.3e" % (float(b.abs().max()), float((a - b).abs().max() / b.abs().max())))
