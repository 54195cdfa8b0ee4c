import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes
dev = torch.device("cuda", 0)
lib = core.C.lib()
sc = scenes.cornell()
g = torch.Generator().manual_seed(1)
params = torch.empty(9216 + 27954112).uniform_(-0.5, 0.5, generator=g)
tables = core.ShadingTables(dev).set_field(params, *sc.voxel_bounds())
x = (torch.rand(2_000_000, 3, generator=g) * 2 - 1).to(dev)
for impl in (0, 1, 0, 1):
    core.C.check(lib.iris_set_option(b"field_forward_impl", impl))
    core.field_forward(tables, x)
torch.cuda.synchronize()
