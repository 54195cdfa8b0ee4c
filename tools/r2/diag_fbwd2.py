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
n = 300000
x = (lo + (hi - lo) * torch.rand(n, 3, generator=g)).to(dev)
dmat = (torch.randn(n, 5, generator=g) * torch.exp(torch.rand(n, 1, generator=g) * -14)).to(dev)
dmat[::5] = 0
mat, enc = core.field_forward(tables, x, want_encoded=True)
wb = lib.iris_field_backward_workspace_bytes(n)
res = {}
for impl in (1, 2, 2):
    core.C.check(lib.iris_set_option(b"field_backward_impl", impl))
    ws = torch.zeros(wb, dtype=torch.uint8, device=dev)
    d = core.field_backward(tables, x, dmat, encoded=enc, workspace=ws)
    torch.cuda.synchronize()
    h = ws.view(torch.float16)
    dx = h[320 * n:384 * n].view(n, 64).float().cpu()
    s = ws[800 * n:804 * n].view(torch.float32).cpu()
    res.setdefault(impl, []).append((dx, s, d[:9216].cpu()))
dx1, s1, w1 = res[1][0]
for k, (dx2, s2, w2) in enumerate(res[2]):
    bad_s = (s1 != s2).nonzero().flatten()
    bad = ((dx1 - dx2).abs().max(1).values > 1e-3).nonzero().flatten()
    print("run", k, "rows with different s:", len(bad_s), "rows with different dx:", len(bad), "of", n)
    if len(bad):
        tiles = torch.unique(bad // 128)
        print("  bad tiles:", tiles[:40].tolist(), "count", len(tiles), " tile %% 296:", torch.unique(tiles % 296)[:40].tolist())
        t0 = int(tiles[0])
        rows = bad[(bad // 128) == t0]
        print("  first bad tile", t0, "bad rows in it", len(rows), "cols differing in first bad row:", ((dx1[rows[0]] - dx2[rows[0]]).abs() > 1e-3).nonzero().flatten().tolist())
        print("  k (tile order within CTA) of bad tiles:", torch.unique(tiles // 296).tolist())
