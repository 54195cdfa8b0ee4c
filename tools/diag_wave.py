import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core
from oracle import estimators as E, field as OF
from oracle.intersect import OracleScene
from tests.golden import cases
from tests.conftest import rel_close
c = cases.build("small"); sc = c["sc"]; dev = torch.device("cuda", 0); spp = c["spp"]; depth = c["depth"]
osc = OracleScene(sc.vertices, sc.faces)
em = E.Emitter(sc.emitter_dict(), sc.slf_dict(c["H"]))
vmin, vmax = sc.voxel_bounds()
mat_fn = lambda x: OF.material(x, c["params"], vmin, vmax)
scene = core.Scene(sc.vertices, sc.faces, 0)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(c["H"]), c["params"], sc.voxel_bounds())
r = torch.as_tensor(c["rays"]); U = torch.as_tensor(c["U"])
for d in (0, 1, 2):
    Ud = U[:, :8 + 6 * d].contiguous()
    with torch.no_grad():
        ref = E.path_tracing(osc, em, mat_fn, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp, d, Ud)
    got = core.path_tracing(scene, tables, r.to(dev), spp, d, core.Sampler(U=Ud.to(dev))).cpu()
    print("path_tracing depth", d, rel_close(got.numpy(), ref.numpy()))
    # per lane
    rl = r.repeat_interleave(spp, 0)
    with torch.no_grad():
        refl = E.path_tracing(osc, em, mat_fn, rl[:, 0:3], rl[:, 3:6], rl[:, 6:9], rl[:, 9:12], 1, d, Ud)
    gotl = core.path_tracing(scene, tables, rl.to(dev), 1, d, core.Sampler(U=Ud.to(dev))).cpu()
    err = (gotl - refl).abs() / torch.maximum(refl.abs(), torch.tensor(1e-5))
    bad = (err > 1e-3).any(1)
    print("   per-lane bad", int(bad.sum()), "of", len(bad))
    for i in torch.nonzero(bad)[:6, 0].tolist():
        print("   ", i, gotl[i].numpy(), refl[i].numpy(), "U5", float(Ud[i, 5]))
pos, nrm, _, tri, _ = osc.ray_intersect(r[:, 0:3], r[:, 3:6])
n = len(pos)
Ui = U[:n, :6 * depth].contiguous()
with torch.no_grad():
    ref = E.trace_indirect(osc, em, mat_fn, pos, -r[:, 3:6], nrm, torch.ones(n, dtype=torch.bool), Ui, depth)
got = core.trace_indirect(scene, tables, pos.to(dev), (-r[:, 3:6]).to(dev), nrm.to(dev), depth, core.Sampler(U=Ui.to(dev))).cpu()
print("trace_indirect", rel_close(got.numpy(), ref.numpy()))
