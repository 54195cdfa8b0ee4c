"""Scratch: where does the BRDF-parameter gradient of path_tracing_single differ? CUDA vs oracle autograd vs golden."""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core
from oracle import estimators as E, field as OF
from oracle.intersect import OracleScene
from tests.golden import cases
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
c = cases.build(name); sc = c["sc"]; dev = torch.device("cuda", 0); spp = c["spp"]
g = np.load(os.path.join("tests/golden", name + ".npz"))
osc = OracleScene(sc.vertices, sc.faces)
em = E.Emitter(sc.emitter_dict(), sc.slf_dict(c["H"]))
vmin, vmax = sc.voxel_bounds()
p = c["params"].clone().requires_grad_(True)
r = torch.as_tensor(c["rays"]); U = torch.as_tensor(c["U"][:, :8]); Gw = torch.as_tensor(c["Gw"])
leaves = {}
def mat_fn(x):
    m = OF.material(x, p, vmin, vmax)
    if not leaves:
        leaves["x"] = x.detach().clone()
        for k, v in m.items():
            v.retain_grad(); leaves[k] = v
    return m
L = E.path_tracing_single(osc, em, mat_fn, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp, U)
(L * Gw).sum().backward()
go = p.grad.numpy()
dmat_o = torch.cat([leaves["albedo"].grad, leaves["roughness"].grad, leaves["metallic"].grad], 1)
scene = core.Scene(sc.vertices, sc.faces, 0)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(c["H"]), c["params"], sc.voxel_bounds())
Lg, rec = core.single_forward(scene, tables, r.to(dev), spp, core.Sampler(U=U.to(dev)), True)
dp = torch.zeros(p.numel(), device=dev)
ws = torch.zeros(core.C.lib().iris_single_workspace_bytes(len(r), spp), dtype=torch.uint8, device=dev)
core.single_backward(tables, Gw.to(dev), spp, rec, True, dp, ws)
gc = dp.cpu().numpy()
n = len(r) * spp
dmat_c = ws[:20 * n].view(torch.float32).reshape(n, 5).cpu()
# isolated: field backward on the ORACLE's positions and d_mat
gi = core.field_backward(tables, leaves["x"].to(dev), dmat_o.to(dev)).cpu().numpy()
def rep(a, b, tag):
    out = []
    for nm, sl in (("W1", slice(0, 4096)), ("W2", slice(4096, 8192)), ("W3", slice(8192, 9216)), ("grid", slice(9216, None))):
        s = np.abs(b[sl]).max(); out.append("%s %.2e (scale %.2e)" % (nm, np.abs(a[sl] - b[sl]).max() / s, s))
    print(tag, " | ".join(out))
rep(gc, go, "cuda    vs oracle:")
rep(gc, np.concatenate([g["d_mlp"], go[9216:]]), "cuda    vs golden(mlp):")
rep(go, np.concatenate([g["d_mlp"], go[9216:]]), "oracle  vs golden(mlp):")
rep(gi, go, "field_bwd(oracle x,dmat) vs oracle:")
d = (dmat_c - dmat_o).abs(); s = dmat_o.abs().max(0).values
print("d_mat per lane: max abs err / column scale", (d.max(0).values / s).numpy(), "lanes with rel err>1e-2:", int(((d / s) > 1e-2).any(1).sum()), "of", n)
print("sum over lanes cuda", dmat_c.sum(0).numpy(), "oracle", dmat_o.sum(0).numpy())
recf = rec.view(torch.float32).reshape(6, n, 4)
x0 = recf[5, :, :3].contiguous(); code = recf[5, :, 3].contiguous().view(torch.int32)
act_o = (dmat_o.abs().sum(1) > 0)
print("code==-2:", int((code == -2).sum()), "oracle active(dmat!=0):", int(act_o.sum()), "x0 max abs diff on active:", float((x0.cpu() - leaves["x"])[act_o].abs().max()))
g2 = core.field_backward(tables, x0, dmat_c.to(dev)).cpu().numpy()
rep(g2, gc, "field_bwd(gpu x0, gpu dmat) vs combined:")
rep(g2, go, "field_bwd(gpu x0, gpu dmat) vs oracle:")
g3 = core.field_backward(tables, leaves["x"].to(dev), dmat_c.to(dev)).cpu().numpy()
rep(g3, go, "field_bwd(oracle x, gpu dmat) vs oracle:")
g4 = core.field_backward(tables, x0, dmat_o.to(dev)).cpu().numpy()
rep(g4, go, "field_bwd(gpu x0, oracle dmat) vs oracle:")
