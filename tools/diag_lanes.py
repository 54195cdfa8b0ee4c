"""Per-lane comparison of path_tracing_single (CUDA) against the oracle on golden case `c1` / `small`: finds which lanes
differ and why (scratch diagnostic)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from iris_b200 import core
from oracle import estimators as E, field as OF
from oracle.intersect import OracleScene
from tests.golden import cases

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
c = cases.build(name); sc = c["sc"]; dev = torch.device("cuda", 0)
spp = c["spp"]
osc = OracleScene(sc.vertices, sc.faces)
em = E.Emitter(sc.emitter_dict(), sc.slf_dict(c["H"]))
vmin, vmax = sc.voxel_bounds()
mat_fn = lambda x: OF.material(x, c["params"], vmin, vmax)
r = torch.as_tensor(c["rays"]); U = torch.as_tensor(c["U"][:, :8])
with torch.no_grad():
    Lo, act, pn, wo, nn, bw, tri0 = E._first_bounce(osc, em, mat_fn, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp, U, 0.0, True)
scene = core.Scene(sc.vertices, sc.faces, 0)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(c["H"]), c["params"], sc.voxel_bounds())
rl = r.repeat_interleave(spp, 0).to(dev)
Lg, _ = core.single_forward(scene, tables, rl, 1, core.Sampler(U=U.to(dev)), False)
Lg = Lg.cpu()
err = (Lg - Lo).abs() / torch.maximum(torch.maximum(Lg.abs(), Lo.abs()), torch.tensor(1e-6 * float(Lo.abs().mean())))
bad = (err > 1e-3).any(1)
print("lanes", len(Lo), "bad", int(bad.sum()), "frac", float(bad.float().mean()))
# primary hit comparison
o, wi0 = E._camera(r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp, U)
t, prim, uv, p, n = scene.intersect_raw(o.to(dev), wi0.to(dev))
print("primary prim mismatch (same rays)", int((prim.cpu().long() != tri0).sum()))
idx = torch.nonzero(bad)[:, 0]
big = err.max(1).values
print("err quantiles over bad lanes", np.quantile(big[idx].numpy(), [0.1, 0.5, 0.9]) if len(idx) else None)
# split of the oracle radiance: direct emission / NEE / BSDF
with torch.no_grad():
    position, normal, _, tri, _ = osc.ray_intersect(o, wi0)
    L0, _, active = em.eval_emitter(position, tri)
    position_s, normal_s, wo_s = E._sanitize(active, position, normal, -wi0)
    mat = mat_fn(position_s)
    Lnee = torch.where(active[:, None], E._nee(osc, em, mat, position_s, normal_s, wo_s, U[:, 2], U[:, 3:5], active, 1e-6, True), torch.zeros_like(L0))
    wi_b, bpdf, bw = E.sample_brdf(U[:, 5], U[:, 6:8], wo_s, normal_s, mat)
    pn, nn2, _, tri_b, _ = osc.ray_intersect(position_s + E.RAY_EPSILON * wi_b, wi_b)
Lb = Lo - L0 - Lnee
order = torch.argsort(-big)[:14]
print("largest-error lanes: lane err | gpu | oracle | L0 | Lnee | Lbsdf | u_lobe roughness tri_b t-ish")
for i in order.tolist():
    print(i, round(float(big[i]), 4), Lg[i].numpy().round(5), Lo[i].numpy().round(5), L0[i].numpy().round(4), Lnee[i].numpy().round(5), Lb[i].numpy().round(5),
          round(float(U[i, 5]), 4), round(float(mat["roughness"][i]), 4), int(tri_b[i]), "diffG-O", (Lg[i] - Lo[i]).numpy().round(5))
for i in idx[:4].tolist():
    print(i, "gpu", Lg[i].numpy(), "oracle", Lo[i].numpy(), "U", U[i].numpy().round(4))
# material parity at the oracle's primary hits
pos, nrm, _, tri, _ = osc.ray_intersect(o, wi0)
mo = mat_fn(pos); mo = torch.cat([mo["albedo"], mo["roughness"], mo["metallic"]], 1)
mg = core.field_forward(tables, pos.to(dev)).cpu()
print("mat max abs diff", float((mo - mg).abs().max()), "lanes with mat diff > 4e-4:", int(((mo - mg).abs().max(1).values > 4e-4).sum()))
