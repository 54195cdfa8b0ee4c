"""Shared, seeded input builders for the golden vectors (used by make_golden.py and by the tests).

Everything is regenerated from seeds -- only OUTPUTS of the reference are stored in the .npz files.
"""
from __future__ import annotations

import numpy as np
import torch

from iris_b200 import scenes
from oracle import field as OF


def golden_params(seed=1):
    """A 'trained-looking' BRDF field: Xavier MLP, grid values U(-0.5,0.5) so that albedo / roughness /
    metallic vary over the scene (tcnn's own init U(+-1e-4) gives a constant 0.5 everywhere)."""
    p = OF.init_params(0)
    g = torch.Generator().manual_seed(seed)
    p[OF.N_MLP:] = (torch.rand(OF.N_GRID, generator=g) * 2 - 1) * 0.5
    return p


CASES = {
    # name: scene, image side, spp, SLF H, sample seed, extra uniform columns
    "small": dict(scene="cornell", side=16, spp=4, H=64, useed=5, depth=2),
    "c1": dict(scene="cornell", side=64, spp=16, H=256, useed=11, depth=0),
}


def build(name):
    c = dict(CASES[name])
    sc = scenes.cornell(seed=0)
    rays = sc.camera_rays(c["side"], c["side"])
    rng = np.random.default_rng(c["useed"])
    n = c["side"] * c["side"] * c["spp"]
    U = rng.random((n, 8 + 6 * c["depth"]), dtype=np.float32)
    # the reference hazard u > cdf[K-1] is never injected
    U = np.minimum(U, np.float32(0.99999))
    Gw = rng.standard_normal((c["side"] * c["side"], 3)).astype(np.float32)
    c.update(sc=sc, rays=rays, U=U, Gw=Gw, params=golden_params())
    return c


def grid_fingerprint(g):
    """Compact summary of a 28M-entry gradient: per-level sums + the 4096 largest-magnitude entries."""
    g = np.asarray(g, np.float64)
    lv_sum, lv_abs = [], []
    for (_, _, size, off) in OF.LEVELS:
        s = g[off * 2:(off + size) * 2]
        lv_sum.append(s.sum())
        lv_abs.append(np.abs(s).sum())
    top = np.argsort(-np.abs(g))[:4096]
    top.sort()
    return np.array(lv_sum), np.array(lv_abs), top.astype(np.int64), g[top].astype(np.float32)


def shading_inputs(n=4096, R=6, seed=21):
    """Seeded inputs of the training-step shading case (also used by the tests): material in the ranges NGPBRDF produces,
    roughness rows 0..2R-1 pinned to the interval ends and the baked levels themselves."""
    g = torch.Generator().manual_seed(seed)
    albedo = torch.rand(n, 3, generator=g)
    roughness = torch.rand(n, 1, generator=g) * 0.98 + 0.02
    metallic = torch.rand(n, 1, generator=g)
    levels = torch.linspace(0.02, 1.0, R)
    roughness[:R, 0] = levels
    roughness[R:2 * R, 0] = (levels + 1e-4).clamp(max=1.0)
    diffuse = torch.rand(n, 3, generator=g) * 2.0
    specular0 = torch.rand(n, R, 3, generator=g) * 3.0
    specular1 = torch.rand(n, R, 3, generator=g) * 0.5
    dL = torch.randn(n, 3, generator=g)
    return dict(albedo=albedo, roughness=roughness, metallic=metallic, diffuse=diffuse, specular0=specular0, specular1=specular1, dL=dL)


def slf_inputs(n_views=3, n=30000, seed=33):
    """Seeded inputs of the SLF-bake case: per view, points in a box (with a few far outliers on one view so that the bounds are
    not symmetric), a validity mask with ~10% misses, radiance.  One view has no valid point at all."""
    g = torch.Generator().manual_seed(seed)
    views, rads = [], []
    for v in range(n_views + 1):
        pos = torch.rand(n, 3, generator=g) * torch.tensor([3.0, 2.4, 3.4]) + torch.tensor([-1.2, 0.1, -0.9])
        valid = torch.rand(n, generator=g) > 0.1
        if v == 1:
            pos[:5] = torch.tensor([[2.6, 2.9, -1.4], [-1.45, 0.0, 0.0], [0.0, 0.0, 2.95], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0]])
            valid[:5] = True
        if v == n_views:
            valid[:] = False
        views.append((pos, valid))
        rads.append(torch.rand(n, 3, generator=g) * 4.0)
    return views, rads


def crf_inputs(n=6000, seed=41):
    """Seeded inputs of the EmorCRF case: HDR values that under- and overshoot [0,1] after exposure, per-row exposures, bin-edge
    values, a weight set of realistic size, LDR values for the inverse."""
    g = torch.Generator().manual_seed(seed)
    hdr = torch.rand(n, 3, generator=g) * 1.6 - 0.2
    hdr[:1024, 0] = torch.linspace(0, 1, 1024)                 # exactly on the knots (exposure 1 below)
    hdr[0, 1], hdr[1, 1], hdr[2, 1] = 0.0, 1.0, 1.0 - 2.0 ** -24
    exposure = torch.rand(n, 1, generator=g) * 1.5 + 0.25
    exposure[:1024] = 1.0
    weight = (torch.rand(3, 11, generator=g) - 0.5) * 0.2
    d_ldr = torch.randn(n, 3, generator=g)
    ldr = torch.rand(n, 3, generator=g) * 1.2 - 0.1
    ldr[:1024, 2] = torch.linspace(0, 1, 1024)
    return dict(hdr=hdr, exposure=exposure, weight=weight, d_ldr=d_ldr, ldr=ldr)


def emitter_extract_inputs(n_views=3, side=40, seed=51):
    """Seeded inputs of the emitter-extraction case (extract_emitter_ldr.py:76-110): camera rays of a few views of the Cornell room and
    LDR colours that saturate on the light, on every 53rd triangle, and NOT QUITE (mean just under the threshold) on every 59th."""
    from oracle.intersect import OracleScene
    sc = scenes.cornell(seed=0)
    osc = OracleScene(sc.vertices, sc.faces)
    g = torch.Generator().manual_seed(seed)
    views = []
    for v in range(n_views):
        rays = sc.camera_rays(side, side, view=v)
        prim = osc.intersect_raw(rays[:, 0:3], rays[:, 3:6])["prim"].astype(np.int64)
        rgb = torch.rand(len(rays), 3, generator=g) * 0.9
        hit = prim >= 0
        lit = hit & (sc.is_emitter[np.maximum(prim, 0)] | (prim % 53 == 0))
        rgb[torch.as_tensor(lit)] = torch.tensor([0.97, 1.0, 0.95])
        almost = torch.as_tensor(hit & (prim % 59 == 0) & ~lit)
        rgb[almost] = torch.tensor([0.9, 0.985, 0.7])
        views.append((rays, rgb.numpy()))
    return sc, views


def slf_refine_inputs(views, seed=37):
    """Seeded LDR colours + exposures for the refine pass (slf_refine.py:88-104) over the points of slf_inputs()."""
    g = torch.Generator().manual_seed(seed)
    ldr = [torch.rand(len(p), 3, generator=g) for p, _ in views]
    exposure = [float(torch.rand(1, generator=g)) + 0.5 for _ in views]
    return ldr, exposure
