"""Seeded procedural indoor scenes for parity tests and the bench (numpy only).

The reference ships no scenes; BASELINE.json's configs name
"synthetic procedurally generated indoor scenes": a Cornell-style room (~10k
triangles, C1), a furnished room (~1M triangles, C2-C4) and a 5M-triangle room
(C5).  Everything here is deterministic in `seed` (const.py:11 uses SEED=0).

What a scene provides, in the on-disk vocabulary of the reference:
  * mesh              vertices (V,3) f32, faces (F,3) i32  (face order == prim index,
                      extract_emitter_ldr.py:73-101 relies on that)
  * emitter file      is_emitter (F,) bool, emitter_vertices (K,3,3), emitter_area (K,),
                      emitter_normal (K,3), emitter_radiance (F,3)
                      (extract_emitter_ldr.py:96-115; note the F-row radiance quirk)
  * SLF file          mask (H,H,H) bool, voxel_min/voxel_max python floats, weight =
                      {inds (H,H,H) i64, radiance (n_occ,3) f32, count (n_occ,) i64}
                      (slf_bake.py:140-145, model/slf.py:30-39)
  * camera rays       (P,12) = [o(3) d(3) dxdu(3) dydv(3)] (utils/dataset/synthetic_ldr.py:50-57)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------- mesh pieces
def _warp01(n, rng):
    """n+1 increasing knots on [0,1] whose spacing varies smoothly by ~10x (scan meshes are dense where the scanner was close)."""
    x = np.linspace(0.0, 1.0, n + 1)
    dens = np.ones_like(x)
    for _ in range(3):
        dens += rng.uniform(0.6, 1.6) * np.exp(-0.5 * ((x - rng.uniform(0, 1)) / rng.uniform(0.05, 0.2)) ** 2)
    w = 1.0 / dens
    k = np.concatenate([[0.0], np.cumsum(0.5 * (w[1:] + w[:-1]))])
    return k / k[-1]


def _grid_quad(p0, du, dv, nu, nv, disp=None, rng=None):
    """Tessellate the parallelogram p0 + s*du + t*dv into nu x nv cells (2 triangles each).

    disp: optional callable (s,t)->offset along the quad normal (displaced walls).
    rng:  irregular ("scan-like") tessellation -- smoothly varying cell sizes, interior vertices jittered inside their cells,
          a random diagonal per cell.  The boundary vertices stay on the boundary, so neighbouring patches still meet.
    Returns (verts (n,3) f64, faces (m,3) i64).
    """
    if rng is not None:
        s, t = _warp01(nu, rng), _warp01(nv, rng)
    else:
        s = np.linspace(0.0, 1.0, nu + 1)
        t = np.linspace(0.0, 1.0, nv + 1)
    S, T = np.meshgrid(s, t, indexing="ij")
    if rng is not None and nu > 1 and nv > 1:
        ds = np.minimum(np.diff(s)[:-1], np.diff(s)[1:])[:, None]
        dt = np.minimum(np.diff(t)[:-1], np.diff(t)[1:])[None, :]
        S[1:-1, 1:-1] += rng.uniform(-0.4, 0.4, size=(nu - 1, nv - 1)) * ds
        T[1:-1, 1:-1] += rng.uniform(-0.4, 0.4, size=(nu - 1, nv - 1)) * dt
    P = p0[None, None, :] + S[..., None] * du[None, None, :] + T[..., None] * dv[None, None, :]
    if disp is not None:
        n = np.cross(du, dv)
        n = n / np.linalg.norm(n)
        P = P + disp(S, T)[..., None] * n[None, None, :]
    verts = P.reshape(-1, 3)
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
    a = idx[:-1, :-1].ravel()
    b = idx[1:, :-1].ravel()
    c = idx[1:, 1:].ravel()
    d = idx[:-1, 1:].ravel()
    if rng is not None:
        flip = rng.random(len(a)) < 0.5                             # the other diagonal: (a,b,d) + (b,c,d), same orientation
        f1 = np.where(flip[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
        f2 = np.where(flip[:, None], np.stack([b, c, d], 1), np.stack([a, c, d], 1))
        return verts, np.concatenate([f1, f2], 0)
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
    return verts, faces


def _box(center, half, rot_y, n, rng=None):
    """Axis box rotated about y, each face an n x n grid."""
    cx = np.asarray(center, np.float64)
    hx, hy, hz = half
    c, s = math.cos(rot_y), math.sin(rot_y)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    ex, ey, ez = R[:, 0] * hx, R[:, 1] * hy, R[:, 2] * hz
    parts = []
    for o, u, v in (
        (cx - ex - ey - ez, 2 * ey, 2 * ez),  # -x
        (cx + ex - ey - ez, 2 * ez, 2 * ey),  # +x
        (cx - ex - ey - ez, 2 * ez, 2 * ex),  # -y
        (cx - ex + ey - ez, 2 * ex, 2 * ez),  # +y
        (cx - ex - ey - ez, 2 * ex, 2 * ey),  # -z
        (cx - ex - ey + ez, 2 * ey, 2 * ex),  # +z
    ):
        parts.append(_grid_quad(o, u, v, n, n, None, rng))
    return parts


def _sphere(center, radius, n):
    """UV sphere with n latitude bands and 2n longitude segments."""
    th = np.linspace(0.0, math.pi, n + 1)
    ph = np.linspace(0.0, 2 * math.pi, 2 * n + 1)[:-1]
    TH, PH = np.meshgrid(th, ph, indexing="ij")
    P = np.stack([np.sin(TH) * np.cos(PH), np.cos(TH), np.sin(TH) * np.sin(PH)], -1) * radius
    verts = P.reshape(-1, 3) + np.asarray(center, np.float64)[None]
    m = 2 * n
    idx = np.arange((n + 1) * m).reshape(n + 1, m)
    a = idx[:-1, :]
    b = idx[1:, :]
    c = np.roll(idx[1:, :], -1, 1)
    d = np.roll(idx[:-1, :], -1, 1)
    f1 = np.stack([a.ravel(), b.ravel(), c.ravel()], 1)[m:]      # skip degenerate north cap
    f2 = np.stack([a.ravel(), c.ravel(), d.ravel()], 1)[:-m]     # skip degenerate south cap
    return [(verts, np.concatenate([f1, f2], 0))]


def _merge(parts):
    vs, fs, off = [], [], 0
    for v, f in parts:
        vs.append(v)
        fs.append(f + off)
        off += len(v)
    return np.concatenate(vs, 0), np.concatenate(fs, 0)


# ----------------------------------------------------------------------------- scene
@dataclass
class Scene:
    vertices: np.ndarray           # (V,3) f32
    faces: np.ndarray              # (F,3) i32
    is_emitter: np.ndarray         # (F,) bool
    emitter_radiance: np.ndarray   # (F,3) f32, rows [0,K) meaningful (reference quirk)
    half: tuple                    # room half extents
    seed: int
    name: str = ""
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def n_tris(self):
        return len(self.faces)

    @property
    def n_emitters(self):
        return int(self.is_emitter.sum())

    # ---- emitter file (extract_emitter_ldr.py:96-115) ----
    def emitter_dict(self):
        tri = self.vertices[self.faces[self.is_emitter]]          # (K,3,3)
        cr = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        nrm = np.linalg.norm(cr, axis=-1)
        return {
            "is_emitter": self.is_emitter.copy(),
            "emitter_vertices": tri.astype(np.float32),
            "emitter_area": (nrm / 2.0).astype(np.float32),
            "emitter_normal": (cr / np.maximum(nrm, 1e-12)[:, None]).astype(np.float32),
            "emitter_radiance": self.emitter_radiance.copy(),
        }

    # ---- SLF file (slf_bake.py:86-145) ----
    def voxel_bounds(self):
        lo = float(self.vertices.min())
        hi = float(self.vertices.max())
        return 1.1 * lo, 1.1 * hi                                  # slf_bake.py:88-90 (scene centred on origin)

    def slf_dict(self, H=256):
        key = ("slf", H)
        if key in self._cache:
            return self._cache[key]
        vmin, vmax = self.voxel_bounds()
        mask = surface_voxels(self.vertices, self.faces, vmin, vmax, H)
        kk, jj, ii = np.nonzero(mask)                              # model/slf.py:30-32 order (z,y,x)
        inds = -np.ones((H, H, H), np.int64)
        inds[kk, jj, ii] = np.arange(len(ii))
        rng = np.random.default_rng(self.seed + 7)
        rad = rng.uniform(0.0, 0.5, size=(len(ii), 3)).astype(np.float32)
        out = {
            "mask": mask, "voxel_min": vmin, "voxel_max": vmax,
            "weight": {"inds": inds, "radiance": rad, "count": np.ones(len(ii), np.int64)},
        }
        self._cache[key] = out
        return out

    # ---- cameras ----
    def camera_rays(self, width, height, view=0, fov_deg=60.0):
        """Pinhole rays in the reference layout (utils/dataset/synthetic_ldr.py:40-57)."""
        hx, hy, hz = self.half
        rng = np.random.default_rng(self.seed * 1000 + 17 + view)
        if view == 0:
            eye = np.array([0.05 * hx, 0.02 * hy, 0.92 * hz])
            tgt = np.array([0.0, -0.15 * hy, -hz])
        else:
            eye = np.array([rng.uniform(-0.55, 0.55) * hx, rng.uniform(-0.3, 0.5) * hy, rng.uniform(-0.55, 0.55) * hz])
            ang = rng.uniform(0, 2 * math.pi)
            tgt = eye + np.array([math.cos(ang), rng.uniform(-0.45, 0.15), math.sin(ang)])
        fwd = tgt - eye
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        f = 0.5 * width / math.tan(0.5 * math.radians(fov_deg))
        u = (np.arange(width) + 0.5 - 0.5 * width) / f
        v = -(np.arange(height) + 0.5 - 0.5 * height) / f
        U, Vv = np.meshgrid(u, v, indexing="xy")                   # (H,W)
        d = fwd[None, None] + U[..., None] * right[None, None] + Vv[..., None] * up[None, None]
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        P = width * height
        rays = np.empty((P, 12), np.float32)
        rays[:, 0:3] = eye[None]
        rays[:, 3:6] = d.reshape(P, 3)
        rays[:, 6:9] = (right / f)[None]
        rays[:, 9:12] = (-up / f)[None]
        return rays


def surface_voxels(vertices, faces, vmin, vmax, H, chunk=1 << 20):
    """Occupancy mask of voxels touched by the surface, indexed [z,y,x] like model/slf.py:49-53."""
    mask = np.zeros((H, H, H), bool)
    vox = (vmax - vmin) / H
    tri = vertices[faces].astype(np.float64)                        # (F,3,3)
    e = np.stack([np.linalg.norm(tri[:, 1] - tri[:, 0], axis=-1),
                  np.linalg.norm(tri[:, 2] - tri[:, 1], axis=-1),
                  np.linalg.norm(tri[:, 0] - tri[:, 2], axis=-1)], -1).max(-1)
    k = np.maximum(1, np.ceil(e / (0.45 * vox)).astype(np.int64))
    for kv in np.unique(k):
        sel = np.nonzero(k == kv)[0]
        i, j = np.meshgrid(np.arange(kv + 1), np.arange(kv + 1), indexing="ij")
        keep = (i + j) <= kv
        b1 = (i[keep] / kv)[None, :, None]
        b2 = (j[keep] / kv)[None, :, None]
        per = max(1, chunk // b1.shape[1])
        for s in range(0, len(sel), per):
            t = tri[sel[s:s + per]]
            p = t[:, None, 0] * (1 - b1 - b2) + t[:, None, 1] * b1 + t[:, None, 2] * b2
            g = np.clip(((p.reshape(-1, 3) - vmin) / (vmax - vmin) * H).astype(np.int64), 0, H - 1)
            mask[g[:, 2], g[:, 1], g[:, 0]] = True
    return mask


def make_room(n_tris=10_000, n_emitters=2, seed=0, half=(1.0, 1.0, 1.0), displaced=False,
              furniture=2, name="cornell", irregular=False):
    """Closed room with tessellated walls, `furniture` primitives and ceiling light quads.

    n_tris is a target; the exact count is whatever the tessellation yields (reported in .n_tris).
    n_emitters = K emitter triangles (two per ceiling quad).
    """
    rng = np.random.default_rng(seed)
    hx, hy, hz = half
    assert n_emitters % 2 == 0 and n_emitters >= 2
    nq = n_emitters // 2
    wall_frac = 0.7 if furniture else 1.0
    n_wall = max(1, int(round(math.sqrt(n_tris * wall_frac / 12.0))))
    amp = 0.004 * min(half) if displaced else 0.0

    irr = np.random.default_rng(seed + 4242) if irregular else None      # irregular: scan-like tessellation (see _grid_quad)

    def disp_fn(ph):
        if not displaced:
            return None
        if irregular:      # scanner noise on top of the smooth displacement: per-vertex, ~1/4 of the smooth amplitude
            return lambda S, T: amp * (np.sin(37.0 * S + ph) * np.cos(29.0 * T + 2 * ph) + 0.5 * np.sin(91.0 * S * T + ph)
                                       + 0.25 * irr.standard_normal(S.shape) * (np.minimum(np.minimum(S, 1 - S), np.minimum(T, 1 - T)) > 0))
        return lambda S, T: amp * (np.sin(37.0 * S + ph) * np.cos(29.0 * T + 2 * ph) + 0.5 * np.sin(91.0 * S * T + ph))

    c = np.array
    parts = [
        _grid_quad(c([-hx, -hy, -hz]), c([2 * hx, 0, 0]), c([0, 0, 2 * hz]), n_wall, n_wall, disp_fn(0.1), irr),   # floor
        _grid_quad(c([-hx, hy, -hz]), c([0, 0, 2 * hz]), c([2 * hx, 0, 0]), n_wall, n_wall, disp_fn(0.7), irr),    # ceiling
        _grid_quad(c([-hx, -hy, -hz]), c([0, 2 * hy, 0]), c([2 * hx, 0, 0]), n_wall, n_wall, disp_fn(1.3), irr),   # back  (-z)
        _grid_quad(c([-hx, -hy, hz]), c([2 * hx, 0, 0]), c([0, 2 * hy, 0]), n_wall, n_wall, disp_fn(1.9), irr),    # front (+z)
        _grid_quad(c([-hx, -hy, -hz]), c([0, 0, 2 * hz]), c([0, 2 * hy, 0]), n_wall, n_wall, disp_fn(2.5), irr),   # left  (-x)
        _grid_quad(c([hx, -hy, -hz]), c([0, 2 * hy, 0]), c([0, 0, 2 * hz]), n_wall, n_wall, disp_fn(3.1), irr),    # right (+x)
    ]
    n_so_far = sum(len(f) for _, f in parts)
    if furniture:
        per = max(12, (n_tris - n_so_far) // furniture)
        for i in range(furniture):
            kind = "box" if (i % 3) != 2 else "sphere"
            px = rng.uniform(-0.6, 0.6) * hx
            pz = rng.uniform(-0.6, 0.3) * hz
            if kind == "box":
                hw = rng.uniform(0.12, 0.3) * min(hx, hz)
                hh = rng.uniform(0.2, 0.6) * hy
                n = max(1, int(round(math.sqrt(per / 12.0))))
                parts += _box((px, -hy + hh + 1e-3 * hy, pz), (hw, hh, hw), rng.uniform(0, math.pi), n, irr)
            else:
                r = rng.uniform(0.12, 0.25) * min(half)
                n = max(3, int(round(math.sqrt(per / 4.0))))
                parts += _sphere((px, -hy + r + rng.uniform(0.0, 0.6) * hy, pz), r, n)
    n_geom = sum(len(f) for _, f in parts)
    # ceiling lights: nq quads on a regular layout, 1% below the ceiling, each 2 triangles
    cols = int(math.ceil(math.sqrt(nq)))
    rows = int(math.ceil(nq / cols))
    lw = 0.5 * hx / cols
    ld = 0.5 * hz / rows
    q = 0
    for r_ in range(rows):
        for c_ in range(cols):
            if q >= nq:
                break
            cx = (-1 + (2 * c_ + 1) / cols) * hx * 0.8
            cz = (-1 + (2 * r_ + 1) / rows) * hz * 0.8
            parts.append(_grid_quad(c([cx - lw / 2, hy * 0.98, cz - ld / 2]), c([0, 0, ld]), c([lw, 0, 0]), 1, 1))
            q += 1
    V, F = _merge(parts)
    is_emitter = np.zeros(len(F), bool)
    is_emitter[n_geom:] = True
    if irregular:          # a scan has no spatial face order: shuffle the faces (prim index == face order, so is_emitter moves along)
        perm = irr.permutation(len(F))
        F, is_emitter = F[perm], is_emitter[perm]
    K = int(is_emitter.sum())
    rad = np.zeros((len(F), 3), np.float32)
    rad[:K] = rng.uniform(5.0, 15.0, size=(K, 3)).astype(np.float32)      # C4: radiance U(5,15)
    return Scene(V.astype(np.float32), F.astype(np.int32), is_emitter, rad, tuple(half), seed, name)


def cornell(seed=0):
    """C1: Cornell-style closed box, walls ~10k triangles + 2 blocks, one ceiling quad (K=2)."""
    return make_room(10_000, 2, seed, (1.0, 1.0, 1.0), displaced=False, furniture=2, name="cornell-10k")


def room(n_tris=1_000_000, n_emitters=16, seed=0, irregular=False):
    """C2-C5: furnished room with displaced walls (half extents 3 x 1.4 x 2.5 m).  irregular=True: the same room with a scan-like
    tessellation (cell sizes varying ~10x, jittered vertices, random diagonals, per-vertex noise, shuffled face order) -- the
    reference's real inputs are ScanNet++ reconstructions, not regular grids."""
    nf = 12 if n_tris >= 100_000 else 4
    return make_room(n_tris, n_emitters, seed, (3.0, 1.4, 2.5), displaced=True, furniture=nf,
                     name="%sroom-%dk" % ("scan-" if irregular else "", n_tris // 1000), irregular=irregular)
