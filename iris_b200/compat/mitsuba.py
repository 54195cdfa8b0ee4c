"""Stand-in for the parts of `mitsuba` the reference's drivers touch: `set_variant`, `load_dict` for a
single-mesh scene (train_emitter.py:57-63), `math.RayEpsilon` (bake_shading.py:117), `OptixDenoiser(wh)(img)`
(bake_shading.py:81,129 -- the library's a-trous filter; parity is defined before the denoiser), `TensorXf` (refine_shading.py:124).
`load_dict` parses the OBJ / PLY mesh and builds the BVH scene the kernels traverse."""
from __future__ import annotations

import types

import numpy as np

math = types.SimpleNamespace(RayEpsilon=1500.0 * 2.0 ** -24)    # fp32 variants of Mitsuba 3.5


class BSDF:   # only so that `class FIPTBSDF(mitsuba.BSDF)` (model/fipt_bsdf.py) can be imported; relighting is out of scope
    pass


def set_variant(*names):
    return None


def variant():
    return "cuda_ad_rgb"


class _Array:
    def __init__(self, a):
        self._a = np.asarray(a)

    def numpy(self):
        return self._a

    def torch(self):
        import torch
        return torch.as_tensor(self._a)


def TensorXf(a):
    return _Array(a.detach().cpu().numpy() if hasattr(a, "detach") else a)


class OptixDenoiser:
    """`denoiser = mitsuba.OptixDenoiser(img_hw[::-1]); Ld = denoiser(Ld).numpy()` (bake_shading.py:81,129,198-199).

    OptiX's denoiser is a learned network (absent, no weights): parity with it cannot be defined, and shading-map parity is defined
    pre-denoise.  With a GPU this runs the library's own a-trous filter (iris_b200.denoise.atrous, csrc/denoise.cuh) on the map;
    `normal=` / `position=` (H,W,3 guides from the bake's primary hits) sharpen its edge stopping when the caller has them.
    `passthrough=True` (or no CUDA device) returns the input unchanged."""

    def __init__(self, input_size, albedo=False, normals=False, temporal=False, passthrough=False, **filter_args):
        self.input_size = tuple(input_size)
        self.passthrough = bool(passthrough)
        self.filter_args = filter_args

    def __call__(self, img, *a, normal=None, position=None, **k):
        arr = img.numpy() if isinstance(img, _Array) else np.asarray(img)
        import torch
        if self.passthrough or not torch.cuda.is_available() or arr.ndim != 3 or arr.shape[-1] != 3:
            return _Array(arr)
        from .. import denoise
        dev = torch.device("cuda", torch.cuda.current_device())
        g = lambda x: None if x is None else torch.as_tensor(np.asarray(x, np.float32)).to(dev)
        out = denoise.atrous(torch.as_tensor(np.ascontiguousarray(arr, np.float32)).to(dev), g(normal), g(position), **self.filter_args)
        return _Array(out.cpu().numpy())


def read_obj(path):
    vs, fs = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                vs.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):                     # fan triangulation keeps file order
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, np.float32).reshape(-1, 3), np.asarray(fs, np.int32).reshape(-1, 3)


def read_ply(path):
    with open(path, "rb") as f:
        header = []
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            header.append(line)
            if line == "end_header":
                break
        fmt = next(h.split()[1] for h in header if h.startswith("format"))
        elems, cur = [], None
        for h in header:
            t = h.split()
            if t[0] == "element":
                cur = dict(name=t[1], count=int(t[2]), props=[])
                elems.append(cur)
            elif t[0] == "property" and cur is not None:
                cur["props"].append(t[1:])
        np_t = dict(char="i1", uchar="u1", short="i2", ushort="u2", int="i4", uint="u4", float="f4", double="f8", int8="i1", uint8="u1",
                    int16="i2", uint16="u2", int32="i4", uint32="u4", float32="f4", float64="f8")
        verts = faces = None
        if fmt == "ascii":
            toks = f.read().decode("ascii").split()
            pos = 0
            for e in elems:
                if e["name"] == "vertex":
                    n = len(e["props"])
                    arr = np.asarray(toks[pos:pos + n * e["count"]], np.float64).reshape(e["count"], n)
                    pos += n * e["count"]
                    names = [p[-1] for p in e["props"]]
                    verts = arr[:, [names.index("x"), names.index("y"), names.index("z")]]
                elif e["name"] == "face":
                    out = []
                    for _ in range(e["count"]):
                        k = int(toks[pos])
                        idx = [int(t) for t in toks[pos + 1:pos + 1 + k]]
                        pos += 1 + k
                        out += [[idx[0], idx[j], idx[j + 1]] for j in range(1, k - 1)]
                    faces = np.asarray(out)
        else:
            end = "<" if "little" in fmt else ">"
            for e in elems:
                if e["name"] == "vertex":
                    dt = np.dtype([(p[-1], end + np_t[p[0]]) for p in e["props"]])
                    arr = np.frombuffer(f.read(dt.itemsize * e["count"]), dt)
                    verts = np.stack([arr["x"], arr["y"], arr["z"]], 1)
                elif e["name"] == "face":
                    p = e["props"][0]                                   # property list <count type> <index type> vertex_indices
                    ct, it = np.dtype(end + np_t[p[1]]), np.dtype(end + np_t[p[2]])
                    out = []
                    for _ in range(e["count"]):
                        k = int(np.frombuffer(f.read(ct.itemsize), ct)[0])
                        idx = np.frombuffer(f.read(it.itemsize * k), it)
                        out += [[idx[0], idx[j], idx[j + 1]] for j in range(1, k - 1)]
                    faces = np.asarray(out)
                else:
                    raise ValueError("read_ply: unsupported element %r before the faces" % e["name"])
    return np.asarray(verts, np.float32).reshape(-1, 3), np.asarray(faces, np.int32).reshape(-1, 3)


class Scene:
    """What `mitsuba.load_dict` returns here: holds the iris_b200.core.Scene (BVH in HBM) in `.iris_scene`."""

    def __init__(self, vertices, faces, device=0):
        from .. import core
        self.vertices, self.faces = vertices, faces
        self.iris_scene = core.Scene(vertices, faces, device)


def load_dict(d, device=0):
    if d.get("type") != "scene":
        raise ValueError("load_dict: only {'type': 'scene', <id>: {'type': 'obj'|'ply', 'filename': ...}} is supported")
    shapes = [v for v in d.values() if isinstance(v, dict) and v.get("type") in ("obj", "ply")]
    if len(shapes) != 1:
        raise ValueError("load_dict: exactly one obj/ply shape expected (the reference's scenes are single meshes)")
    s = shapes[0]
    v, f = (read_obj if s["type"] == "obj" else read_ply)(s["filename"])
    return Scene(v, f, device)
