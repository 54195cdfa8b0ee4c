"""CPU test of the host logic behind iris_scene_create: the binned-SAH -> 8-wide compressed BVH builder and the device
traversal code (compiled for the host with IEEE shims, tests/host/traverse_host.cpp) must reproduce the oracle's closest
hit bit for bit -- on random rays, rays leaving surfaces, rays aimed at mesh vertices and axis-parallel rays."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from iris_b200 import scenes
from oracle.intersect import OracleScene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=[0, 1], ids=["nodes80B", "nodes128B_fp16"])
def host_lib(tmp_path_factory, request):
    # both node formats of bvh8.h: the compact default and the compile-time 128-byte fp16 variant kept for A/B measurements
    out = str(tmp_path_factory.mktemp("host") / "libtraverse_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-DIRIS_NODE_FP16=%d" % request.param,
                           "-I/usr/local/cuda/include", "-o", out,
                           os.path.join(ROOT, "tests", "host", "traverse_host.cpp"), os.path.join(ROOT, "iris_b200", "csrc", "bvh_build.cpp")])
    L = ctypes.CDLL(out)
    L.host_trace.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6
    return L


def _host_trace(L, sc, o, d):
    n = len(o)
    v = np.ascontiguousarray(sc.vertices, np.float32)
    f = np.ascontiguousarray(sc.faces, np.int32)
    out = dict(t=np.empty(n, np.float32), prim=np.empty(n, np.int32), uv=np.empty((n, 2), np.float32), p=np.empty((n, 3), np.float32),
               n=np.empty((n, 3), np.float32))
    info = np.zeros(2, np.int64)
    rc = L.host_trace(v.ctypes.data, len(v), f.ctypes.data, len(f), o.ctypes.data, d.ctypes.data, n, out["t"].ctypes.data,
                      out["prim"].ctypes.data, out["uv"].ctypes.data, out["p"].ctypes.data, out["n"].ctypes.data, info.ctypes.data)
    assert rc == 0
    return out, info


@pytest.mark.parametrize("which", ["cornell", "room60k", "scanroom60k"])
def test_bvh8_traversal_matches_oracle(host_lib, which):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import _rays_for_parity
    # scanroom: the irregular, scan-like tessellation (varying cell sizes, jittered vertices, random diagonals, shuffled faces)
    sc = scenes.cornell() if which == "cornell" else scenes.room(60_000, 16, seed=3, irregular=(which == "scanroom60k"))
    osc = OracleScene(sc.vertices, sc.faces)
    o, d = _rays_for_parity(sc, osc, 40_000, 1)
    o, d = np.ascontiguousarray(o), np.ascontiguousarray(d)
    ref = osc.intersect_raw(o, d, "bvh")
    got, info = _host_trace(host_lib, sc, o, d)
    assert 0 < info[1] <= 32 and info[0] < sc.n_tris
    for k in ("prim", "t", "uv", "p", "n"):
        assert np.array_equal(got[k], ref[k]), k
    if which == "cornell":                       # and both agree with the brute-force loop
        sub = slice(0, len(o), 7)
        br = osc.intersect_raw(o[sub], d[sub], "brute")
        assert np.array_equal(br["prim"], got["prim"][sub]) and np.array_equal(br["t"], got["t"][sub])
