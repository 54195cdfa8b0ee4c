#!/usr/bin/env python
"""bench.py -- path samples / second, forward + backward, for the differentiable shading hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A *step* = one pass of `path_tracing_single` forward + adjoint over one batch of synthetic camera rays:
workload "c3" (default; BASELINE.json configs[2]): 1M-triangle room, 8 views 1280x960 per GPU, SPP=256 as 8 chunks of spp=32,
random-init hash-grid+MLP BRDF field, K=16 emitter triangles, MIS on, MSE loss, gradients to the field parameters (27.96M) and
to emitter.radiance ("c4" = configs[3]: emitter-radiance gradient only, the train_emitter.py step).  The scene/BVH, SLF and BRDF
field are replicated, pixels are sharded by view, and the only collective is the allreduce of the flat gradient buffer.
A *path sample* = one (pixel, spp index) lane through the whole estimator (3 ray casts, BRDF-field evaluation, emitter / SLF
lookups, MIS, adjoint replay).

Prints ONE JSON line (rank 0).  `value` times the steps with inputs resident in HBM; `e2e` times the same steps through the
public API with the rays in pinned HOST memory (H2D inside the timed region, loss + gradient read back).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: tris, emitters, views per GPU, width, height, SPP, spp, estimator
    "c3": dict(tris=1_000_000, emitters=16, views=8, width=1280, height=960, SPP=256, spp=32, brdf_grad=True,
               desc="training step: path_tracing_single -> EmorCRF -> MSE, fwd+bwd with gradients to the BRDF field (hash grid + MLP), the emitter radiance and the CRF weight, 1M-tri room, 8 views 1280x960/GPU, SPP=256 (8 x spp 32), random-init field, K=16, MIS on"),
    "c4": dict(tris=1_000_000, emitters=16, views=8, width=1280, height=960, SPP=256, spp=32, brdf_grad=False,
               desc="train_emitter step: path_tracing_single -> EmorCRF (fixed) -> MSE, fwd+bwd, emitter-radiance gradient, 1M-tri room, 8 views 1280x960/GPU, SPP=256 (8 x spp 32), K=16, MIS on"),
    "c5": dict(tris=5_000_000, emitters=16, views=64, width=1920, height=1440, SPP=128, spp=128, brdf_grad=True, strong=True,
               desc="scaling sweep: 5M-tri room, 64 views 1920x1440 in total (sharded over the GPUs), spp=128, path_tracing_single fwd+bwd, field + emitter gradients"),
    "c2": dict(tris=1_000_000, emitters=16, views=1, width=640, height=480, SPP=64, spp=64, brdf_grad=False, bake=True,
               desc="shading-map bake (bake_shading.py): 1M-tri room, 640x480, spp=64, diffuse map + 6 roughness levels x 2 Fresnel maps, forward only"),
    "brdf": dict(tris=1_000_000, emitters=16, views=8, width=1280, height=960, SPP=1, spp=1, brdf_grad=True, maps=True,
                 desc="train_brdf_crf step (SURVEY 8f-2): NGPBRDF field at the primary hits of 8 views 1280x960 -> kd/ks shading from baked maps (6 roughness levels) -> EmorCRF -> MSE + diffuse regulariser, backward to field params and CRF weight"),
    "c1": dict(tris=10_000, emitters=2, views=1, width=64, height=64, SPP=16, spp=16, brdf_grad=True,
               desc="Cornell ~10k tris, 64x64, spp=16, path_tracing_single fwd+bwd (reference's CPU-runnable case)"),
}


def ray_bytes(n_tris):
    """one root-to-leaf path of an 8-wide 80-byte BVH + one 48-byte triangle."""
    return math.ceil(math.log(max(n_tris, 8) / 4.0, 8)) * 80 + 48


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--views", type=int, default=None)
    ap.add_argument("--tile", type=int, default=262144, help="pixels per forward/backward tile")
    ap.add_argument("--cpu-sample-pixels", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_leg(w, sc, n_pixels, steps, warmup):
    """The reference's own algorithm for this path on the host cores: oracle/estimators.path_tracing_single forward + autograd
    backward to emitter radiance, on the first n_pixels pixels of view 0 with spp = w['spp'] (one chunk)."""
    import torch
    from oracle import estimators as E
    from oracle import field as OF
    from oracle.intersect import OracleScene
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    osc = OracleScene(sc.vertices, sc.faces)
    em = E.Emitter(sc.emitter_dict(), sc.slf_dict(256), learn=True)
    params = bench_params().requires_grad_(bool(w.get("brdf_grad")))
    vmin, vmax = sc.voxel_bounds()
    mat_fn = lambda x: OF.material(x, params, vmin, vmax)
    rays = torch.as_tensor(sc.camera_rays(w["width"], w["height"], view=1))
    stride = max(1, len(rays) // n_pixels)
    r = rays[::stride][:n_pixels]
    spp = w["spp"]
    rng = np.random.default_rng(0)
    times = []
    for it in range(warmup + steps):
        U = torch.as_tensor(np.minimum(rng.random((len(r) * spp, 8), dtype=np.float32), np.float32(0.99999)))
        t0 = time.perf_counter()
        L = E.path_tracing_single(osc, em, mat_fn, r[:, 0:3], r[:, 3:6], r[:, 6:9], r[:, 9:12], spp, U)
        loss = ((L - 0.5) ** 2).mean()
        em.radiance.grad = None
        params.grad = None
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per = float(np.mean(times))
    return dict(value=len(r) * spp / per, unit="samples/s", cores=cores, kind="port",
                sample="%d pixels (every %d-th of view 1) x spp %d, one chunk, fwd+bwd to %s" % (len(r), stride, spp, "field params + emitter.radiance" if w.get("brdf_grad") else "emitter.radiance")), per


def bench_params():
    """Random-init BRDF field: tcnn-style grid U(+-1e-4) + Xavier MLP."""
    import torch
    g = torch.Generator().manual_seed(0)
    p = torch.empty(9216 + 27954112)
    b = math.sqrt(6.0 / 128)
    p[:8192] = (torch.rand(8192, generator=g) * 2 - 1) * b
    p[8192:9216] = (torch.rand(1024, generator=g) * 2 - 1) * math.sqrt(6.0 / 80)
    p[9216:] = (torch.rand(27954112, generator=g) * 2 - 1) * 1e-4
    return p


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ bake (configs[1], forward only)
def bench_bake(a, w, sc, scene, tables, dev, config, stats):
    """One step = the bake of one 640x480 training view: primary hits (ray_intersect), then per pixel spp secondary rays for the
    diffuse map and for each of the 6 roughness levels (two Fresnel maps each) -- 7 fused launches (bake_shading.py:93-204)."""
    import torch
    from iris_b200 import core
    lib = core.C.lib()
    rays = torch.as_tensor(sc.camera_rays(w["width"], w["height"], view=1)).to(dev)
    spp = w["spp"]
    levels = [0.02 + 0.98 * i / 5 for i in range(6)]

    def step(s):
        t, prim, uv, p, n = scene.intersect_raw(rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous())
        valid = prim >= 0
        pos, nrm, wo = p[valid].contiguous(), n[valid].contiguous(), (-rays[:, 3:6])[valid].contiguous()
        outs = [core.bake(scene, tables, 0, 1.0, pos, nrm, None, spp, core.Sampler(seed=10 + s))]
        for i, r in enumerate(levels):
            outs += list(core.bake(scene, tables, 1, r, pos, nrm, wo, spp, core.Sampler(seed=100 * (i + 1) + s)))
        return pos.shape[0], outs

    for s in range(max(a.warmup, 3)):
        npx, _ = step(s)
    torch.cuda.synchronize()
    l0 = lib.iris_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        npx, outs = step(100 + s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_rays = (npx * spp * 7 + rays.shape[0]) * a.steps
    peak = 6518.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    bpr = ray_bytes(sc.n_tris) + 16 + 36.0 / spp
    value = n_rays / (ms * 1e-3)
    line = dict(metric="shading_map_bake_rays_per_sec", value=value, unit="rays/s", n_gpus=1, steps=a.steps, warmup=max(a.warmup, 3), ms_per_step=ms / a.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=config,
                gpu_launches=int(lib.iris_launch_count() - l0),
                roofline=dict(bound="hbm", kernel="k_bake_persistent", achieved=value * bpr / 1e9, peak=peak, unit="GB/s", frac=value * bpr / 1e9 / peak, traffic=None,
                              algorithmic_bytes_per_ray=bpr),
                scene=dict(bvh_nodes=stats["n_nodes"], bvh_build_ms=stats["build_ms"], bvh_depth=stats["max_depth"]),
                maps_finite=bool(all(torch.isfinite(o).all() for o in outs)))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ train_brdf_crf step (8f-2)
def bench_brdf(a, w, sc, scene, dev, config, stats):
    """One step = the shading block of train_brdf_crf.py:176-211 over every valid pixel of the views: ray_intersect, field forward,
    kd/ks shading from the baked maps, EmorCRF, MSE + loss_d, and the adjoint down to mlp.params and the CRF weight."""
    import torch
    from iris_b200 import core, ops
    from iris_b200.crf import EmorCRF
    lib = core.C.lib()
    g = torch.Generator().manual_seed(0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp = torch.nn.Module()
            self.mlp.params = torch.nn.Parameter(bench_params().to(dev))
            self.voxel_min, self.voxel_max = sc.voxel_bounds()
    net = Net()
    x = np.linspace(0.0, 1.0, 1024, dtype=np.float32)
    crf = EmorCRF(dim=11, tables=(x ** (1 / 2.2), np.stack([np.sin((k + 1) * np.pi * x) * 0.05 for k in range(11)]).astype(np.float32))).to(dev)
    views = []
    for v in range(w["views"]):
        rays = torch.as_tensor(sc.camera_rays(w["width"], w["height"], view=v + 1)).to(dev)
        n = rays.shape[0]
        views.append(dict(o=rays[:, 0:3].contiguous(), d=rays[:, 3:6].contiguous(), diffuse=torch.rand(n, 3, generator=g).to(dev),
                          spec0=torch.rand(n, 6, 3, generator=g).to(dev), spec1=(torch.rand(n, 6, 3, generator=g) * 0.2).to(dev),
                          rgb=torch.rand(n, 3, generator=g).to(dev)))
    exposure = torch.ones(1, device=dev)

    def step():
        net.mlp.params.grad = None
        crf.weight.grad = None
        n_valid, total = 0, 0.0
        for V in views:
            t, prim, uv, p, nrm = scene.intersect_raw(V["o"], V["d"])
            valid = prim >= 0
            L, mat = ops.brdf_shading(net, p[valid], V["diffuse"][valid], V["spec0"][valid], V["spec1"][valid])
            ldr = crf(L, exposure)
            loss = torch.nn.functional.mse_loss(ldr, V["rgb"][valid]) + 5e-4 * ((mat["roughness"] - 1).abs().mean() + mat["metallic"].mean())
            loss.backward()
            n_valid += int(L.shape[0])
            total += float(loss.detach())
        return n_valid, total

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = lib.iris_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        n_valid, loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    value = n_valid * a.steps / (ms * 1e-3)
    peak = 6518.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    bpp = ray_bytes(sc.n_tris) + 2048 + 176 + 12 + 24 + 20      # primary cast + field fwd+bwd gathers/scatters + maps + L + rgb + mat
    line = dict(metric="brdf_crf_step_pixels_per_sec", value=value, unit="pixels/s", n_gpus=1, steps=a.steps, warmup=max(a.warmup, 3), ms_per_step=ms / a.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=config, gpu_launches=int(lib.iris_launch_count() - l0),
                roofline=dict(bound="hbm", kernel="step", achieved=value * bpp / 1e9, peak=peak, unit="GB/s", frac=value * bpp / 1e9 / peak, traffic=None,
                              algorithmic_bytes_per_pixel=bpp),
                scene=dict(bvh_nodes=stats["n_nodes"], bvh_build_ms=stats["build_ms"], bvh_depth=stats["max_depth"]),
                loss=loss, d_params_abs_sum=float(net.mlp.params.grad.abs().sum()), d_crf_weight_abs_sum=float(crf.weight.grad.abs().sum()))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ main
def main():
    a = parse()
    w = dict(WORKLOADS[a.workload])
    if a.views:
        w["views"] = a.views
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    from iris_b200 import scenes
    n_chunks = w["SPP"] // w["spp"]
    pix_per_view = w["width"] * w["height"]
    config = dict(workload=a.workload, description=w["desc"], triangles=None, views_per_gpu=w["views"], width=w["width"], height=w["height"],
                  SPP=w["SPP"], spp=w["spp"], emitters=w["emitters"], slf_H=256, sharding="weak scaling: N x views_per_gpu views in total, each rank takes pixel band rank/N of every view (all spp of a pixel on one rank); scene/SLF/field replicated; one allreduce of the gradient buffer per step",
                  l2="inputs larger than L2 (BVH+triangles 62 MB, SLF 64 MB+, hash grid 56 MB, per-tile records > 400 MB)")

    if a.impl == "reference":
        if rank != 0:
            return
        sc = scenes.cornell() if a.workload == "c1" else scenes.room(w["tris"], w["emitters"], seed=0)
        config["triangles"] = sc.n_tris
        cb, per = cpu_leg(w, sc, a.cpu_sample_pixels, a.steps, a.warmup)
        line = dict(metric="path_samples_per_sec_fwd_bwd", value=cb["value"], unit="samples/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=per * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=config,
                    impl="reference", cpu_baseline=cb, e2e=dict(value=cb["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return

    import torch
    from iris_b200 import core
    from iris_b200 import dist as idist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    sc = scenes.cornell() if a.workload == "c1" else scenes.room(w["tris"], w["emitters"], seed=0)
    config["triangles"] = sc.n_tris
    scene = core.Scene(sc.vertices, sc.faces, local)
    stats = scene.stats()
    tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), bench_params(), sc.voxel_bounds())
    if w.get("bake"):
        return bench_bake(a, w, sc, scene, tables, dev, config, stats)
    if w.get("maps"):
        return bench_brdf(a, w, sc, scene, dev, config, stats)
    if w.get("strong"):                                   # fixed total work, views sharded over the ranks
        lo_v, hi_v = idist.shard_range(w["views"], rank, world)
        views = [v + 1 for v in range(lo_v, hi_v)]
    else:
        views = [v + 1 for v in range(world * w["views"])]   # weak scaling: world x views in total ...
    per_view = []
    for v in views:
        r = torch.as_tensor(sc.camera_rays(w["width"], w["height"], view=v))
        if not w.get("strong") and world > 1:                # ... and every rank takes the same pixel band of EVERY view, so that the
            lo_p, hi_p = idist.shard_range(r.shape[0], rank, world)   # ranks' work is balanced whatever the views cost
            r = r[lo_p:hi_p].clone()
        per_view.append(r)
    rays_host = torch.cat(per_view).pin_memory()
    del per_view
    rays_dev = rays_host.to(dev)
    P = rays_host.shape[0]
    spp = w["spp"]
    tile = min(a.tile, P)
    lib = core.C.lib()
    ws = torch.empty(lib.iris_single_workspace_bytes(tile, spp), dtype=torch.uint8, device=dev)   # forward streams | d_mat + activation streams
    recs = [torch.empty(lib.iris_single_record_bytes(tile, spp), dtype=torch.uint8, device=dev) for _ in range(n_chunks)]
    want_par = bool(w.get("brdf_grad"))
    # encoded field inputs of every sample, kept from forward to adjoint only when the field gradient is wanted
    encs = [torch.empty(lib.iris_single_encoded_bytes(tile, spp), dtype=torch.uint8, device=dev) if want_par else None for _ in range(n_chunks)]
    target = torch.full((P, 3), 0.5, device=dev)
    n_samples_rank = P * w["SPP"]
    step_no = [0]

    want_par = bool(w.get("brdf_grad"))
    d_par_buf = torch.zeros(9216 + 27954112, device=dev) if want_par else None

    from iris_b200.crf import EmorCRF
    xs = np.linspace(0.0, 1.0, 1024, dtype=np.float32)     # synthetic response: gamma 2.2 mean curve + 11 smooth basis functions
    crf = EmorCRF(dim=11, tables=(xs ** (1 / 2.2), np.stack([np.sin((k + 1) * np.pi * xs) * 0.05 for k in range(11)]).astype(np.float32))).to(dev)
    crf.weight.requires_grad_(want_par)                      # train_emitter keeps the response fixed
    exposure = torch.ones(1, device=dev)

    def step(host_inputs):
        """One training step of this rank: returns (loss tensor, d_radiance)."""
        crf.weight.grad = None
        d_rad = torch.zeros(tables.K, 3, device=dev)
        if want_par:
            d_par_buf.zero_()
        loss = torch.zeros((), device=dev)
        s = step_no[0]
        step_no[0] += 1
        for t0 in range(0, P, tile):
            t1 = min(t0 + tile, P)
            rays = rays_host[t0:t1].to(dev, non_blocking=True) if host_inputs else rays_dev[t0:t1]
            L = torch.zeros(t1 - t0, 3, device=dev)
            for c in range(n_chunks):
                smp = core.Sampler(seed=1000 + s, lane_offset=(t0 * n_chunks + c * (t1 - t0)) * spp)
                Lc = torch.empty(t1 - t0, 3, device=dev)
                P_, S_ = tables.c(), smp.c()
                core.C.check(lib.iris_single_forward(scene.handle, P_, core.C.ptr(rays), t1 - t0, spp, S_, core.C.ptr(Lc), core.C.ptr(recs[c]),
                                                     core.C.ptr(encs[c]), core.C.ptr(ws), ws.numel(), core.C.stream_ptr()))
                L += Lc
            L /= n_chunks
            # camera response + MSE, as every trainer does right after the estimator (train_emitter.py:191-193, train_brdf_crf.py:208-209):
            # EmorCRF forward / adjoint kernels, gradient to L and -- in the train_brdf_crf configuration -- to the CRF weight
            L.requires_grad_(True)
            ldr = crf(L, exposure)
            loss_t = ((ldr - target[t0:t1]) ** 2).sum() / (P * 3 * world)
            loss_t.backward()
            loss += loss_t.detach()
            dL = L.grad / n_chunks
            for c in range(n_chunks):
                P_ = tables.c()
                core.C.check(lib.iris_single_backward(P_, core.C.ptr(dL), t1 - t0, spp, core.C.ptr(recs[c]), core.C.ptr(encs[c]), core.C.ptr(d_rad), core.C.ptr(d_par_buf),
                                                      core.C.ptr(ws) if want_par else None, ws.numel() if want_par else 0, core.C.stream_ptr()))
        if dist is not None:
            idist.allreduce_gradients([d_rad, d_par_buf, crf.weight.grad])
            dist.all_reduce(loss)
        if host_inputs:
            # the optimiser consumes gradients on the device; what leaves the GPU per step is the loss, the K x 3 emitter gradient
            # and (to make the field gradient observable) its L1 norm
            return loss.cpu(), d_rad.cpu(), (d_par_buf.abs().sum().cpu() if want_par else None)
        return loss, d_rad, (d_par_buf.abs().sum() if want_par else None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host_inputs, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = step(host_inputs)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    # ---- warm-up, then the timed region (inputs resident in HBM)
    for _ in range(max(a.warmup, 3)):
        step(False)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    lib.iris_profile_enable(1)
    n_kernels = 0
    while lib.iris_profile_name(n_kernels):
        n_kernels += 1
    for k in range(n_kernels):
        lib.iris_profile_read(k, None, None, 1)
    l0 = lib.iris_launch_count()
    ms, (loss, d_rad, d_par_l1) = timed(False, a.steps)
    launches = lib.iris_launch_count() - l0
    prof = {}
    for k in range(n_kernels):
        n_, t_ = core.C.c_i64(), core.C.ctypes.c_double()
        lib.iris_profile_read(k, core.C.ctypes.byref(n_), core.C.ctypes.byref(t_), 1)
        if n_.value:
            prof[lib.iris_profile_name(k).decode()] = dict(launches=n_.value, total_ms=t_.value)
    lib.iris_profile_enable(0)
    clk = clocks.stop() if rank == 0 else None
    value = n_samples_rank * world * a.steps / (ms * 1e-3)

    # ---- the same steps end to end: rays in pinned host memory, loss + gradient read back
    e2e = None
    if not a.no_e2e:
        step(True)
        ms_e, _ = timed(True, a.steps)
        e2e = dict(value=n_samples_rank * world * a.steps / (ms_e * 1e-3), unit="samples/s", h2d_bytes_per_step=int(P * 12 * 4 * world),
                   d2h_bytes_per_step=int((tables.K * 3 + 2) * 4 * world), ms_per_step=ms_e / a.steps)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: k_trace_queue (2 of the 3 ray casts of a sample, through the ray queue); with the
    #      fused bounce selected (iris_set_option single_impl 0) it is k_bounce_single (the same casts + BSDF/emitter/SLF/MIS + record)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    R = ray_bytes(sc.n_tris)
    est_bytes = 3 * R + (2048 if want_par else 1024) + 16 + 72.0 / spp      # `single` estimator (+1024 B of grid-gradient writes with BRDF grads)
    kname = "k_trace_queue" if "k_trace_queue" in prof else "k_bounce_single"
    kb = prof.get(kname)
    roof = None
    if kb:
        # this kernel's share per sample.  fused: 2 casts + SLF + pixel out/dL.  queue: 2 casts + 2 x (32 B ray in + 16 B hit out)
        k_bytes = 2 * R + 96 if kname == "k_trace_queue" else 2 * R + 16 + 24.0 / spp
        per_launch_samples = n_samples_rank * a.steps / kb["launches"]
        avg_ms = kb["total_ms"] / kb["launches"]
        ach = per_launch_samples * k_bytes / (avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname)
        except Exception:
            pass
        roof = dict(bound="hbm", kernel=kname, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                    peak_source="measured (MEASURED_PEAKS.json)" if peaks else "fallback", algorithmic_bytes_per_sample=k_bytes,
                    samples_per_launch=per_launch_samples, avg_launch_ms=avg_ms,
                    step=dict(algorithmic_bytes_per_sample=est_bytes, achieved=value / max(world, 1) * est_bytes / 1e9, frac=value / max(world, 1) * est_bytes / 1e9 / peak),
                    kernel_share_of_step={k: v["total_ms"] / ms for k, v in prof.items()})

    cb = None
    if not a.no_cpu_baseline and world == 1:
        cb, _ = cpu_leg(w, sc, a.cpu_sample_pixels, 1, 1)

    line = dict(metric="path_samples_per_sec_fwd_bwd", value=value, unit="samples/s", n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
                ms_per_step=ms / a.steps, higher_is_better=True, scaling="strong" if w.get("strong") else "weak", vs_baseline=None, dtype="f32", data="synthetic", config=config,
                clocks=clk, e2e=e2e, gpu_launches=int(launches), roofline=roof, cpu_baseline=cb,
                scene=dict(bvh_nodes=stats["n_nodes"], bvh_build_ms=stats["build_ms"], bvh_depth=stats["max_depth"]),
                loss=float(loss), d_radiance_abs_sum=float(d_rad.abs().sum()), d_params_abs_sum=(float(d_par_l1) if d_par_l1 is not None else None),
                d_crf_weight_abs_sum=(float(crf.weight.grad.abs().sum()) if crf.weight.grad is not None else None))
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
