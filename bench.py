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
public API with the rays in pinned HOST memory (H2D inside the timed region, loss + gradient read back).  The default (c3) line
also carries `extra`: short device-timed runs (1 warm-up + 2 steps) of the other BASELINE configurations -- c4, c2, the per-GPU
shard of c5 (8 views of the 5M-triangle room at 1920x1440, spp 128), the c3 step on a scan-like irregular mesh, path_tracing with 5 indirect bounces, and the per-call latency
at the trainers' real batch shape (8192 pixels x spp 32, eager and CUDA graph) -- so that they are measured by whoever runs this file.
`--workload c5` is the full strong-scaling configuration (64 views dealt to the ranks).
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
    ap.add_argument("--no-extras", action="store_true", help="default workload only: skip the short c4 / c2 / c5-shard / path_tracing / latency measurements")
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


# ------------------------------------------------------------------------------------------------ estimator workloads (c1, c3, c4, c5)
class Ctx:
    """Process context of one bench rank."""
    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        self.scenes = {}

    def scene(self, w):
        """(procedural scene, device Scene, stats, tables) -- built once per triangle count."""
        import torch
        from iris_b200 import core, scenes
        key = "c1" if w.get("cornell") else (w["tris"], bool(w.get("irregular")))
        if not self.scenes and not getattr(self, "_warm", False):
            # context creation and the lazy loading of the builder's kernels are not part of a BVH build: spend them on a toy scene, so that
            # scene.bvh_build_ms below is the build (iris_scene_create's own wall clock, allocations included)
            torch.zeros(1, device=self.dev)
            toy = scenes.cornell()
            core.Scene(toy.vertices, toy.faces, self.local)
            torch.cuda.synchronize()
            self._warm = True
        if key not in self.scenes:
            sc = scenes.cornell() if w.get("cornell") else scenes.room(w["tris"], w["emitters"], seed=0, irregular=bool(w.get("irregular")))
            scene = core.Scene(sc.vertices, sc.faces, self.local, builder=os.environ.get("IRIS_BENCH_BUILDER", "auto"))   # A/B only: "sah" / "lbvh"
            tables = core.ShadingTables.from_dicts(self.dev, sc.emitter_dict(), sc.slf_dict(256), bench_params(), sc.voxel_bounds())
            self.scenes[key] = (sc, scene, scene.stats(), tables)
        return self.scenes[key]

    def drop_scenes(self):
        import torch
        self.scenes.clear()
        torch.cuda.empty_cache()


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_estimator(ctx, a, w, steps, warmup, want_e2e, want_prof, tile):
    """path_tracing_single -> EmorCRF -> MSE, forward + adjoint, over this rank's pixel shard; returns a dict of measurements
    (value = samples/s over all ranks, device-timed, max over ranks)."""
    import torch
    from iris_b200 import core
    from iris_b200 import dist as idist
    from iris_b200.crf import EmorCRF
    dist, dev, rank, world = ctx.dist, ctx.dev, ctx.rank, ctx.world
    sc, scene, stats, tables = ctx.scene(w)
    lib = core.C.lib()
    n_chunks = w["SPP"] // w["spp"]
    if w.get("strong"):                                   # fixed total work (BASELINE configs[4]): the views are dealt to the ranks
        views = [v + 1 for v in range(w["views"])][rank::world]
    else:
        views = [v + 1 for v in range(world * w["views"])]   # weak scaling: world x views in total ...
    per_view = []
    for v in views:
        r = torch.as_tensor(sc.camera_rays(w["width"], w["height"], view=v))
        if not w.get("strong") and world > 1:                # ... every rank takes blocks of 256 consecutive pixels of EVERY view round-robin,
            r = r[idist.interleaved_rows(r.shape[0], 256, rank, world)].clone()   # so the ranks' loads are equal whatever a region costs
        per_view.append(r)
    rays_host = torch.cat(per_view).pin_memory()
    del per_view
    rays_dev = rays_host.to(dev)
    P = rays_host.shape[0]
    spp = w["spp"]
    tile = min(tile, P)
    want_par = bool(w.get("brdf_grad"))
    ws = torch.empty(lib.iris_single_workspace_bytes(tile, spp), dtype=torch.uint8, device=dev)   # forward streams | d_mat + activation streams
    recs = [torch.empty(lib.iris_single_record_bytes(tile, spp), dtype=torch.uint8, device=dev) for _ in range(n_chunks)]
    # encoded field inputs of every sample, kept from forward to adjoint only when the field gradient is wanted
    encs = [torch.empty(lib.iris_single_encoded_bytes(tile, spp), dtype=torch.uint8, device=dev) if want_par else None for _ in range(n_chunks)]
    target = torch.full((P, 3), 0.5, device=dev)
    P_total = torch.tensor([P], device=dev, dtype=torch.int64)
    if dist is not None:
        dist.all_reduce(P_total)
    P_total = int(P_total.item())
    n_samples_all = P_total * w["SPP"]
    step_no = [0]
    d_par_buf = torch.zeros(9216 + 27954112, device=dev) if want_par else None
    xs = np.linspace(0.0, 1.0, 1024, dtype=np.float32)     # synthetic response: gamma 2.2 mean curve + 11 smooth basis functions
    crf = EmorCRF(dim=11, tables=(xs ** (1 / 2.2), np.stack([np.sin((k + 1) * np.pi * xs) * 0.05 for k in range(11)]).astype(np.float32))).to(dev)
    crf.weight.requires_grad_(want_par)                      # train_emitter keeps the response fixed
    exposure = torch.ones(1, device=dev)

    def step(host_inputs):
        """One training step of this rank: returns (loss tensor, d_radiance, |d_params|_1)."""
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
                smp = core.Sampler(seed=1000 + s, lane_offset=((rank * (1 << 40)) + t0 * n_chunks + c * (t1 - t0)) * spp)
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
            loss_t = ((ldr - target[t0:t1]) ** 2).sum() / (P_total * 3)
            loss_t.backward()
            loss += loss_t.detach()
            dL = L.grad / n_chunks
            for c in range(n_chunks):
                P_ = tables.c()
                core.C.check(lib.iris_single_backward(P_, core.C.ptr(dL), t1 - t0, spp, core.C.ptr(recs[c]), core.C.ptr(encs[c]), core.C.ptr(d_rad), core.C.ptr(d_par_buf),
                                                      core.C.ptr(ws) if want_par else None, ws.numel() if want_par else 0, core.C.stream_ptr()))
        if dist is not None:
            idist.allreduce_gradients([d_rad, d_par_buf, crf.weight.grad, loss])
        if host_inputs:
            # the optimiser consumes gradients on the device; what leaves the GPU per step is the loss, the K x 3 emitter gradient
            # and (to make the field gradient observable) its L1 norm
            return loss.cpu(), d_rad.cpu(), (d_par_buf.abs().sum().cpu() if want_par else None)
        return loss, d_rad, (d_par_buf.abs().sum() if want_par else None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host_inputs, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = step(host_inputs)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(warmup):
        step(False)
    clocks = ClockSampler(ctx.local)
    if rank == 0 and want_prof:
        clocks.start()
    prof, n_kernels = {}, 0
    if want_prof:
        lib.iris_profile_enable(1)
        while lib.iris_profile_name(n_kernels):
            n_kernels += 1
        for k in range(n_kernels):
            lib.iris_profile_read(k, None, None, 1)
    l0 = lib.iris_launch_count()
    ms, (loss, d_rad, d_par_l1) = timed(False, steps)
    launches = lib.iris_launch_count() - l0
    for k in range(n_kernels):
        n_, t_ = core.C.c_i64(), core.C.ctypes.c_double()
        lib.iris_profile_read(k, core.C.ctypes.byref(n_), core.C.ctypes.byref(t_), 1)
        if n_.value:
            prof[lib.iris_profile_name(k).decode()] = dict(launches=n_.value, total_ms=t_.value)
    if want_prof:
        lib.iris_profile_enable(0)
    clk = clocks.stop() if (rank == 0 and want_prof) else None
    value = n_samples_all * steps / (ms * 1e-3)
    e2e = None
    if want_e2e:        # the same steps end to end: rays in pinned host memory, loss + gradient read back
        step(True)
        ms_e, _ = timed(True, steps)
        h2d = torch.tensor([P * 12 * 4], device=dev, dtype=torch.int64)
        if dist is not None:
            dist.all_reduce(h2d)
        e2e = dict(value=n_samples_all * steps / (ms_e * 1e-3), unit="samples/s", h2d_bytes_per_step=int(h2d.item()),
                   d2h_bytes_per_step=int((tables.K * 3 + 2) * 4 * world), ms_per_step=ms_e / steps)
    R = ray_bytes(sc.n_tris)
    est_bytes = 3 * R + (2048 if want_par else 1024) + 16 + 72.0 / spp      # `single` estimator (+1024 B of grid-gradient writes with BRDF grads)
    out = dict(value=value, ms=ms, steps=steps, launches=int(launches), prof=prof, clocks=clk, e2e=e2e, loss=float(loss), n_samples_rank=P * w["SPP"],
               d_radiance_abs_sum=float(d_rad.abs().sum()), d_params_abs_sum=(float(d_par_l1) if d_par_l1 is not None else None),
               d_crf_weight_abs_sum=(float(crf.weight.grad.abs().sum()) if crf.weight.grad is not None else None),
               est_bytes=est_bytes, R=R, stats=stats, triangles=sc.n_tris, pixels_rank=P, pixels_total=P_total)
    del recs, encs, ws, rays_dev, rays_host, target, d_par_buf
    torch.cuda.empty_cache()
    return out


def run_path_tracing(ctx, w, steps, warmup):
    """path_tracing(indir_depth=5) forward (render.py:171-176) on one 1280x960 view per rank, spp 16: samples/s (12 ray casts and up to
    7 field evaluations per sample on shrinking lane sets)."""
    import torch
    from iris_b200 import core
    sc, scene, stats, tables = ctx.scene(w)
    dev = ctx.dev
    rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1 + ctx.rank)).to(dev)
    spp, depth = 16, 5
    wsb = torch.empty(core.C.lib().iris_wave_workspace_bytes(rays.shape[0] * spp), dtype=torch.uint8, device=dev)
    for s in range(warmup):
        core.path_tracing(scene, tables, rays, spp, depth, core.Sampler(seed=s), wsb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        L = core.path_tracing(scene, tables, rays, spp, depth, core.Sampler(seed=10 + s), wsb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(metric="path_tracing_depth5_samples_per_sec_fwd", value=rays.shape[0] * spp * steps / (ms * 1e-3) * ctx.world, unit="samples/s", ms_per_step=ms / steps,
                config="1280x960 x spp 16, indir_depth 5, one view per GPU", finite=bool(torch.isfinite(L).all()))


def run_latency(ctx, w, steps):
    """The reference trainers' real call shape (configs/config.py:9-12, train_emitter.py:184-189): B = 8192 pixels x spp 32 through the
    Python operator layer (ops.path_tracing_single + autograd backward to emitter.radiance), per call: eager, and with the forward
    captured in a CUDA graph (fixed sample seed; workspace and record reused)."""
    import torch
    from iris_b200 import core, ops
    sc, scene, stats, tables = ctx.scene(w)
    dev = ctx.dev
    B, spp = 8192, 32
    rays = torch.as_tensor(sc.camera_rays(1280, 960, view=1))
    rays = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(0))[:B]].to(dev)       # a random pixel batch, like the trainers'

    lib = core.C.lib()

    def eager():
        L, rec = core.single_forward(scene, tables, rays, spp, core.Sampler(seed=3), True, want_encoded=False, workspace=eager.ws)
        return core.single_backward(tables, eager.g, spp, rec)
    eager.ws = torch.empty(lib.iris_single_workspace_bytes(B, spp), dtype=torch.uint8, device=dev)
    eager.g = torch.randn(B, 3, device=dev)
    for _ in range(5):
        eager()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        eager()
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) * 1e3 / steps
    # forward + adjoint captured once, replayed
    graph_ms = None
    try:
        L = torch.empty(B, 3, device=dev)
        rec = torch.empty(lib.iris_single_record_bytes(B, spp), dtype=torch.uint8, device=dev)
        d_rad = torch.zeros(tables.K, 3, device=dev)
        P_, S_ = tables.c(), core.Sampler(seed=3).c()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            def body():
                core.C.check(lib.iris_single_forward(scene.handle, P_, core.C.ptr(rays), B, spp, S_, core.C.ptr(L), core.C.ptr(rec), None,
                                                     core.C.ptr(eager.ws), eager.ws.numel(), core.C.stream_ptr()))
                d_rad.zero_()
                core.C.check(lib.iris_single_backward(P_, core.C.ptr(eager.g), B, spp, core.C.ptr(rec), None, core.C.ptr(d_rad), None, None, 0, core.C.stream_ptr()))
            body()
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                body()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            g.replay()
        torch.cuda.synchronize()
        graph_ms = (time.perf_counter() - t0) * 1e3 / steps
    except Exception as e:                                 # a capture failure must not cost the headline
        graph_ms = "capture failed: %s" % (str(e)[:120],)
    n = B * spp
    return dict(metric="train_emitter_call_latency", config="B=8192 x spp 32, path_tracing_single fwd + adjoint to emitter.radiance, wall clock per call incl. host launch overhead",
                eager_ms=eager_ms, eager_samples_per_s=n / (eager_ms * 1e-3), graph_ms=graph_ms,
                graph_samples_per_s=(n / (graph_ms * 1e-3) if isinstance(graph_ms, float) else None))


def run_bake(ctx, w, steps, warmup):
    """One step = the bake of one 640x480 training view (bake_shading.py:93-204): primary hits, then spp secondary rays per pixel for the
    diffuse map and for each of the 6 roughness levels (two Fresnel maps each) -- 7 persistent launches."""
    import torch
    from iris_b200 import core
    sc, scene, stats, tables = ctx.scene(w)
    dev = ctx.dev
    lib = core.C.lib()
    rays = torch.as_tensor(sc.camera_rays(w["width"], w["height"], view=1 + ctx.rank)).to(dev)
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

    for s in range(warmup):
        npx, _ = step(s)
    torch.cuda.synchronize()
    l0 = lib.iris_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        npx, outs = step(100 + s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_rays = (npx * spp * 7 + rays.shape[0]) * steps
    bpr = ray_bytes(sc.n_tris) + 16 + 36.0 / spp
    value = n_rays / (ms * 1e-3)
    return dict(value=value * ctx.world, ms=ms, steps=steps, launches=int(lib.iris_launch_count() - l0), bpr=bpr, stats=stats, triangles=sc.n_tris,
                finite=bool(all(torch.isfinite(o).all() for o in outs)))


# ------------------------------------------------------------------------------------------------ main
def main():
    a = parse()
    w = dict(WORKLOADS[a.workload])
    if a.views:
        w["views"] = a.views
    if a.workload == "c1":
        w["cornell"] = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    from iris_b200 import scenes
    config = dict(workload=a.workload, description=w["desc"], triangles=None, views_per_gpu=w["views"], width=w["width"], height=w["height"],
                  SPP=w["SPP"], spp=w["spp"], emitters=w["emitters"], slf_H=256,
                  sharding=("strong scaling: %d views in total, dealt round-robin to the ranks" % w["views"]) if w.get("strong") else
                  "weak scaling: N x views_per_gpu views in total; blocks of 256 consecutive pixels of every view are dealt round-robin to the ranks (all spp of a pixel on one rank); scene/SLF/field replicated; one allreduce of the gradient buffer per step",
                  l2="inputs larger than L2 (BVH+triangles 62 MB at 1M triangles / 300 MB at 5M, SLF 64 MB+, hash grid 56 MB, per-tile records > 400 MB)")

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
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    ctx = Ctx()
    if os.environ.get("IRIS_BENCH_OPTIONS"):          # A/B runs only: "name=value,name=value" -> iris_set_option (the defaults are what is benchmarked)
        from iris_b200 import core as _core
        for kv in os.environ["IRIS_BENCH_OPTIONS"].split(","):
            k, v = kv.split("=")
            _core.C.check(_core.C.lib().iris_set_option(k.strip().encode(), int(v)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=ctx.dev)
        ctx.dist = dist
    warmup = max(a.warmup, 3)
    peak, peak_src = peak_hbm()

    def finish(line):
        if rank == 0:
            print(json.dumps(line))
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()

    if w.get("bake"):
        r = run_bake(ctx, w, a.steps, warmup)
        config["triangles"] = r["triangles"]
        st = r["stats"]
        return finish(dict(metric="shading_map_bake_rays_per_sec", value=r["value"], unit="rays/s", n_gpus=world, steps=a.steps, warmup=warmup, ms_per_step=r["ms"] / a.steps,
                           higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=config, gpu_launches=r["launches"],
                           roofline=dict(bound="hbm", kernel="k_bake_persistent", achieved=r["value"] / world * r["bpr"] / 1e9, peak=peak, unit="GB/s",
                                         frac=r["value"] / world * r["bpr"] / 1e9 / peak, traffic=None, algorithmic_bytes_per_ray=r["bpr"]),
                           scene=dict(bvh_nodes=st["n_nodes"], bvh_build_ms=st["build_ms"], bvh_depth=st["max_depth"]), maps_finite=r["finite"]))
    if w.get("maps"):
        sc, scene, stats, tables = ctx.scene(w)
        config["triangles"] = sc.n_tris
        if rank == 0:
            bench_brdf(a, w, sc, scene, ctx.dev, config, stats)
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return

    r = run_estimator(ctx, a, w, a.steps, warmup, not a.no_e2e, True, a.tile)
    config["triangles"] = r["triangles"]
    value, ms, prof, stats = r["value"], r["ms"], r["prof"], r["stats"]

    # ---- roofline of the dominant kernel: k_trace_queue (2 of the 3 ray casts of a sample, through the ray queue); with the
    #      fused bounce selected (iris_set_option single_impl 0) it is k_bounce_single (the same casts + BSDF/emitter/SLF/MIS + record)
    R, est_bytes, spp = r["R"], r["est_bytes"], w["spp"]
    kname = "k_trace_queue" if "k_trace_queue" in prof else "k_bounce_single"
    kb = prof.get(kname)
    roof = None
    if kb:
        # this kernel's ALGORITHMIC bytes per sample: the two secondary ray casts it performs, 2 x R(F)
        k_bytes = 2 * R if kname == "k_trace_queue" else 2 * R + 16 + 24.0 / spp
        per_launch_samples = r["n_samples_rank"] * a.steps / kb["launches"]
        avg_ms = kb["total_ms"] / kb["launches"]
        ach = per_launch_samples * k_bytes / (avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname + ("_5m" if r["triangles"] > 2_000_000 else ""))
        except Exception:
            pass
        roof = dict(bound="hbm", kernel=kname, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                    peak_source=peak_src, algorithmic_bytes_per_sample=k_bytes,
                    samples_per_launch=per_launch_samples, avg_launch_ms=avg_ms,
                    step=dict(algorithmic_bytes_per_sample=est_bytes, achieved=value / max(world, 1) * est_bytes / 1e9, frac=value / max(world, 1) * est_bytes / 1e9 / peak),
                    kernel_share_of_step={k: v["total_ms"] / ms for k, v in prof.items()},
                    kernel_samples_per_s={k: r["n_samples_rank"] * a.steps / (v["total_ms"] * 1e-3) for k, v in prof.items()})

    # ---- short, driver-timed measurements of the other BASELINE configurations (device-resident inputs; 1 warm-up + 2 steps each)
    extra = None
    if a.workload == "c3" and not a.no_extras:
        extra = {}
        try:
            x = run_estimator(ctx, a, dict(WORKLOADS["c4"]), 2, 1, False, False, a.tile)
            extra["c4"] = dict(metric="path_samples_per_sec_fwd_bwd", value=x["value"], unit="samples/s", ms_per_step=x["ms"] / 2, config="BASELINE configs[3]: train_emitter step, emitter-radiance gradient only, 8 views 1280x960/GPU, SPP 256",
                               roofline_step_frac=x["value"] / world * x["est_bytes"] / 1e9 / peak, algorithmic_bytes_per_sample=x["est_bytes"])
            x = run_bake(ctx, dict(WORKLOADS["c2"]), 3, 2)
            extra["c2"] = dict(metric="shading_map_bake_rays_per_sec", value=x["value"], unit="rays/s", ms_per_step=x["ms"] / 3, config="BASELINE configs[1]: 13-map bake of one 640x480 view per GPU, spp 64",
                               roofline_frac=x["value"] / world * x["bpr"] / 1e9 / peak, algorithmic_bytes_per_ray=x["bpr"])
            extra["path_tracing"] = run_path_tracing(ctx, dict(WORKLOADS["c3"]), 2, 1)
            if rank == 0:
                extra["latency"] = run_latency(ctx, dict(WORKLOADS["c3"]), 50)
            ctx.drop_scenes()
            ws_ = dict(WORKLOADS["c3"])
            ws_["irregular"], ws_["views"] = True, 2       # the c3 step on a scan-like tessellation of the same room (scenes.room(irregular=True))
            x = run_estimator(ctx, a, ws_, 2, 1, False, True, a.tile)
            kq = x["prof"].get("k_trace_queue")
            extra["c3_scan_mesh"] = dict(metric="path_samples_per_sec_fwd_bwd", value=x["value"], unit="samples/s", ms_per_step=x["ms"] / 2, triangles=x["triangles"],
                                         config="c3 step (2 views/GPU) on the irregular 1M-triangle room: cell sizes varying ~10x, jittered vertices, random diagonals, per-vertex noise, shuffled face order",
                                         trace_queue_rays_per_s=(2 * x["n_samples_rank"] * 2 / (kq["total_ms"] * 1e-3) if kq else None),
                                         bvh_nodes=x["stats"]["n_nodes"], bvh_depth=x["stats"]["max_depth"], bvh_build_ms=x["stats"]["build_ms"])
            ctx.drop_scenes()
            w5 = dict(WORKLOADS["c5"])
            w5["views"], w5["strong"] = 8, False             # the per-GPU shard of the 64-view sweep at N = 8: 8 views of the 5M-triangle room per GPU
            x = run_estimator(ctx, a, w5, 2, 1, False, True, a.tile)
            kq = x["prof"].get("k_trace_queue")
            extra["c5_shard"] = dict(metric="path_samples_per_sec_fwd_bwd", value=x["value"], unit="samples/s", ms_per_step=x["ms"] / 2, triangles=x["triangles"],
                                     config="BASELINE configs[4] per-GPU shard: 5M-tri room, 8 views 1920x1440 per GPU, spp 128, field + emitter gradients (the full 64-view strong-scaling runs are under profiles/)",
                                     roofline_step_frac=x["value"] / world * x["est_bytes"] / 1e9 / peak, algorithmic_bytes_per_sample=x["est_bytes"],
                                     trace_queue_rays_per_s=(2 * x["n_samples_rank"] * 2 / (kq["total_ms"] * 1e-3) if kq else None),
                                     kernel_share_of_step={k: v["total_ms"] / x["ms"] for k, v in x["prof"].items()},
                                     bvh_nodes=x["stats"]["n_nodes"], bvh_build_ms=x["stats"]["build_ms"])
            ctx.drop_scenes()
        except Exception as e:                              # the extras never cost the headline line
            extra["error"] = "%s: %s" % (type(e).__name__, str(e)[:300])

    cb = None
    if not a.no_cpu_baseline and world == 1:
        sc = ctx.scene(w)[0]
        cb, _ = cpu_leg(w, sc, a.cpu_sample_pixels, 1, 1)

    line = dict(metric="path_samples_per_sec_fwd_bwd", value=value, unit="samples/s", n_gpus=world, steps=a.steps, warmup=warmup,
                ms_per_step=ms / a.steps, higher_is_better=True, scaling="strong" if w.get("strong") else "weak", vs_baseline=None, dtype="f32", data="synthetic", config=config,
                clocks=r["clocks"], e2e=r["e2e"], gpu_launches=r["launches"], roofline=roof, cpu_baseline=cb,
                scene=dict(bvh_nodes=stats["n_nodes"], bvh_build_ms=stats["build_ms"], bvh_depth=stats["max_depth"]),
                loss=r["loss"], d_radiance_abs_sum=r["d_radiance_abs_sum"], d_params_abs_sum=r["d_params_abs_sum"],
                d_crf_weight_abs_sum=r["d_crf_weight_abs_sum"], extra=extra)
    finish(line)


if __name__ == "__main__":
    main()
