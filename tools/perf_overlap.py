"""Does running the forward of one tile next to the adjoint of another tile (two CUDA streams) beat running them back to back?

The forward is bound by instruction issue (ray casts), the adjoint by the LSU's reduction rate (grid-gradient scatter): complementary
resources.  For the two to share every SM the persistent ray-cast grid and the scatter grid are capped (persist_ctas_per_sm,
scatter_ctas_per_sm).  Prints ms for: forward alone, adjoint alone, both on two streams, per cap setting.
    python tools/perf_overlap.py [tile_pixels]
"""
import sys
import torch
sys.path.insert(0, ".")
from iris_b200 import core, scenes

dev = torch.device("cuda", 0)
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
spp, n_chunks = 32, 4
sc = scenes.room(1_000_000, 16, seed=0)
scene = core.Scene(sc.vertices, sc.faces, 0)
params = torch.empty(9216 + 27954112).uniform_(-1e-4, 1e-4)
params[:9216].uniform_(-0.2, 0.2)
tables = core.ShadingTables.from_dicts(dev, sc.emitter_dict(), sc.slf_dict(256), params, sc.voxel_bounds())
rays_all = torch.as_tensor(sc.camera_rays(1280, 960, view=1)).to(dev)
raysA, raysB = rays_all[:tile].contiguous(), rays_all[tile:2 * tile].contiguous()
lib = core.C.lib()
wsF = torch.empty(lib.iris_single_workspace_bytes(tile, spp), dtype=torch.uint8, device=dev)
wsB = torch.empty(lib.iris_single_workspace_bytes(tile, spp), dtype=torch.uint8, device=dev)
dp = torch.zeros(9216 + 27954112, device=dev)
dL = torch.randn(tile, 3, device=dev)
sF, sB = torch.cuda.Stream(), torch.cuda.Stream()


def fwd(rays, seed0, ws):
    return [core.single_forward(scene, tables, rays, spp, core.Sampler(seed=seed0 + c), True, workspace=ws)[1] for c in range(n_chunks)]


def bwd(recs, ws):
    for r in recs:
        core.single_backward(tables, dL, spp, r, True, dp, workspace=ws)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


recsA = fwd(raysA, 0, wsF)
torch.cuda.synchronize()
keep = {}


def both():
    ev = torch.cuda.Event()
    ev.record()
    with torch.cuda.stream(sF):
        sF.wait_event(ev)
        keep["r"] = fwd(raysB, 10, wsF)
    with torch.cuda.stream(sB):
        sB.wait_event(ev)
        bwd(recsA, wsB)
    torch.cuda.current_stream().wait_stream(sF)
    torch.cuda.current_stream().wait_stream(sB)


def bwd_two_streams():
    """The adjoint chunks alternate between two streams (own workspaces): the latency-bound tcgen05 kernel of one chunk can sit next
    to the reduction-bound grid scatter of the other."""
    ev = torch.cuda.Event()
    ev.record()
    for k, r in enumerate(recsA):
        st, ws = (sF, wsF) if k % 2 == 0 else (sB, wsB)
        with torch.cuda.stream(st):
            if k < 2:
                st.wait_event(ev)
            core.single_backward(tables, dL, spp, r, True, dp, workspace=ws)
    torch.cuda.current_stream().wait_stream(sF)
    torch.cuda.current_stream().wait_stream(sB)


print("tile %d px x spp %d x %d chunks" % (tile, spp, n_chunks))
t1 = timed(lambda: bwd(recsA, wsB))
t2 = timed(bwd_two_streams)
print("adjoint of %d chunks: one stream %.2f ms, two streams %.2f ms (%.3f)" % (n_chunks, t1, t2, t2 / t1))
if len(sys.argv) > 2 and sys.argv[2] == "bwd":
    sys.exit(0)
for persist, scat in [(8, 0), (8, 2), (6, 2), (5, 2), (4, 2), (4, 3), (4, 0), (3, 3), (6, 1)]:
    core.C.check(lib.iris_set_option(b"persist_ctas_per_sm", persist))
    core.C.check(lib.iris_set_option(b"scatter_ctas_per_sm", scat))
    tf = timed(lambda: fwd(raysB, 10, wsF))
    tb = timed(lambda: bwd(recsA, wsB))
    tfb = timed(both)
    print("persist %d scatter %d : fwd %.2f ms  bwd %.2f ms  sum %.2f  two streams %.2f ms  (%.3f of the default serial sum)" % (persist, scat, tf, tb, tf + tb, tfb, 0.0 if "base" not in keep else tfb / keep["base"]))
    if "base" not in keep:
        keep["base"] = tf + tb
