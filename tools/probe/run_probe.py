"""Runs tools/probe/umma_mn_probe.cu: which TMEM lanes hold D = A^T B (M = 64) and is the MN-major reading of a K-major tile right?"""
import ctypes, sys, os
import torch
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "iris_b200", "_lib", "ab", "libprobe.so"))
L.probe_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
A = torch.randint(-3, 4, (128, 64), generator=g).half().to(dev)      # small integers: exact in fp16 / fp32
B = torch.randint(-3, 4, (128, 64), generator=g).half().to(dev)
want = (A.float().t() @ B.float()).cpu()                               # (64 features of A) x (64 features of B)
for swap in (0, 1):
    out = torch.full((128, 64), float("nan"), device=dev)
    rc = L.probe_launch(A.data_ptr(), B.data_ptr(), out.data_ptr(), swap)
    o = out.cpu()
    print("swap_lbo_sbo", swap, "rc", rc, "finite rows", int(torch.isfinite(o).all(-1).sum()))
    hits = {}
    for m in range(64):
        eq = (o == want[m]).all(-1).nonzero().reshape(-1).tolist()
        hits[m] = eq
    mapped = {m: l for m, l in hits.items() if l}
    print("  rows of D found in TMEM lanes:", len(mapped), "of 64;", "first few:", {m: mapped[m] for m in list(mapped)[:6]})
    if len(mapped) == 64:
        print("  lane(m) for m = 0..63:", [mapped[m][0] for m in range(64)])

# dgrad: D[sample][in] = sum_out A[sample][out] W[out][in], B read MN-major from the forward's weight tile
L.probe_dgrad_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
W = torch.randint(-3, 4, (64, 64), generator=g).half().to(dev)
out = torch.full((128, 64), float("nan"), device=dev)
rc = L.probe_dgrad_launch(A.data_ptr(), W.data_ptr(), out.data_ptr())
print("dgrad rc", rc, "A@W exact:", bool(torch.equal(out.cpu(), (A.float() @ W.float()).cpu())), " A@W^T:", bool(torch.equal(out.cpu(), (A.float() @ W.float().t()).cpu())))
