"""world_size-2 gloo test of the N>1 host logic (pixel sharding, flat gradient allreduce, global-mean loss normalisation,
shard reassembly).  The compute kernels need a GPU; here each rank's "estimator" is a linear stand-in so the sharded result can
be checked against the single-process one exactly."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iris_b200 import dist as idist


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000003):
        for ws in (1, 2, 3, 8):
            r = [idist.shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_rows_partition_and_balance():
    for n, block, ws in ((1000, 64, 3), (1228800, 256, 8), (5, 256, 2), (0, 16, 4)):
        parts = [idist.interleaved_rows(n, block, r, ws) for r in range(ws)]
        allidx = torch.cat(parts).sort().values
        assert torch.equal(allidx, torch.arange(n))                       # a partition
        for p in parts:                                                   # whole blocks only: the spp samples of a pixel and its neighbours stay together
            assert all(int(b) // block == int(a) // block or int(b) % block == 0 for a, b in zip(p[:-1][:2000], p[1:][:2000]))
        if n >= block * ws * 4:
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= block


def _worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    torch.manual_seed(0)
    P, K = 1001, 5
    A = torch.randn(P, 3, K * 3)                     # stand-in estimator: L[p] = A[p] @ radiance.flatten()
    radiance = torch.randn(K, 3)
    target = torch.randn(P, 3)
    lo, hi = idist.shard_range(P)
    L_local = A[lo:hi] @ radiance.reshape(-1)
    step = idist.ShardedStep(P)
    loss_local, dL = step.mse_and_cotangent(L_local, target[lo:hi])
    d_rad = torch.einsum("pc,pck->k", dL, A[lo:hi]).reshape(K, 3)
    d_crf = torch.full((3, 2), float(rank + 1))
    d_big = torch.full((1 << 12,), float(rank + 1))                      # reduced in place by its own collective (`big` lowered for the test)
    keep_ptr = d_big.data_ptr()
    n = idist.allreduce_gradients([d_rad, None, d_crf, d_big], big=1 << 12)
    assert d_big.data_ptr() == keep_ptr and bool((d_big == 3.0).all())
    assert idist.rank_seed(5) == (5 + rank * 0x9E3779B97F4A7C15) % 2 ** 64
    loss = loss_local.clone()
    dist.all_reduce(loss)
    L_all = idist.gather_rows(L_local, P)
    if rank == 0:
        q.put((n, loss.item(), d_rad, d_crf, L_all))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_step_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n, loss, d_rad, d_crf, L_all = q.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    torch.manual_seed(0)
    P, K = 1001, 5
    A = torch.randn(P, 3, K * 3)
    radiance = torch.randn(K, 3, requires_grad=True)
    target = torch.randn(P, 3)
    L = A @ radiance.reshape(-1)
    ref = ((L - target) ** 2).mean()
    ref.backward()
    assert n == K * 3 + 6 + (1 << 12)
    assert abs(loss - ref.item()) < 1e-5 * abs(ref.item())
    assert torch.allclose(d_rad, radiance.grad, rtol=1e-4, atol=1e-6)
    assert torch.equal(d_crf, torch.full((3, 2), 3.0))
    assert torch.allclose(L_all, L.detach(), rtol=1e-5, atol=1e-6)
