"""Multi-GPU plumbing: one process per GPU (torchrun), path samples sharded by contiguous pixel ranges, scene/BVH, SLF and BRDF
field replicated, and ONE allreduce per optimiser step over the flat gradient buffer [emitter radiance K*3 | mlp+grid | crf].
The reference has no distributed code; this is the data-parallel layout of section 8e.  Works with any
torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous range [lo,hi) of rank `rank`: sizes differ by at most one, all spp samples of a pixel stay on one rank."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(t, rank=None, world_size=None):
    lo, hi = shard_range(t.shape[0], rank, world_size)
    return t[lo:hi]


def interleaved_rows(n_items, block=256, rank=None, world_size=None):
    """Index tensor of the rows rank `rank` owns when blocks of `block` consecutive rows (pixels of one image region) are dealt
    round-robin to the ranks: every rank sees every part of every view, so the ranks' loads are equal whatever a region costs
    ("interleave tiles across ranks rather than whole views").  All spp samples of a pixel stay on one rank."""
    if rank is None or world_size is None:
        rank, world_size = world()
    n_items, block = int(n_items), int(block)
    idx = torch.arange(n_items)
    if world_size == 1:
        return idx
    blk = idx // block
    return idx[(blk % world_size) == rank]


def allreduce_gradients(tensors, average=False, big=1 << 20):
    """Sum (or average) the given gradient tensors over all ranks, in place.  Contiguous fp32 tensors of at least `big` elements
    (the 112 MB field gradient) are reduced where they lie, each by its own collective; the small ones (emitter rows, CRF weights)
    travel together in one flat buffer.  All collectives are issued asynchronously and waited for at the end, so the small reduction
    and the copies around it overlap the big one.  None entries are skipped; returns the number of elements communicated."""
    ts = [t for t in tensors if t is not None]
    rank, ws = world()
    if ws == 1 or not ts:
        return 0
    in_place = [t for t in ts if t.numel() >= big and t.is_contiguous() and t.dtype == torch.float32]
    small = [t for t in ts if not any(t is b for b in in_place)]
    work = [dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True) for t in in_place]
    flat = None
    if small:
        flat = torch.cat([t.reshape(-1).float() for t in small])
        work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
    for w in work:
        w.wait()
    if average:
        for t in in_place:
            t /= ws
    if flat is not None:
        if average:
            flat /= ws
        o = 0
        for t in small:
            n = t.numel()
            t.copy_(flat[o:o + n].view_as(t))
            o += n
    return int(sum(t.numel() for t in ts))


def rank_seed(seed):
    """Decorrelate the Philox streams of data-parallel ranks that were seeded alike (const.set_random_seed seeds every process with
    the same value): rank r uses seed + r * golden-ratio constant (mod 2^64); rank 0 / single process keep the seed."""
    rank, ws = world()
    return (int(seed) + rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


def gather_rows(local, n_total):
    """Reassemble the per-rank pixel shards (rank order) into the full (n_total, ...) tensor on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


class ShardedStep:
    """Loss normalisation for a data-parallel step: each rank renders its pixel shard, the loss is the mean over ALL pixels, so the
    per-rank cotangent is d(mean)/dL = 2 (L - target) / (n_total * channels) and gradients are summed (not averaged) across ranks."""

    def __init__(self, n_total_pixels, channels=3):
        self.n_total = int(n_total_pixels)
        self.channels = channels

    def mse_and_cotangent(self, L_local, target_local):
        diff = L_local - target_local
        denom = float(self.n_total * self.channels)
        return (diff * diff).sum() / denom, diff * (2.0 / denom)
