"""Multi-GPU driver of the path: image pairs are independent (ImagePair touches only its two frames and
the global K, reference source/front-end/image-pair.cpp:30-71,143), so a batch is sharded contiguously
over the ranks of one box (one process per GPU), every rank keeps the frame table resident, and the only
collective is one final gather of the fixed-size result records (NCCL over NVLink on GPUs, gloo in the
CPU tests).  There is no collective inside the hot path.

Feature extraction of a window is replicated, not sharded: every rank extracts all frames it matches against (512
Tsukuba-size frames take ~4 ms on one B200, against ~100 ms per rank for the all-pairs matching they feed), which is
cheaper than an all-gather of variable-length feature lists and keeps the path free of collectives."""
import numpy as np

from . import capi


def shard_bounds(n_pairs, world, rank):
    """Contiguous [lo, hi) slice of the pair list owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(local, n_total, dist, device=None, dst=0, dtype=None):
    """Gather per-rank fixed-size record arrays (contiguous shards, in rank order) on rank `dst`: pair records
    (RESULT_DTYPE, the default), or the records of the other independent-unit stages (PNP_RESULT_DTYPE, BA_RESULT_DTYPE).
    `dist` is torch.distributed (initialised); returns the full array on dst, None elsewhere."""
    import torch
    dtype = capi.RESULT_DTYPE if dtype is None else dtype
    world, rank = dist.get_world_size(), dist.get_rank()
    item = dtype.itemsize
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = np.zeros(cap * item, np.uint8)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    buf[:raw.size] = raw
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    bucket = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bucket, dst=dst)
    if rank != dst:
        return None
    out = np.empty(n_total, dtype)
    for r, (lo, hi) in enumerate(sizes):
        out[lo:hi] = bucket[r].cpu().numpy()[:(hi - lo) * item].view(dtype)
    return out


def solve_pnp_sharded(ctx, worlds, images, K, dist=None, device=None, **kw):
    """pnp_solve problems are independent units like pairs: contiguous shard per rank, sampling keyed by the global
    problem index (problem_id_base), one final gather of the fixed-size records."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = shard_bounds(len(worlds), world, rank)
    res = np.zeros(0, capi.PNP_RESULT_DTYPE)
    if hi > lo:
        res, _ = ctx.pnp_solve_batch(worlds[lo:hi], images[lo:hi], K, problem_id_base=lo, **kw)
    if dist is None:
        return res
    return gather_records(res, len(worlds), dist, device=device, dtype=capi.PNP_RESULT_DTYPE)


def solve_pairs_sharded(ctx, descs, kps, pairs, K, dist=None, device=None, **kw):
    """Upload the frame table on this rank, solve this rank's shard, gather the records on rank 0.
    Sampling is sharding-invariant: pair i always draws its samples with pair_id = i."""
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = shard_bounds(len(pairs), world, rank)
    ctx.frames_upload(descs, kps)
    res = np.zeros(0, capi.RESULT_DTYPE)
    if hi > lo:
        res, _ = ctx.pair_batch(pairs[lo:hi], K, pair_id_base=lo, details=False, **kw)
    if dist is None:
        return res
    return gather_records(res, len(pairs), dist, device=device)
