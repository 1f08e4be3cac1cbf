"""Multi-GPU host layer (one process per GPU, torch.distributed over NCCL/NVLink as plumbing).

The front-end shards only where the reference's data model allows it:

  * frame pairs of a replayed sequence are independent units -> contiguous blocks per rank, NO
    data-path collective (each rank builds its own pyramids and tracks its own pairs); results are
    optionally gathered for the host-side map bookkeeping (`gather_rows`).
  * the descriptor matcher (C5: 1M x 1M) has one real exchange: the train set is replicated
    (broadcast from the owning rank), query rows are sharded, and the per-row top-2 results are
    all-gathered.  Top-2 is per query row, so there is no cross-rank reduction.

The compute call is injected (`match_fn`), so the same sharding logic is exercised on CPU with the
gloo backend in tests/ and with libslamfe + NCCL on the GPUs.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n units owned by `rank` (blocks differ by at most one unit)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_rows(local, n_total, group=None):
    """All-gather row blocks of unequal length (rank order == row order). local: torch tensor [m, ...]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = shard_counts(n_total, world)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def match_hamming256_sharded(match_fn, q_all_or_local, t, nq_total, train_src=0, group=None, q_is_local=False):
    """Sharded brute-force Hamming top-2 (BASELINE config 5).

    match_fn(q_rows, t_rows) -> (idx[m,2], dist[m,2], pass[m]) as torch tensors on q_rows.device; it must enqueue
    on torch's current stream (FrontEnd.use_torch_stream()), which is what orders it after the broadcast and before
    the all-gather.
    q_all_or_local: either the full query set (every rank slices its own block) or, with
    q_is_local=True, this rank's block.  t: the train set; only rank `train_src`'s copy is used
    (it is broadcast, i.e. replicated over NVLink).  Returns the gathered (idx, dist, pass) for all
    nq_total query rows on every rank.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(nq_total, rank, world)
    q_local = q_all_or_local if q_is_local else q_all_or_local[lo:hi]
    assert q_local.shape[0] == hi - lo
    t = t.contiguous()
    dist.broadcast(t, src=train_src, group=group)
    idx, dst, ok = match_fn(q_local.contiguous(), t)
    return (gather_rows(idx, nq_total, group), gather_rows(dst, nq_total, group), gather_rows(ok, nq_total, group))


def shard_pairs(n_pairs, rank, world):
    """Frame pairs [lo, hi) of a replayed sequence owned by `rank` (no halo needed: a pair carries
    both of its frames; a frame shared by two pairs at a block edge is simply built on both ranks)."""
    return shard_range(n_pairs, rank, world)
