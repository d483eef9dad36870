"""Multi-GPU plumbing: one process per GPU (torch.distributed), locations / prediction sites /
parameter sets sharded across ranks; the only collective on the fit path is a sum of a handful of
FP64 partials (sum log F, Q, or the (p+1)^2 cross forms) -- SURVEY §8e.
"""
from __future__ import annotations

import numpy as np


def nn_offset(i: int, m: int) -> int:
    """Offset of location i in the reference's ragged neighbour layout (src/util.cpp:55-70)."""
    return (i * (i - 1)) // 2 if i < m else (m * (m - 1)) // 2 + (i - m) * m


def shard_range(n: int, m: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous location range of `rank`, balanced by neighbour count (sum of m_i), so the ragged
    first m locations do not unbalance small problems.  Ranges of all ranks partition [0, n)."""
    total = nn_offset(n, m) + n  # weight m_i + 1 per location
    def boundary(r):
        if r <= 0:
            return 0
        if r >= world:
            return n
        target = total * r // world
        lo, hi = 0, n
        while lo < hi:  # first i with cumulative weight >= target
            mid = (lo + hi) // 2
            if nn_offset(mid, m) + mid < target:
                lo = mid + 1
            else:
                hi = mid
        return lo
    return boundary(rank), boundary(rank + 1)


def rebalance(sizes, times, n: int, damp: float = 1.0):
    """New contiguous shard sizes from measured per-shard kernel times: every shard's throughput (locations per unit
    time) is taken as measured and the n locations are re-dealt so that all shards would finish together.  The shards
    of a fit are NOT equally expensive per location: the head of an ordering has far-apart neighbours, which sends the
    Matern correlation through its slow branches -- at 8 shards of 10^6 locations the first shard's launch takes 0.132 ms
    against 0.114 ms for the others, and the slowest rank paces every sampler iteration.  Returns integer sizes that
    sum to n (order preserved)."""
    sizes = np.asarray(sizes, dtype=np.float64); times = np.asarray(times, dtype=np.float64)
    rate = sizes / np.maximum(times, 1e-12)                 # locations per unit time, per shard
    target = rate / rate.sum() * n                          # equal finishing time
    new = sizes + damp * (target - sizes)
    out = np.floor(new).astype(np.int64)
    out = np.maximum(out, 1)
    out[-1] += n - int(out.sum())                           # the remainder goes to the last shard
    if out[-1] < 1:
        raise ValueError("rebalance: degenerate shard sizes")
    return out


def balanced_shard_range(engine, n: int, m: int, rank: int, world: int, set_data, time_launch, allgather, rounds: int = 3):
    """Shard boundaries equalised by MEASURED launch time (setup, not part of a timed region).
    set_data(lo, hi) loads this rank's shard into the engine; time_launch() returns the B/F launch time of that shard;
    allgather(x) returns the list of every rank's float x.  Starts from `shard_range` (balanced by neighbour count) and
    applies `rebalance` `rounds` times; every rank computes the same boundaries.  Returns (lo, hi) with the engine
    holding that shard."""
    bounds = [shard_range(n, m, r, world) for r in range(world)]
    sizes = np.array([hi - lo for lo, hi in bounds], dtype=np.int64)
    for _ in range(rounds):
        lo = int(sizes[:rank].sum()); hi = lo + int(sizes[rank])
        set_data(lo, hi)
        times = np.array(allgather(float(time_launch())), dtype=np.float64)
        sizes = rebalance(sizes, times, n)
    lo = int(sizes[:rank].sum()); hi = lo + int(sizes[rank])
    set_data(lo, hi)
    return lo, hi


def split_even(count: int, rank: int, world: int) -> tuple[int, int]:
    """Even contiguous split (prediction sites, theta.alpha sets)."""
    return count * rank // world, count * (rank + 1) // world


def make_allreduce(device=None, group=None):
    """In-place sum over ranks of a small float64 numpy vector.  NCCL when `device` is a CUDA device
    (one staging tensor, reused), gloo otherwise.  Returns None when not running distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    use_cuda = device is not None and str(device).startswith("cuda")
    staging = torch.zeros(256, dtype=torch.float64, device=device if use_cuda else "cpu")

    def allreduce(buf: np.ndarray) -> None:
        k = buf.shape[0]
        t = staging[:k]
        t.copy_(torch.from_numpy(buf))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        buf[:] = t.cpu().numpy()

    return allreduce


def conj_kfold_sharded(engine, y, X, coords, m, cov, theta_alpha, sigmaSqIG, ho_idx, search_type="cb", rank=0, world=1,
                       allreduce=None):
    """k-fold scores of spConjNNGP with the parameter sets split over the ranks (zero communication
    until the end).  Each rank runs every fold for its sets (`Engine.conj_kfold`, set_range); the g x 2 score
    matrices have disjoint non-zero rows, so one sum over ranks assembles the result on every rank."""
    g = np.asarray(theta_alpha).shape[1]
    lo, hi = split_even(g, rank, world)
    scores = engine.conj_kfold(y, X, coords, m, cov, theta_alpha, sigmaSqIG, ho_idx, search_type, set_range=(lo, hi))
    if allreduce is not None and world > 1:
        flat = np.ascontiguousarray(scores.T).ravel()
        allreduce(flat)
        scores = flat.reshape(2, g).T.copy()
    return scores
