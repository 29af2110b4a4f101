"""Multi-GPU plumbing for the kNN step: one process per GPU, queries sharded, results gathered.

The distinct-guide table is replicated on every rank (<= 80 MB for 10^7 guides); the query rows are
cut into ``world_size`` contiguous ranges; each rank computes the complete top-k of its rows, so
no cross-rank merge is needed -- only one fixed-size all-gather of ``(int32 idx[k], uint8 dist[k])``
per query (NCCL over NVLink on the GPU box; gloo in the CPU tests).  SURVEY.md section 8(e).
"""
from __future__ import annotations

import numpy as np


def world():
    """(rank, world_size) of the initialised torch.distributed group, else (0, 1).

    torch is consulted only if the caller has already imported it: without that there cannot be an initialised
    process group, and importing torch here would add seconds to a single-GPU run."""
    import sys
    if "torch" not in sys.modules:
        return 0, 1
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous, balanced row range of `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _gather_rows(local: np.ndarray, n_total: int, rank: int, world_size: int) -> np.ndarray:
    """all-gather row blocks of unequal size (pad to the largest shard, then trim)."""
    import torch
    import torch.distributed as dist
    rows = max(shard_bounds(n_total, r, world_size)[1] - shard_bounds(n_total, r, world_size)[0] for r in range(world_size))
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad = np.zeros((rows,) + local.shape[1:], dtype=local.dtype)
    pad[: len(local)] = local
    t = torch.from_numpy(pad).to(device)
    out = torch.empty((world_size * rows,) + tuple(local.shape[1:]), dtype=t.dtype, device=device)
    dist.all_gather_into_tensor(out, t)
    out = out.cpu().numpy().reshape((world_size, rows) + local.shape[1:])
    parts = []
    for r in range(world_size):
        lo, hi = shard_bounds(n_total, r, world_size)
        parts.append(out[r, : hi - lo])
    return np.concatenate(parts, axis=0)


def sharded_knn(index, q2bit: np.ndarray, k: int):
    """kNN of all queries with the rows split over the ranks; every rank returns the full result."""
    rank, ws = world()
    if ws == 1:
        return index.knn_packed(q2bit, k)
    lo, hi = shard_bounds(len(q2bit), rank, ws)
    idx, dist_ = index.knn_packed(q2bit[lo:hi], k)
    return _gather_rows(idx, len(q2bit), rank, ws), _gather_rows(dist_, len(q2bit), rank, ws)


def sharded_min_dist(index, q2bit: np.ndarray) -> np.ndarray:
    rank, ws = world()
    if ws == 1:
        return index.min_dist_packed(q2bit)
    lo, hi = shard_bounds(len(q2bit), rank, ws)
    d = index.min_dist_packed(q2bit[lo:hi])
    return _gather_rows(d, len(q2bit), rank, ws)
