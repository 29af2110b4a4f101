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


def nccl_group() -> bool:
    """True iff a torch.distributed group with the NCCL backend is initialised (device-side collectives possible)"""
    if world()[1] == 1:
        return False
    import torch.distributed as dist
    return dist.get_backend() == "nccl"


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


def broadcast_rank0(a: np.ndarray) -> np.ndarray:
    """rank 0's array on every rank (same shape and dtype everywhere); identity for a single process."""
    rank, ws = world()
    if ws == 1:
        return a
    import torch
    import torch.distributed as dist
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    raw = np.ascontiguousarray(a).view(np.uint8).copy()
    t = torch.from_numpy(raw).to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().view(a.dtype).reshape(a.shape)


def sharded_knn(index, q2bit: np.ndarray, k: int):
    """kNN of all queries with the rows split over the ranks; every rank returns the full result."""
    rank, ws = world()
    if ws == 1:
        return index.knn_packed(q2bit, k)
    lo, hi = shard_bounds(len(q2bit), rank, ws)
    idx, dist_ = index.knn_packed(q2bit[lo:hi], k)
    return _gather_rows(idx, len(q2bit), rank, ws), _gather_rows(dist_, len(q2bit), rank, ws)


_pinned = {}


def _pinned_like(name: str, shape, dtype):
    """page-locked host staging buffers, kept between calls (allocating them costs more than the copy)"""
    import torch
    key = (name, tuple(shape), dtype)
    buf = _pinned.get(name)
    if buf is None or buf[0] != key:
        buf = (key, torch.empty(tuple(shape), dtype=dtype, pin_memory=True))
        _pinned[name] = buf
    return buf[1]


def _session_gather_dev(sess, engine, qmask: np.ndarray, k: int):
    """NCCL path: every rank compacts and searches ITS slice of the query rows on the device; the fixed-size result rows
    are all-gathered device to device.  -> (idx, dist) torch tensors [nq, k] on the device, query-row order."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    qrows = np.flatnonzero(qmask)
    nq = len(qrows)
    lo, hi = shard_bounds(nq, rank, ws)
    local = np.zeros(len(qmask), np.uint8)
    local[qrows[lo:hi]] = 1
    dev = torch.device("cuda", torch.cuda.current_device())
    sizes = [shard_bounds(nq, r, ws)[1] - shard_bounds(nq, r, ws)[0] for r in range(ws)]
    rows = max(sizes)
    d_idx = torch.full((rows, k), -1, dtype=torch.int32, device=dev)
    d_dist = torch.full((rows, k), 255, dtype=torch.uint8, device=dev)
    g_idx = torch.empty((ws * rows, k), dtype=torch.int32, device=dev)
    g_dist = torch.empty((ws * rows, k), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    sess.knn_dev(engine, local, k, d_idx.data_ptr(), d_dist.data_ptr(), stream)
    dist.all_gather_into_tensor(g_idx, d_idx)
    dist.all_gather_into_tensor(g_dist, d_dist)
    if nq != ws * rows:                                     # unequal shards: drop the padding rows (on the device)
        g_idx = torch.cat([g_idx[r * rows: r * rows + sizes[r]] for r in range(ws)])
        g_dist = torch.cat([g_dist[r * rows: r * rows + sizes[r]] for r in range(ws)])
    return g_idx, g_dist


def sharded_session_knn(sess, engine, qmask: np.ndarray, k: int):
    """kNN of the masked rows of a device-resident scan (``_capi.Session``) with the query rows split over the ranks.

    Single process: one call, results straight into host arrays.  NCCL: shards searched and gathered on the devices
    (``_session_gather_dev``), then copied once into page-locked host memory -- nothing bounces through the host between
    the kernel and the collective."""
    rank, ws = world()
    if ws == 1:
        return sess.knn(engine, qmask, k)
    import torch
    import torch.distributed as dist
    if dist.get_backend() != "nccl":                          # CPU test plumbing (gloo): host arrays through _gather_rows
        qrows = np.flatnonzero(qmask)
        lo, hi = shard_bounds(len(qrows), rank, ws)
        local = np.zeros(len(qmask), np.uint8)
        local[qrows[lo:hi]] = 1
        idx, dist_ = sess.knn(engine, local, k)
        return _gather_rows(idx, len(qrows), rank, ws), _gather_rows(dist_, len(qrows), rank, ws)
    g_idx, g_dist = _session_gather_dev(sess, engine, qmask, k)
    h_idx = _pinned_like("idx", g_idx.shape, torch.int32)
    h_dist = _pinned_like("dist", g_dist.shape, torch.uint8)
    h_idx.copy_(g_idx, non_blocking=True)
    h_dist.copy_(g_dist, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h_idx.numpy().copy(), h_dist.numpy().copy()


def sharded_session_neighbors(sess, engine, qmask: np.ndarray, k: int, editdist: int):
    """get_neighbors' selection with the search sharded over the ranks: gather the rows on the devices, then every rank
    filters the gathered table on its own GPU (``gm_session_filter_dev``) and copies only the kept rows to the host.
    -> (codes, idx, dist, n_short) as ``Session.neighbors``."""
    import torch
    g_idx, g_dist = _session_gather_dev(sess, engine, qmask, k)
    stream = torch.cuda.current_stream().cuda_stream
    return sess.filter_dev(qmask, k, editdist, g_idx.data_ptr(), g_dist.data_ptr(), stream)


def sharded_min_dist(index, q2bit: np.ndarray) -> np.ndarray:
    rank, ws = world()
    if ws == 1:
        return index.min_dist_packed(q2bit)
    lo, hi = shard_bounds(len(q2bit), rank, ws)
    d = index.min_dist_packed(q2bit[lo:hi])
    return _gather_rows(d, len(q2bit), rank, ws)
