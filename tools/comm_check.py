"""gm_knn_sharded under torchrun (torch is used ONLY to carry the 128-byte NCCL id from rank 0 to the others): every rank
must receive the full, oracle-exact result.   torchrun --nproc-per-node N tools/comm_check.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402
from guidemaker_b200 import _capi  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
_capi.init(local)
dist.init_process_group("gloo")
box = [_capi.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
comm = _capi.Comm(box[0], rank, world)
rng = np.random.default_rng(3)
t = np.unique(rng.integers(0, 1 << 40, size=400000, dtype=np.uint64))
ok = True
for nq in (1, max(world - 1, 1), 1001, 250007):
    q = np.concatenate([t[: nq // 2], rng.integers(0, 1 << 40, size=nq - nq // 2, dtype=np.uint64)])
    ix = _capi.Index(t, 20, 0)
    t0 = time.perf_counter()
    idx, d = comm.knn(ix, q, 5)
    dt = time.perf_counter() - t0
    rows = np.arange(nq) if nq < 2000 else np.unique(rng.integers(0, nq, size=300))
    oi, od = O.c_knn(t, q[rows], 20, 0, 5)
    good = bool(np.array_equal(idx[rows], oi) and np.array_equal(d[rows], od))
    si, sd = ix.knn(q, 5)                                              # the single-GPU call on the same rows
    good = good and np.array_equal(si, idx) and np.array_equal(sd, d)
    ok = ok and good
    print("rank %d/%d: q=%d -> %s (%.1f ms)" % (rank, world, nq, "ok" if good else "MISMATCH", 1e3 * dt), flush=True)
    ix.close()
comm.close()
dist.barrier()
dist.destroy_process_group()
if not ok:
    raise SystemExit(1)
sys.stdout.write("COMM_OK %d\n" % rank)          # one write: the ranks share the pipe
sys.stdout.flush()
