"""Repeats the bench's e2e step (gm_index_create + gm_knn + gm_index_free on pinned host buffers) on a random table and
prints the per-phase trace (GM_TRACE=1): where does a step stall on the host?   python tools/e2e_stall_probe.py [n] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GM_TRACE", "1")
from guidemaker_b200 import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
_capi.init(0)
rng = np.random.default_rng(0)
t = np.unique(rng.integers(0, 1 << 40, size=n, dtype=np.uint64))
q = t.copy()
idx = np.zeros((len(q), 5), np.int32); dist = np.zeros((len(q), 5), np.uint8)
for r in range(reps):
    t0 = time.perf_counter()
    ix = _capi.Index(t, 20, 0)
    ix.knn(q, 5, out_idx=idx, out_dist=dist)
    ix.close()
    print("step %d: %.1f ms" % (r, 1e3 * (time.perf_counter() - t0)), flush=True)
