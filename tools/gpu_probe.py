"""Exploration script for gpurun calls: instruction-rate microbenchmarks + pair-scan kernel timing
sweeps on a configs[1]-scale synthetic genome.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi  # noqa: E402


def genome(n, gc, seed):
    rng = np.random.default_rng(seed)
    return rng.choice(np.frombuffer(b"GCAT", np.uint8), size=n, p=[gc / 2, gc / 2, (1 - gc) / 2, (1 - gc) / 2]).tobytes()


def main():
    out = {}
    _capi.init(0)
    out["device"] = _capi.device_info()
    for name, what in (("popc", 0), ("lop3", 1), ("imad", 2)):
        out["mb_" + name] = _capi.microbench(what)
        print(name, "%.3e lane-ops/s" % out["mb_" + name], flush=True)
    n = int(os.environ.get("PROBE_BASES", 6_300_000))
    t0 = time.time()
    seq = genome(n, 0.66, 2)
    t1 = time.time()
    g, s, p, nf, nr = _capi.pam_scan(seq, "NGG", False, 20)
    t2 = time.time()
    first = _capi.first_occurrence(g)
    uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
    t3 = time.time()
    print("genome %.2fs scan %.3fs (%d hits) uniq %.3fs (%d)" % (t1 - t0, t2 - t1, len(g), t3 - t2, len(uniq)), flush=True)
    out["n_targets"], out["n_uniq"] = len(g), len(uniq)
    ix = _capi.Index(uniq, 20, 0)
    nq = int(os.environ.get("PROBE_QUERIES", 0))
    if nq:
        g = np.ascontiguousarray(g[:nq])
    _capi.prof_enable(True)
    rows = []
    variants = [(8, 0, -1, 0), (8, 0, -1, 1), (8, 1, -1, 1), (8, 2, -1, 1), (8, 4, -1, 1), (8, 0, 0, 1), (8, 0, 16384, 1)]
    if len(sys.argv) > 1:
        variants = [tuple(int(x) for x in v.split(",")) for v in sys.argv[1:]]
    ref = None
    for (r, sp, warm, eng) in variants:
        _capi.knn_tune(r, sp, warm)
        _capi.knn_engine(eng)
        for rep in range(2):
            _capi.prof_reset()
            t0 = time.time()
            idx, dist = ix.knn(g, 5)
            wall = time.time() - t0
            pr = _capi.prof_read()
        if ref is None:
            ref = (idx, dist)
        same = bool(np.array_equal(ref[0], idx) and np.array_equal(ref[1], dist))
        rate = pr["pairs"] / (pr["scan_kernel_ms"] * 1e-3)
        rows.append({"engine": eng, "R": r, "splits": sp, "warm": warm, "scan_ms": pr["scan_kernel_ms"], "wall_s": wall, "pairs_per_s": rate, "same": same})
        print(rows[-1], flush=True)
    out["variants"] = rows
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
