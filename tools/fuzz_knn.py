"""Randomised cross-check of gm_knn / gm_min_dist (both Hamming engines) against the CPU oracle on odd sizes.

    python tools/fuzz_knn.py [seconds] [seed]
Every case draws a table size, a query count, k and L, runs K3b and K3a, requires identical outputs and compares a
sample of rows (all rows for small cases) with oracle/gm_oracle.c.  Exits non-zero on the first mismatch."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def guides(rng, n, L, n_base):
    base = rng.integers(0, 4, size=(max(n_base, 1), L), dtype=np.uint64)
    rows = base[rng.integers(0, len(base), size=n)]
    mut = rng.random(n) < 0.5
    pos = rng.integers(0, L, size=n)
    rows[mut, pos[mut]] = rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint64)
    g = np.zeros(n, dtype=np.uint64)
    for i in range(L):
        g |= rows[:, i] << np.uint64(2 * i)
    return g


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    _capi.init(0)
    t_end, n_cases = time.time() + budget, 0
    while time.time() < t_end:
        L = int(rng.choice([1, 2, 5, 10, 16, 17, 20, 23, 27]))
        n_t = int(rng.choice([1, 3, 40, 129, 1024, 1025, 5000, 70000, 140000, 300000]))
        n_q = int(rng.choice([1, 31, 512, 513, 4097, 76000, 76289, 160000]))
        k = int(rng.choice([1, 2, 5, 20, 32]))
        t, _ = O.unique_first_order(guides(rng, n_t, L, max(n_t * 3 // 4, 1)))
        q = guides(rng, n_q, L, max(n_q // 2, 1))
        q[: min(len(q), len(t))] = t[: min(len(q), len(t))][::-1]          # exact hits and ties
        ix = _capi.Index(t, L, 0)
        _capi.knn_tune(8, 0, -1)
        _capi.knn_engine(1)
        bi, bd = ix.knn(q, k)
        md = ix.min_dist(q)
        _capi.knn_engine(0)
        ai, ad = ix.knn(q, k)
        ix.close()
        rows = np.arange(len(q)) if len(q) * len(t) < 3e8 else np.unique(np.concatenate([rng.integers(0, len(q), size=150), [0, len(q) - 1]]))
        oi, od = O.c_knn(t, q[rows], L, 0, k)
        ok = (np.array_equal(ai, bi) and np.array_equal(ad, bd) and np.array_equal(bi[rows], oi) and np.array_equal(bd[rows], od)
              and np.array_equal(md, bd[:, 0]))
        n_cases += 1
        print("case %d: L=%d targets=%d queries=%d k=%d -> %s" % (n_cases, L, len(t), len(q), k, "ok" if ok else "MISMATCH"), flush=True)
        if not ok:
            raise SystemExit(1)
    _capi.knn_engine(1)
    print("fuzz: %d cases, all identical to the oracle" % n_cases)


if __name__ == "__main__":
    main()
