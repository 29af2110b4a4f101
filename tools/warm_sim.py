"""CPU simulation of K3b's list process under different warm starts (no GPU needed): how many candidate events a query
causes -- targets that pass the bound the query holds when the scan reaches them -- with (a) no warm start, (b) the
first-8192-guides sample whose lists split 0 inherits, (c) the neighbourhood bound of warm.cu (k-th distance over the
guides around the query's rank in `copies` sorted copies of the table, inclusive, lists start empty).

    python tools/warm_sim.py [workload] [queries] [window] [copies]

Measured on the GPU (profiles/r02_k3b_ablations.md): 12.5 events per query with (b), 8.8 with (c)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200.synth import config_genome  # noqa: E402
from oracle import oracle as O  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_bacterial_6.3Mb"
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 200
W = int(sys.argv[3]) if len(sys.argv) > 3 else 256
COPIES = int(sys.argv[4]) if len(sys.argv) > 4 else 3
K, L = 5, 20

recs = config_genome(name)
g = np.concatenate([O.c_pam_scan(r.seq.encode(), "NGG", False, L)[0] for r in recs])
uniq, _ = O.unique_first_order(g)
n = len(uniq)
print(name, "guides", n, "queries sampled", NQ, "window", W, "copies", COPIES, flush=True)


def planes(x):
    lo = np.zeros(len(x), np.uint32); hi = np.zeros(len(x), np.uint32)
    for i in range(L):
        b = (x >> np.uint64(2 * i)) & np.uint64(3)
        lo |= ((b & np.uint64(1)).astype(np.uint32) << np.uint32(i))
        hi |= ((b >> np.uint64(1)).astype(np.uint32) << np.uint32(i))
    return lo, hi


def rot(code, h):
    """positions rotated right by h within the L positions of a 2-bit code"""
    if h == 0:
        return code
    m = np.uint64((1 << (2 * L)) - 1)
    return ((code >> np.uint64(2 * h)) | (code << np.uint64(2 * (L - h)))) & m


tlo, thi = planes(uniq)
popc = np.array([bin(i).count("1") for i in range(1 << 16)], np.uint8)


def dist_to_all(q):
    qlo, qhi = planes(np.array([q], np.uint64))
    x = (tlo ^ qlo[0]) | (thi ^ qhi[0])
    return popc[x & 0xFFFF] + popc[x >> 16]


sorted_copies = []
for c in range(COPIES):
    key = rot(uniq, c * L // COPIES)
    order = np.argsort(key, kind="stable")
    sorted_copies.append((key[order], order))


def events(d, bound_key, lst):
    """scan in index order; returns the number of targets that pass the current bound (each one is an insertion)"""
    ev = 0
    lst = list(lst)
    cand = np.flatnonzero(d <= (bound_key[0] if bound_key[0] < 99 else 99))      # nothing above the initial distance ever passes
    for i in cand:
        key = (int(d[i]), int(i))
        if key < bound_key and key not in lst:
            ev += 1
            lst.append(key)
            lst.sort()
            del lst[K:]
            if len(lst) == K:
                bound_key = min(bound_key, lst[-1])
    return ev


rng = np.random.default_rng(1)
tot = {"none": 0, "first 8192 (lists inherited)": 0, "neighbourhood (inclusive bound)": 0, "oracle bound (final k-th distance, inclusive)": 0}
for q in g[rng.choice(len(g), NQ, replace=False)]:
    d = dist_to_all(q)
    # (a) no warm start
    tot["none"] += events(d, (99, -1), [])
    # (b) first 8192 guides: their top-k is inherited, the scan continues behind them
    w = sorted((int(d[i]), int(i)) for i in np.argsort(d[:8192], kind="stable")[:K])
    d_b = d.copy(); d_b[:8192] = 99
    tot["first 8192 (lists inherited)"] += events(d_b, w[-1], w)
    # (c) neighbourhood bound
    ids = []
    for c, (skey, order) in enumerate(sorted_copies):
        pos = int(np.searchsorted(skey, rot(np.array([q], np.uint64), c * L // COPIES)[0]))
        lo = min(max(pos - W // 2, 0), max(n - W, 0))
        ids.append(order[lo: lo + W])
    ids = np.unique(np.concatenate(ids))
    w0 = int(np.sort(d[ids])[K - 1])
    tot["neighbourhood (inclusive bound)"] += events(d, (w0 + 1, -1), [])
    # (d) the floor: the final k-th distance known in advance
    tot["oracle bound (final k-th distance, inclusive)"] += events(d, (int(np.sort(d)[K - 1]) + 1, -1), [])
for k, v in tot.items():
    print(f"{k:46s} {v / NQ:6.2f} events per query")
