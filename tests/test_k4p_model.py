"""CPU model of K4p (guidemaker_b200/csrc/knn.cu: knn_leven_prefix_kernel) -- checks the ARGUMENT, not the kernel.

Claims: (1) Myers' bit-parallel state (Pv, Mv) after j text bases depends only on the text's first j bases, so a target
may resume from the state its predecessor in the prefix-sorted table left at any kept level <= their common prefix;
(2) with the kernel's level rule (keep levels c0 .. c0+NL-1, resume from min(lcp, top), from scratch when lcp < c0) every
distance equals the plain DP; (3) full (distance, index) keys make the lists independent of the arrival order."""
import numpy as np
import pytest

from oracle import oracle as O

MASK32 = 0xFFFFFFFF


def myers_step(eq, pv, mv):
    """one text base (distance.cuh: myers_planes / knn.cu: GM_MYERS_STEP), 32-bit words"""
    xv = eq | mv
    xh = ((((eq & pv) + pv) & MASK32) ^ pv) | eq
    ph = mv | (~(xh | pv) & MASK32)
    mh = pv & xh
    ph = ((ph << 1) | 1) & MASK32
    mh = (mh << 1) & MASK32
    return (mh | (~(xv | ph) & MASK32)), (ph & xv)


def pattern_masks(query):
    pm = [0, 0, 0, 0]
    for i, b in enumerate(query):
        pm[b] |= 1 << i
    return pm


def distance_of(pv, mv, L):
    lm = (1 << L) - 1
    return L + bin(pv & lm).count("1") - bin(mv & lm).count("1")


def prefix_sharing_distances(query, targets_sorted, L, c0, levels):
    """distances of `query` to the prefix-sorted targets with the kernel's state reuse; also returns the steps executed"""
    pm = pattern_masks(query)
    top = c0 + levels - 1
    kept = {}                                         # level -> (pv, mv) of the previous target
    prev = None
    out, steps = [], 0
    for t in targets_sorted:
        pl = 0
        if prev is not None:
            lcp = next((j for j in range(L) if t[j] != prev[j]), L)
            pl = min(lcp, top)
            if pl < c0:
                pl = 0
        pv, mv = (MASK32, 0) if pl == 0 else kept[pl]
        for j in range(pl, L):
            pv, mv = myers_step(pm[t[j]], pv, mv)
            steps += 1
            if c0 <= j + 1 <= top:
                kept[j + 1] = (pv, mv)
        out.append(distance_of(pv, mv, L))
        prev = t
    return out, steps


@pytest.mark.parametrize("L,n,c0,levels", [(8, 300, 2, 2), (12, 500, 3, 2), (20, 400, 2, 4), (9, 64, 1, 2), (27, 200, 1, 3)])
def test_resumed_states_give_the_plain_distances(L, n, c0, levels):
    rng = np.random.default_rng(L * 1000 + n)
    base = rng.integers(0, 4, size=(max(n // 6, 1), L))
    t = base[rng.integers(0, len(base), size=n)].copy()               # families with long common prefixes
    mut = rng.integers(0, L, size=n)
    t[np.arange(n), mut] = rng.integers(0, 4, size=n)
    t = np.unique(t, axis=0)
    order = np.lexsort(t.T[::-1])                                     # base 0 is the most significant key
    ts = [tuple(int(b) for b in row) for row in t[order]]
    total = 0
    for _ in range(6):
        q = tuple(int(b) for b in rng.integers(0, 4, size=L))
        got, steps = prefix_sharing_distances(q, ts, L, c0, levels)
        total += steps
        qs = "".join("ACGT"[b] for b in q)
        want = [O.py_leven(qs, "".join("ACGT"[b] for b in row)) for row in ts]
        assert got == want
    assert total < 6 * len(ts) * L                                    # and it does skip work


def test_full_keys_make_the_lists_order_independent():
    """K4p sees the targets in prefix order, the plain kernel in index order: inserting by (distance, original index) with
    'key < bound' gives the same k smallest keys for any arrival order, ties at the k-th distance included"""
    rng = np.random.default_rng(7)
    for _ in range(50):
        n, k = int(rng.integers(5, 60)), int(rng.integers(1, 6))
        d = rng.integers(0, 4, size=n)                                 # few distinct distances: ties everywhere
        want = sorted((int(d[i]), i) for i in range(n))[:k]
        lst, bound = [], (np.inf, np.inf)
        for i in rng.permutation(n):
            key = (int(d[i]), int(i))
            if key < bound:
                lst.append(key)
                lst.sort()
                del lst[k:]
                if len(lst) == k:
                    bound = lst[-1]
        assert lst == want
