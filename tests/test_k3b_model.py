"""CPU model of the K3b filter protocol (guidemaker_b200/csrc/knn_tc.cu) -- checks the ARGUMENT, not the kernel.

The tensor-core pass only produces flags "target t is strictly closer to query q than q's current bound"; the bound an
MMA sees (the bias byte of A) may be stale, and the candidate warps serve the flagged (tile, query) events in any
order, inserting by full (distance, index) key.  The kernel's claim: as long as the tiles of a query enter the tensor
pipe in ASCENDING order, the final lists equal the exact top-k under (distance, index) whatever the staleness and the
service order.  This model draws random staleness and service orders and compares with brute force; it also shows
that the claim fails when tiles may overtake each other (the bug the issue token in the kernel prevents)."""
import numpy as np
import pytest

KEY_EMPTY = np.iinfo(np.int64).max


def brute_topk(dist_row, k):
    order = np.lexsort((np.arange(len(dist_row)), dist_row))[:k]
    return [(int(dist_row[i]), int(i)) for i in order]


def run_protocol(dist, k, tile, rng, in_order=True, max_lag=6, warm=0, subset=None, inherit_subset=False):
    """dist: (Q, N) exact distances.  Returns per-query sorted lists of (distance, index)."""
    Q, N = dist.shape
    n_tiles = (N + tile - 1) // tile
    lists = [[] for _ in range(Q)]                     # sorted by (d, idx), at most k entries
    bound = [(np.inf, np.inf)] * Q                     # insert iff key < bound (full key)
    bias = np.full(Q, np.inf)                          # what the MMA compares with: flag iff d < bias (strict)
    if warm:                                           # warm start: lists of the first `warm` targets are inherited
        for q in range(Q):
            lists[q] = brute_topk(dist[q, :warm], k)
            if len(lists[q]) == k:
                bound[q] = lists[q][-1]
                bias[q] = lists[q][-1][0]
    if subset is not None:
        # neighbourhood warm start (warm.cu): per query an ARBITRARY subset of the table (any positions).  The kernel keeps
        # only the subset's k-th distance w0, as an inclusive bound (w0 + 1, index 0) with empty lists.  inherit_subset
        # models the tempting alternative -- inherit the subset's lists and flag strictly below w0 -- which loses ties.
        for q in range(Q):
            best = sorted((int(dist[q, i]), int(i)) for i in subset[q])[:k]
            if len(best) < k:
                continue
            if inherit_subset:
                lists[q] = list(best)
                bound[q] = best[-1]
                bias[q] = best[-1][0]
            else:
                bound[q] = (best[-1][0] + 1, -1)
                bias[q] = best[-1][0] + 1
    first = warm // tile
    order = list(range(first, n_tiles))
    if not in_order:                                   # tiles may overtake each other by a few positions
        for i in range(len(order) - 1):
            if rng.random() < 0.5:
                order[i], order[i + 1] = order[i + 1], order[i]
    pending = []                                       # queued events: (ready_time, tile, q)
    pending_bias = []                                  # bound tightenings on their way to the bias byte: (time, q, d)
    clock = 0
    for t in order:
        clock += 1
        # bias updates that have "landed" by now
        for item in [p for p in pending_bias if p[0] <= clock]:
            bias[item[1]] = min(bias[item[1]], item[2])
        pending_bias = [p for p in pending_bias if p[0] > clock]
        lo, hi = t * tile, min((t + 1) * tile, N)
        flagged = (dist[:, lo:hi] < bias[:, None]).any(axis=1)          # the accumulator filter, per query
        for q in np.flatnonzero(flagged):
            pending.append((clock + int(rng.integers(0, max_lag)), t, int(q)))
        # candidate warps: serve a random subset of the ready events, in random order
        ready = [e for e in pending if e[0] <= clock]
        rng.shuffle(ready)
        served = ready[: int(rng.integers(0, len(ready) + 1))] if clock < len(order) else ready
        for e in served:
            pending.remove(e)
            _, tt, q = e
            a, b = tt * tile, min((tt + 1) * tile, N)
            for i in range(a, b):                      # exact re-check of every target of the chunk
                key = (int(dist[q, i]), i)
                if key < bound[q] and key not in lists[q]:
                    lists[q].append(key)
                    lists[q].sort()
                    del lists[q][k:]
                    if len(lists[q]) == k:
                        bound[q] = min(bound[q], lists[q][-1])
                        pending_bias.append((clock + int(rng.integers(0, max_lag)), q, lists[q][-1][0]))
    for _, tt, q in pending:                           # drain
        a, b = tt * tile, min((tt + 1) * tile, N)
        for i in range(a, b):
            key = (int(dist[q, i]), i)
            if key < bound[q] and key not in lists[q]:
                lists[q].append(key)
                lists[q].sort()
                del lists[q][k:]
                if len(lists[q]) == k:
                    bound[q] = min(bound[q], lists[q][-1])
    return lists


def small_case(rng, Q=24, N=700, L=6):
    q = rng.integers(0, 4, size=(Q, L))
    t = rng.integers(0, 4, size=(N, L))
    return (q[:, None, :] != t[None, :, :]).sum(axis=2)               # many ties: distances in 0..6


@pytest.mark.parametrize("seed", range(12))
def test_filter_protocol_is_exact_for_any_staleness_and_service_order(seed):
    rng = np.random.default_rng(seed)
    dist = small_case(rng)
    k = int(rng.choice([1, 2, 3, 5]))
    warm = int(rng.choice([0, 64, 128]))
    lists = run_protocol(dist, k, tile=32, rng=rng, in_order=True, warm=warm)
    for q in range(dist.shape[0]):
        assert lists[q] == brute_topk(dist[q], k), (seed, q)


def _random_subsets(rng, Q, N, size):
    return [rng.choice(N, size=size, replace=False) for _ in range(Q)]


@pytest.mark.parametrize("seed", range(10))
def test_bound_from_an_arbitrary_subset_is_exact_when_inclusive(seed):
    """warm.cu + knn_tc.cu (warm_any_subset): the k-th distance over any k guides, taken as an inclusive bound with empty
    lists, yields the exact (distance, index) top-k under the same staleness / service-order freedom"""
    rng = np.random.default_rng(300 + seed)
    dist = small_case(rng)
    k = int(rng.choice([1, 2, 3, 5]))
    subset = _random_subsets(rng, dist.shape[0], dist.shape[1], int(rng.choice([k, 8, 40])))
    lists = run_protocol(dist, k, tile=32, rng=rng, in_order=True, subset=subset)
    for q in range(dist.shape[0]):
        assert lists[q] == brute_topk(dist[q], k), (seed, q)


def test_inheriting_a_subsets_lists_with_a_strict_filter_loses_ties():
    """why the kernel does NOT inherit the warm lists: a guide at exactly the subset's k-th distance but with a lower
    index than the inherited entry is never flagged by the strict filter"""
    bad = 0
    for seed in range(30):
        rng = np.random.default_rng(2000 + seed)
        dist = small_case(rng)
        subset = _random_subsets(rng, dist.shape[0], dist.shape[1], 40)
        lists = run_protocol(dist, 3, tile=32, rng=rng, in_order=True, subset=subset, inherit_subset=True)
        bad += sum(lists[q] != brute_topk(dist[q], 3) for q in range(dist.shape[0]))
    assert bad > 0


def test_out_of_order_tiles_can_lose_ties():
    """Without the issue token a later tile's entry can tighten the bias before an earlier tile is multiplied; a target
    of the earlier tile at exactly that distance is then not flagged although it wins the tie by index."""
    bad = 0
    for seed in range(40):
        rng = np.random.default_rng(1000 + seed)
        dist = small_case(rng)
        lists = run_protocol(dist, 3, tile=32, rng=rng, in_order=False, max_lag=1)
        bad += sum(lists[q] != brute_topk(dist[q], 3) for q in range(dist.shape[0]))
    assert bad > 0


# ---- the K encoding of knn_tc.cu: 3 bytes per position, ternary query side, bias word last -----------------------------
def _permute_bits(x):
    r = 0
    for p in range(32):
        r |= ((x >> p) & 1) << ((p >> 2) + 8 * (p & 3))
    return r


def _k_chunks(L):
    words = 3 * ((L + 3) // 4) + 1
    return (((words + 3) // 4) + 1) & ~1


def _planes(codes):
    lo = sum((int(c) & 1) << i for i, c in enumerate(codes))
    hi = sum((int(c) >> 1) << i for i, c in enumerate(codes))
    return _permute_bits(lo), _permute_bits(hi)


def _bytes_of(word):
    return [(word >> (8 * m)) & 0xFF for m in range(4)]


def _target_row(codes, L):
    """B operand row of one target: int8 values, as the producer warps write them"""
    lo, hi = _planes(codes)
    e = [lo & ~hi, hi & ~lo, lo & hi]
    kw = 4 * _k_chunks(L)
    row = []
    for w in range(kw):
        word = (1 | (64 << 8)) if w == kw - 1 else (e[w % 3] >> (w // 3)) & 0x01010101
        row += _bytes_of(word)
    return np.array(row, dtype=np.uint8).view(np.int8).astype(np.int64)


def _query_row(c1, c2, tau1, tau2, L):
    """A operand row of two queries (weights 1 and 64) with their bias bytes"""
    lmask = _permute_bits((1 << L) - 1)
    kw = 4 * _k_chunks(L)
    pl = [_planes(c1), _planes(c2)]
    eA = [~(lo | hi) & lmask for lo, hi in pl]
    e = [[lo & ~hi, hi & ~lo, lo & hi] for lo, hi in pl]
    row = []
    for w in range(kw):
        if w == kw - 1:
            b = [31 - L + t + bin(a).count("1") for t, a in zip((tau1, tau2), eA)]
            row += [b[0], b[1], 0, 0]
        else:
            x = [(e[s][w % 3] >> (w // 3)) & 0x01010101 for s in range(2)]
            a = [(eA[s] >> (w // 3)) & 0x01010101 for s in range(2)]
            pos, neg = _bytes_of(x[0] + (x[1] << 6)), _bytes_of(a[0] + (a[1] << 6))
            row += [(p - n) & 0xFF for p, n in zip(pos, neg)]                  # __vsub4
    return np.array(row, dtype=np.uint8).view(np.int8).astype(np.int64)


@pytest.mark.parametrize("L", [1, 4, 5, 8, 9, 12, 16, 17, 20, 21, 23, 24, 25, 27])
def test_k_encoding_counts_matches_and_flags(L):
    rng = np.random.default_rng(L)
    assert _k_chunks(L) == (2 if L <= 8 else 4 if L <= 20 else 6)
    for _ in range(40):
        q1, q2, t = (rng.integers(0, 4, size=L) for _ in range(3))
        if rng.random() < 0.3:
            q1 = np.zeros(L, dtype=np.int64)                                   # all-A query: the most negative data bytes
        if rng.random() < 0.3:
            t = q2.copy()
        tau1, tau2 = int(rng.integers(0, 32)), int(rng.integers(0, 32))
        A, B = _query_row(q1, q2, tau1, tau2, L), _target_row(t, L)
        assert np.abs(A).max() <= 127 and len(A) == 16 * _k_chunks(L)
        acc = int(A @ B)
        m1, m2 = int((q1 == t).sum()), int((q2 == t).sum())
        assert acc == (m1 + 31 - L + tau1) + 64 * (m2 + 31 - L + tau2)
        assert 0 <= acc <= 4030                                                 # fits the packed 16-bit TMEM read-out
        assert bool(acc & (1 << 5)) == ((L - m1) < tau1)                        # flag iff strictly closer than the bound
        assert bool(acc & (1 << 11)) == ((L - m2) < tau2)
