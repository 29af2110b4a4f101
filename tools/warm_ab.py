"""A/B of K3b's warm start on the GPU: neighbourhood windows in sorted copies of the table (default) against the
first-8192-guides sample and against no warm start; results must be identical.  `python tools/warm_ab.py [workload]`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi  # noqa: E402
from guidemaker_b200.synth import config_genome  # noqa: E402

_capi.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c2_bacterial_6.3Mb"
recs = config_genome(name)
buf = b"N".join(r.seq.encode() for r in recs)
g, s, p, nf, nr = _capi.pam_scan(buf, "NGG", False, 20)
first = _capi.first_occurrence(g)
uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
print(name, "queries", len(g), "guides", len(uniq), flush=True)


def run(ix, q, k, reps):
    _capi.prof_enable(True)
    out = None
    best = 1e30
    for i in range(reps + 1):
        _capi.prof_reset()
        out = ix.knn(q, k)
        ms = _capi.prof_read()["scan_kernel_ms"]
        if i > 0:
            best = min(best, ms)
    return out, best


ix = _capi.Index(uniq, 20, 0)
ref = None
for label, warm in (("neighbourhood windows (default)", -1), ("first 8192 guides", 8192), ("none", 0), ("window again", -1)):
    ix.tune(engine=1, warm_sample=warm)
    (idx, dist), ms = run(ix, g, 5, 3)
    same = "" if ref is None else ("same" if np.array_equal(idx, ref[0]) and np.array_equal(dist, ref[1]) else "DIFFERENT")
    if ref is None:
        ref = (idx, dist)
    print(f"{label:28s} {ms:9.2f} ms  {len(g) * len(uniq) / ms / 1e9:8.2f} e12 cmp/s  {same}", flush=True)

# parity over k and sparse query sets (windows straight from L2), against the XOR/POPC engine
rng = np.random.default_rng(5)
for k in (1, 2, 5, 10, 32):
    for nq in (1000, 70000):
        q = g[rng.choice(len(g), nq, replace=False)]
        ix.tune(engine=1, warm_sample=-1)
        a = ix.knn(q, k)
        ix.tune(engine=0, warm_sample=-1)
        b = ix.knn(q, k)
        ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        print("k", k, "queries", nq, "same" if ok else "DIFFERENT", flush=True)
        assert ok
# short guides: ties everywhere (4^8 = 65536 possible guides, 200 000 distinct impossible -> L = 10)
for L in (10, 13, 27):
    t = np.unique(rng.integers(0, 1 << (2 * L), 300000, dtype=np.uint64))
    rng.shuffle(t)
    q = np.concatenate([t[:20000], rng.integers(0, 1 << (2 * L), 20000, dtype=np.uint64)])
    jx = _capi.Index(t, L, 0)
    for k in (1, 3, 8):
        jx.tune(engine=1, warm_sample=-1)
        a = jx.knn(q, k)
        jx.tune(engine=0, warm_sample=-1)
        b = jx.knn(q, k)
        ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        print("L", L, "guides", len(t), "k", k, "same" if ok else "DIFFERENT", flush=True)
        assert ok
    jx.close()
print("ok")
