"""A/B of the Levenshtein scan on the GPU: prefix-sharing kernel over the sorted table (K4p, engine 1) against the plain
kernel (engine 0); outputs must be identical.  `python tools/leven_ab.py [workload] [queries]`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi  # noqa: E402
from guidemaker_b200.synth import config_genome  # noqa: E402

_capi.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c2_bacterial_6.3Mb"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
recs = config_genome(name)
buf = b"N".join(r.seq.encode() for r in recs)
g, s, p, nf, nr = _capi.pam_scan(buf, "NGG", False, 20)
first = _capi.first_occurrence(g)
uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
q = g[:nq]
print(name, "queries", len(q), "guides", len(uniq), flush=True)
ix = _capi.Index(uniq, 20, 1)
_capi.prof_enable(True)
ref = None
for label, eng, r in (("K4p prefix-sharing, R=8", 1, 8), ("K4 plain, R=8", 0, 8), ("K4p prefix-sharing, R=4", 1, 4)):
    ix.tune(engine=eng, queries_per_thread=r)
    best = 1e30
    for i in range(3):
        _capi.prof_reset()
        out = ix.knn(q, 5)
        if i:
            best = min(best, _capi.prof_read()["scan_kernel_ms"])
    same = "" if ref is None else ("same" if np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]) else "DIFFERENT")
    if ref is None:
        ref = out
    print(f"{label:28s} {best:9.2f} ms  {len(q) * len(uniq) / best / 1e6:8.2f} e9 cmp/s  {same}", flush=True)
print("ok")
