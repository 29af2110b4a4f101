"""Short, single-purpose command for `ncu --set full`: one exact kNN pass (k=5) on the configs[1]
workload, twice (first launch warms up).  Run it plainly first; never report its timings."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi  # noqa: E402
from guidemaker_b200.synth import config_genome  # noqa: E402

_capi.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c2_bacterial_6.3Mb"
metric = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 0
engine = int(sys.argv[4]) if len(sys.argv) > 4 else 0
recs = config_genome(name)
buf = b"N".join(r.seq.encode() for r in recs)
g, s, p, nf, nr = _capi.pam_scan(buf, "NGG", False, 20)
first = _capi.first_occurrence(g)
uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
dup = _capi.seed_dedup(g, 20, 10, False)
q = g if nq <= 0 else g[:nq]
ix = _capi.Index(uniq, 20, metric)
ix.tune(engine=engine)                                  # per-handle: 1 = K3b tcgen05, 0 = K3a xor/popc
for _ in range(2):
    idx, dist = ix.knn(q, 5)
print("ok", len(g), len(uniq), int(dup.sum()), int(dist[:, 1].min()), "engine", engine, "metric", metric)
