"""Small end-to-end case for compute-sanitizer: every kernel of the library, both Hamming engines and Levenshtein."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi
from guidemaker_b200.synth import synthetic_genome
_capi.init(0)
recs = synthetic_genome(60000, 3, 0.5, 9)
buf = b"N".join(r.seq.encode() for r in recs)
g, s, p, nf, nr = _capi.pam_scan(buf, "NGG", False, 20)
dup = _capi.seed_dedup(g, 20, 10, False)
first = _capi.first_occurrence(g)
uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
out = []
for engine in (1, 0):
    _capi.knn_engine(engine)
    for tune in ((8, 0, -1), (8, 3, 1024)):
        _capi.knn_tune(*tune)
        ix = _capi.Index(uniq, 20, 0)
        out.append(ix.knn(g, 5)); md = ix.min_dist(g[:1000]); ix.close()
assert all(np.array_equal(out[0][0], o[0]) and np.array_equal(out[0][1], o[1]) for o in out)
ix = _capi.Index(uniq, 20, 1)
li, ld = ix.knn(g[:2000], 3)
print("sanitize target ok", len(g), len(uniq), int(dup.sum()), int(ld[:, 1].min()))
