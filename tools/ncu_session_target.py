"""Short command for ncu on the HBM-bound kernels of the session path (K1 scan, K2 dedupe, text columns): one
device-resident scan of a named synthetic genome + seed flags + distinct-guide table.  Never report its timings."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi  # noqa: E402
from guidemaker_b200.synth import config_genome  # noqa: E402

_capi.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c5_arabidopsis_120Mb"
recs = config_genome(name)
buf = np.frombuffer(b"N".join(r.seq.encode() for r in recs), np.uint8)
rec_start = np.zeros(len(recs) + 1, np.int64)
rec_start[1:] = np.cumsum([len(r) + 1 for r in recs])
for _ in range(2):
    s = _capi.Session(buf, rec_start, "NGG", False, 20)
    rows = s.fetch_rows()
    t, c, e = s.fetch_text(30)
    dup = s.seed_dedup(10)
    ix, uniq, r2u = s.build_index(0)
    print("ok", s.n_rows, len(uniq), int(dup.sum()), int(e.sum()))
    ix.close(); s.close()
