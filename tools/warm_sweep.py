"""Sweep of the neighbourhood warm start (window size x sorted copies); one process per setting (env read at first use)."""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from guidemaker_b200 import _capi
    from guidemaker_b200.synth import config_genome
    _capi.init(0)
    recs = config_genome(sys.argv[2])
    buf = b"N".join(r.seq.encode() for r in recs)
    g, s, p, nf, nr = _capi.pam_scan(buf, "NGG", False, 20)
    first = _capi.first_occurrence(g)
    uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
    ix = _capi.Index(uniq, 20, 0)
    _capi.prof_enable(True)
    best = 1e30
    for i in range(4):
        _capi.prof_reset()
        out = ix.knn(g, 5)
        if i:
            best = min(best, _capi.prof_read()["scan_kernel_ms"])
    print(f"W {os.environ.get('GM_WARM_WINDOW')} copies {os.environ.get('GM_WARM_COPIES')}: {best:.2f} ms  checksum {int(out[0].sum()) + int(out[1].sum())}", flush=True)
else:
    name = sys.argv[1] if len(sys.argv) > 1 else "c2_bacterial_6.3Mb"
    for copies in (1, 2, 3, 4):
        for w in (256, 512, 1024, 2048):
            env = dict(os.environ, GM_WARM_WINDOW=str(w), GM_WARM_COPIES=str(copies))
            subprocess.run([sys.executable, __file__, "one", name], env=env, check=False)
