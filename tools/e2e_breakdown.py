"""Where does the host-buffer (e2e) step spend its time?  Wall-clock per sub-call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guidemaker_b200 import _capi
from guidemaker_b200.synth import config_genome

_capi.init(0)
recs = config_genome("c2_bacterial_6.3Mb")
buf = b"N".join(r.seq.encode() for r in recs)
g, *_ = _capi.pam_scan(buf, "NGG", False, 20)
first = _capi.first_occurrence(g)
uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
h_uniq = pin(uniq.view(np.int64)).view(np.uint64); h_q = pin(g.view(np.int64)).view(np.uint64)
h_idx = pin(np.empty((len(g), 5), np.int32)); h_dist = pin(np.empty((len(g), 5), np.uint8))
for rep in range(4):
    t = [time.perf_counter()]
    ix = _capi.Index(h_uniq, 20, 0); t.append(time.perf_counter())
    _capi.prof_enable(True); _capi.prof_reset()
    ix.knn(h_q, 5, out_idx=h_idx, out_dist=h_dist); t.append(time.perf_counter())
    pr = _capi.prof_read()
    ix.close(); t.append(time.perf_counter())
    print("rep %d: index_create %.1f ms | knn %.1f ms (scan kernel %.1f ms) | free %.1f ms" %
          (rep, 1e3 * (t[1] - t[0]), 1e3 * (t[2] - t[1]), pr["scan_kernel_ms"], 1e3 * (t[3] - t[2])), flush=True)
# same index, repeated knn
ix = _capi.Index(h_uniq, 20, 0)
for rep in range(3):
    t0 = time.perf_counter(); _capi.prof_reset()
    ix.knn(h_q, 5, out_idx=h_idx, out_dist=h_dist)
    print("reuse rep %d: knn %.1f ms (scan %.1f)" % (rep, 1e3 * (time.perf_counter() - t0), _capi.prof_read()["scan_kernel_ms"]), flush=True)
