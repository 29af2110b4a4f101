"""Print every gm_microbench() rate of this GPU (roofline denominators and the MMA shape probes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guidemaker_b200 import _capi
_capi.init(0)
names = ["popc", "lop3", "imad", "i8 128x256x32 SS", "i8 128x128x32 SS", "i8 128x128x32 TS", "i8 128x256x32 TS", "i8 128x128x32 SS cycling tiles", "i8 128x128x32 TS cycling tiles", "i8 128x128x32 SS cycling, 1 commit / 3 MMAs", "i8 128x128x32 SS cycling, 2 commits / 3 MMAs", "i8 128x64x32 SS cycling, 2 commits / 3 MMAs"]
for w, n in enumerate(names):
    v = _capi.microbench(w)
    extra = "" if w < 3 else "  = %.0f MAC/clk/SM at 1.965 GHz" % (v / 2 / 148 / 1.965e9)
    print("%-32s %.4e ops/s%s" % (n, v, extra), flush=True)
