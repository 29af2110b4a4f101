"""Build and time variants of the K3b kernel (knn_tc.cu) side by side.

    python tools/tc_ablate.py build NAME=FLAGS [NAME=FLAGS ...]   (here, CPU: nvcc cross-compiles)
    python tools/tc_ablate.py run [NAME ...]                      (on the GPU box, inside one gpurun call)

`build` compiles guidemaker_b200/lib/variants/libgm_NAME.so with extra nvcc FLAGS (comma-separated, e.g.
-DGM_TC_ABL=1).  `run` times every variant on the configs[1]-scale table in its own subprocess (GM_B200_LIB
selects the library) and writes gpurun_out/tc_ablate.json.  Variants with GM_TC_ABL return wrong results by
construction: only their timing is meaningful; `same` says whether a variant matched K3a.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "guidemaker_b200", "lib", "variants")


def build(specs):
    from guidemaker_b200 import _build
    os.makedirs(VAR, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(VAR, f"libgm_{name}.so")
        cmd = [_build._nvcc()] + _build.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + \
              [os.path.join(_build.CSRC, s) for s in _build.SOURCES]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        print(name, "ok" if p.returncode == 0 else "FAILED\n" + out, flush=True)


def child():
    import numpy as np
    from guidemaker_b200 import _capi
    from tools.gpu_probe import genome
    _capi.init(0)
    n = int(os.environ.get("PROBE_BASES", 6_300_000))
    g, s, p, nf, nr = _capi.pam_scan(genome(n, 0.66, 2), "NGG", False, 20)
    first = _capi.first_occurrence(g)
    uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
    ix = _capi.Index(uniq, 20, 0)
    _capi.prof_enable(True)
    if os.environ.get("ABL_WARM"):
        _capi.knn_tune(8, 0, int(os.environ["ABL_WARM"]))
    res = {}
    ref = None
    for eng in ((0, 1) if os.environ.get("ABL_CHECK", "1") == "1" else (1,)):
        _capi.knn_engine(eng)
        best = 1e30
        for rep in range(3 if eng else 1):
            _capi.prof_reset()
            idx, dist = ix.knn(g, 5)
            best = min(best, _capi.prof_read()["scan_kernel_ms"])
        if eng == 0:
            ref = (idx, dist)
        else:
            res = {"scan_ms": best, "pairs_per_s": len(g) * len(uniq) / (best * 1e-3),
                   "same": None if ref is None else bool(np.array_equal(ref[0], idx) and np.array_equal(ref[1], dist))}
    print("RESULT " + json.dumps(res), flush=True)


def run(names):
    if not names:
        names = sorted(f[6:-3] for f in os.listdir(VAR) if f.startswith("libgm_") and f.endswith(".so"))
    out = {}
    for name in names:
        lib, _, opt = name.partition("+")                    # NAME+dbg: same library with GM_TC_DEBUG=1 (event counters; build it with -DGM_TC_STATS)
        env = dict(os.environ, GM_B200_LIB=os.path.join(VAR, f"libgm_{lib}.so"))
        if opt == "dbg":
            env["GM_TC_DEBUG"] = "1"
        elif opt.startswith("w"):                            # NAME+w16384: warm-start sample size
            env["ABL_WARM"] = opt[1:]
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True,
                               timeout=float(os.environ.get("ABL_TIMEOUT", 90)))
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
            out[name] = json.loads(line[-1][7:]) if line else {"error": (r.stdout + r.stderr)[-600:]}
            dbg = [ln for ln in r.stderr.splitlines() if ln.startswith("[tc_dbg]")]
            if dbg:
                out[name]["dbg"] = dbg[-5:]
        except subprocess.TimeoutExpired:
            out[name] = {"error": "timeout"}
        print(name, out[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tc_ablate.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    {"build": lambda: build(sys.argv[2:]), "run": lambda: run(sys.argv[2:]), "child": child}[sys.argv[1]]()
