#!/usr/bin/env python
"""bench.py -- headline benchmark of the off-target hot path (BASELINE.json: "off-target comparisons/s
(20-nt Hamming kNN) + genome wall-time, 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N ...            # CPU arm: exact brute force on the host cores

WORKLOAD = BASELINE.json configs[4], the configuration the metric's target is quoted on ("120 Mb synthetic eukaryotic
genome, ~10^7 NGG targets, exact all-vs-all Hamming kNN"): seeded synthetic 120 Mb / 5 records / 36 % GC, PAM NGG 3prime,
20-nt guides, k = 5.  It fits one B200 (62 MB guide table), so it is the workload at every N (strong scaling).
One STEP = one exact all-vs-all kNN pass: every PAM target row (Q = 7.77e6) against the table of distinct guides
(N_u = 7.70e6) = 5.98e13 comparisons.  A comparison = one (query, indexed guide) distance evaluation.

  value    comparisons/s with queries and the guide table already resident in HBM (gm_knn_dev on torch's current
           stream), K steps timed with CUDA events between barrier + synchronize, max over ranks.  N > 1: the table is
           replicated, query rows are sharded over the ranks and the per-rank top-k is all-gathered with NCCL inside the
           timed region.
  e2e      the same metric with HOST buffers: every step copies the guide table and that step's query rows
           host->device from pinned memory, runs the search and reads the (idx, dist) rows device->host.  N = 1: the
           host-pointer C ABI (gm_index_create + gm_knn).  N > 1: gm_index_create + gm_knn_dev on the rank's shard, NCCL
           all-gather of the device rows, one copy into pinned host memory on every rank.
  roofline the pair-scan kernel (the dominant kernel), K3b tcgen05 kind::i8 GEMM: algorithmic int8 tensor ops
           (2*4L = 160 per comparison) against the kind::i8 MMA rate measured live on this GPU; the executed
           ops (K = 64 bytes per row of two queries = 64 per comparison) are reported beside it.
  cpu_baseline  the CPU oracle port (oracle/gm_oracle.c, exact brute force, AVX-512 VPOPCNTD + OpenMP over all host
           cores: "tuned": true) on a bounded sample of the same workload.
  extras   genome_wall: find_targets -> get_neighbors through the public Python API on the same genome, at every N
           (query rows sharded under torchrun); c4_controls: BASELINE configs[3] (12 Mb / 16 records) with the first
           control round of n = 100 000 (10^6 random queries, seed 40); c2: BASELINE configs[1] (6.3 Mb), the round-1
           workload, with the INT-pipe engine K3a beside it; leven: the K4 Levenshtein kernel (alt_metric).

The reference arm times the CPU brute force as its whole measurement: the reference's own engine (nmslib 2.1.1 HNSW)
is a third-party package absent from this image and from /opt/wheelhouse, so the oracle port -- the published
definition of the search nmslib approximates -- stands in.  Reference HNSW recall is therefore not
measurable here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "off-target comparisons/s (20-nt Hamming kNN)"
UNIT = "comparisons/s"
WORKLOAD = "c5_arabidopsis_120Mb"
WORKLOAD_C2 = "c2_bacterial_6.3Mb"
WORKLOAD_C4 = "c4_yeast_12Mb"
K_NEIGHBORS = 5
GUIDE_LEN = 20
PUBLISHED_REF_BRUTEFORCE = 2.17e8      # tests/GridOptimization.ipynb:147, nmslib brute_force, 4 threads, 3814^2 (other hardware)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        if self.thread:
            self.thread.join(timeout=5)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
def build_workload(use_gpu_scan: bool, name: str = WORKLOAD):
    """-> (records, guides u64[Q] in reference row order, uniq u64[N_u] in first-occurrence order, info)"""
    from guidemaker_b200.synth import CONFIGS, config_genome
    total, records, gc, seed = CONFIGS[name]
    t0 = time.perf_counter()
    recs = config_genome(name)
    t_gen = time.perf_counter() - t0
    buf = b"N".join(r.seq.encode() for r in recs)
    rec_start = np.zeros(len(recs) + 1, np.int64)
    rec_start[1:] = np.cumsum([len(r) + 1 for r in recs])
    if use_gpu_scan:
        from guidemaker_b200 import _capi
        sess = _capi.Session(np.frombuffer(buf, np.uint8), rec_start, "NGG", False, GUIDE_LEN)
        g = sess.fetch_rows()[0]
        ix, uniq, _ = sess.build_index(0)
        ix.close(); sess.close()
    else:
        from oracle import oracle as O
        g, gs, _, nf, nr = O.c_pam_scan(buf, "NGG", False, GUIDE_LEN)
        rec = np.searchsorted(rec_start, gs.astype(np.int64), side="right") - 1
        g = g[np.argsort(rec, kind="stable")]                       # the reference's row order: per record, forward then reverse
        first = O.c_first_occurrence(g)
        uniq = np.ascontiguousarray(g[first == np.arange(len(g))])
    info = {"workload": name, "genome_bases": total, "records": records, "gc": gc, "genome_seed": seed,
            "pam": "NGG", "pam_orientation": "3prime", "guide_len": GUIDE_LEN, "metric_space": "hamming", "k": K_NEIGHBORS,
            "queries": int(len(g)), "indexed_guides": int(len(uniq)), "comparisons_per_step": float(len(g)) * float(len(uniq))}
    return recs, g, uniq, info


def cpu_bruteforce_rate(uniq, queries, seconds_target: float):
    """exact CPU brute force (oracle port, all cores) on a bounded sample: ~seconds_target of work"""
    from oracle import oracle as O
    cores = os.cpu_count() or O.num_threads()          # all host threads, even under torchrun's OMP_NUM_THREADS=1
    probe = queries[: min(1024, len(queries))]
    t0 = time.perf_counter()
    O.c_knn_hamming_fast(uniq, probe, GUIDE_LEN, K_NEIGHBORS, threads=cores)
    rate = len(probe) * len(uniq) / (time.perf_counter() - t0)
    n = int(min(len(queries), max(1024, seconds_target * rate / len(uniq))))
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(len(queries), size=n, replace=False))
    sample = np.ascontiguousarray(queries[rows])
    t0 = time.perf_counter()
    O.c_knn_hamming_fast(uniq, sample, GUIDE_LEN, K_NEIGHBORS, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n * len(uniq) / dt, "unit": UNIT, "cores": cores, "kind": "port", "tuned": True,
            "simd": "AVX-512 VPOPCNTD, 16 pairs per instruction" if O.has_avx512_popcnt() else "scalar POPCNT (no AVX-512 VPOPCNTDQ on this host)",
            "sample": "%d of %d query rows (seeded random subset) x all %d indexed guides, %.1f s, oracle/gm_oracle.c gmo_knn_hamming_fast "
                      "(OpenMP, queries blocked 64 x 32 KB target chunks; same bits as the scalar checker gmo_knn)"
                      % (n, len(queries), len(uniq), dt)}, n, dt


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm.  Rank 0 only; other ranks exit without work."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    _, g, uniq, info = build_workload(use_gpu_scan=False, name=args.workload)
    per_step_s = 6.0
    times, ns = [], []
    for i in range(args.warmup + args.steps):
        cb, n, dt = cpu_bruteforce_rate(uniq, g, per_step_s)
        if i >= args.warmup:
            times.append(dt); ns.append(n)
    value = sum(ns) * float(len(uniq)) / sum(times)
    cb["value"] = value
    cb["sample"] = "each step: %d of %d query rows x all %d indexed guides (bounded sample of the workload)" % (ns[-1], len(g), len(uniq))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": info, "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "reference_recall": "not measurable: nmslib absent from the image and the wheelhouse",
            "note": "reference engine nmslib==2.1.1 (HNSW) is not installable here; this arm times the exact CPU brute force "
                    "it approximates (oracle port, tuned: AVX-512 + OpenMP) on all host cores. Reference's own published brute_force "
                    "figure: %.3g comparisons/s (4 threads, Carsonella 3814^2, unspecified laptop)" % PUBLISHED_REF_BRUTEFORCE}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def _api_config(min_hm=7, mult=(10, 100, 1000, 10000)):
    import tempfile
    import yaml
    cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9},
                    "CONTROL": {"MINIMUM_HMDIST": min_hm, "CONTROL_SEARCH_MULTIPLE": list(mult)}}, cfg)
    cfg.close()
    return cfg.name


def api_genome_wall(recs, k, controls=0):
    """find_targets -> get_neighbors (-> get_control_seqs) through the public Python API; every rank of an initialised
    process group takes part (query rows sharded), each ends with the full result"""
    import guidemaker_b200 as gmk
    cfg = _api_config(1, (10, 100)) if controls else _api_config()
    t = [time.perf_counter()]
    df = gmk.PamTarget("NGG", "3prime", "hamming").find_targets(recs, GUIDE_LEN); t.append(time.perf_counter())
    tp = gmk.TargetProcessor(df, lsr=10, editdist=2, knum=k)
    tp.check_restriction_enzymes([]); tp.find_unique_near_pam(); t.append(time.perf_counter())
    tp.create_index(cfg); t.append(time.perf_counter())
    tp.get_neighbors(cfg); t.append(time.perf_counter())
    wall = {"total_s": round(t[-1] - t[0], 3), "find_targets_s": round(t[1] - t[0], 3), "find_unique_near_pam_s": round(t[2] - t[1], 3),
            "create_index_s": round(t[3] - t[2], 3), "get_neighbors_s": round(t[4] - t[3], 3), "targets": len(df),
            "indexed_guides": len(tp.nmslib_index), "guides_kept": len(tp.neighbors),
            "api": "guidemaker_b200.PamTarget/TargetProcessor (pandas in/out)"}
    if controls:
        np.random.seed(40)
        t0 = time.perf_counter()
        cmin, cmed, cdf = tp.get_control_seqs(recs, configpath=cfg, length=GUIDE_LEN, n=controls)
        wall.update({"get_control_seqs_s": round(time.perf_counter() - t0, 3), "controls": controls, "control_queries": int(tp.ncontrolsearched),
                     "control_min_dist": float(cmin), "control_median_dist": float(cmed),
                     "control_round": "first round only (10 x n random GC-matched 20-mers, numpy seed 40): MINIMUM_HMDIST 1"})
    os.unlink(cfg)
    return wall


def run_ours(args):
    # Libraries (NCCL's version banner, torchrun notices) must not reach stdout: rank 0 prints ONE JSON line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from guidemaker_b200 import _capi
    from guidemaker_b200.sharding import shard_bounds

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    _capi.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=__import__("datetime").timedelta(seconds=600))
    dev = torch.device("cuda", local_rank)
    t_start = time.perf_counter()

    def barrier():
        if world > 1:
            dist.barrier()

    _capi.knn_engine(args.engine)
    recs, g, uniq, info = build_workload(use_gpu_scan=True, name=args.workload)
    k = K_NEIGHBORS
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    class Resident:
        """queries (this rank's shard) and the guide table in HBM; one step = one sharded kNN pass + all-gather"""

        def __init__(self, g, uniq, metric=0):
            self.Q, self.NU = len(g), len(uniq)
            self.lo, self.hi = shard_bounds(self.Q, rank, world)
            self.rows_max = max(shard_bounds(self.Q, r, world)[1] - shard_bounds(self.Q, r, world)[0] for r in range(world))
            self.d_uniq = torch.from_numpy(uniq.view(np.int64)).to(dev)
            self.d_q = torch.from_numpy(np.ascontiguousarray(g[self.lo:self.hi]).view(np.int64)).to(dev)
            self.ix = _capi.Index(None, GUIDE_LEN, metric, device_ptr=self.d_uniq.data_ptr(), n=self.NU, stream=stream)
            self.d_idx = torch.full((self.rows_max, k), -1, dtype=torch.int32, device=dev)
            self.d_dist = torch.full((self.rows_max, k), 255, dtype=torch.uint8, device=dev)
            if world > 1:
                self.g_idx = torch.empty((world * self.rows_max, k), dtype=torch.int32, device=dev)
                self.g_dist = torch.empty((world * self.rows_max, k), dtype=torch.uint8, device=dev)

        def step(self):
            flush.zero_()                                               # L2 flush between iterations
            self.ix.knn_dev(self.d_q.data_ptr(), self.hi - self.lo, k, self.d_idx.data_ptr(), self.d_dist.data_ptr(), stream)
            if world > 1:
                dist.all_gather_into_tensor(self.g_idx, self.d_idx)
                dist.all_gather_into_tensor(self.g_dist, self.d_dist)

        def full_result(self):
            fi = (self.g_idx if world > 1 else self.d_idx).cpu().numpy()
            fd = (self.g_dist if world > 1 else self.d_dist).cpu().numpy()
            if world > 1:
                pi, pd_ = [], []
                for r in range(world):
                    a, b = shard_bounds(self.Q, r, world)
                    pi.append(fi[r * self.rows_max: r * self.rows_max + (b - a)]); pd_.append(fd[r * self.rows_max: r * self.rows_max + (b - a)])
                fi, fd = np.concatenate(pi), np.concatenate(pd_)
            return fi, fd

        def close(self):
            self.ix.close()

    def timed(res, steps):
        """-> (ms total over `steps` steps, max over ranks; profile of the pair-scan kernel)"""
        _capi.prof_enable(True)
        _capi.prof_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            res.step()
        e1.record()
        torch.cuda.synchronize(); barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        prof = _capi.prof_read()
        _capi.prof_enable(False)
        return float(ms.item()), prof

    main = Resident(g, uniq)
    Q, NU = main.Q, main.NU
    lo, hi, rows_max = main.lo, main.hi, main.rows_max

    # the clock sampler starts with the warm-up: the warm-up steps are the identical load
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w = time.perf_counter()
    for _ in range(args.warmup):
        main.step()
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t_w], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)                      # every rank derives the SAME number of extra steps
    el = float(el.item())
    extra = int(min(500, max(0, -(-(1.0 - el) // (el / args.warmup)))))  # keep the GPU under load for >= 1 s before timing
    for _ in range(extra):
        main.step()
    torch.cuda.synchronize()
    n_warm = args.warmup + extra

    # ---- timed region: `value` ----------------------------------------------------------------------
    ms_total, prof = timed(main, args.steps)
    clocks = sampler.stop()
    clocks["window"] = "warm-up (%d steps, same load) + the %d timed steps" % (n_warm, args.steps)
    comparisons = float(Q) * float(NU)
    value = comparisons * args.steps / (ms_total * 1e-3)
    launches_per_step = prof["all_kernel_launches"] / args.steps

    # ---- result check (outside the timed region): rank 0 verifies a sample against the CPU oracle ------
    checked = None
    if rank == 0:
        from oracle import oracle as O
        full_idx, full_dist = main.full_result()
        rows = np.random.default_rng(1).integers(0, Q, size=256)
        oi, od = O.c_knn(uniq, g[rows], GUIDE_LEN, 0, k, threads=os.cpu_count())
        checked = bool(np.array_equal(full_idx[rows], oi) and np.array_equal(full_dist[rows], od))
        del full_idx, full_dist
        if not checked:
            raise SystemExit("bench.py: GPU result differs from the CPU oracle -- refusing to report a number")

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    popc_rate = _capi.microbench(0)                                  # lane-POPC/s of this GPU, measured now
    scan_ms = prof["scan_kernel_ms"] / max(prof["scan_kernel_launches"], 1)
    pairs_per_launch = prof["pairs"] / max(prof["scan_kernel_launches"], 1)
    achieved = pairs_per_launch / (scan_ms * 1e-3)
    alg_bytes = float(NU) * 8 + float(hi - lo) * (8 + 5 * k)
    traffic = {}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            traffic = json.load(open(tr))
        except Exception:  # noqa: BLE001
            traffic = {}
    if args.engine == 1:
        i8_rate = _capi.microbench(3)                                # int8 tensor ops/s (2 per MAC), measured now
        ops_alg, ops_exec = 2 * 4 * GUIDE_LEN, 64                    # 2*4L per comparison; executed: K = 64 bytes, 2 queries per row
        roofline = {"bound": "tensor", "achieved": achieved * ops_alg / 1e12, "peak": i8_rate / 1e12, "unit": "TOP/s (int8)",
                    "frac": achieved * ops_alg / i8_rate,
                    "peak_source": "measured live: gm_microbench(3), back-to-back tcgen05.mma kind::i8 128x256x32 on all SMs "
                                   "(MEASURED_PEAKS.json has no int8 figure; its 2 x bf16 burst proxy is 3354 TOP/s)",
                    "algorithmic_ops_per_comparison": ops_alg, "executed_ops_per_comparison": ops_exec,
                    "frac_executed": achieved * ops_exec / i8_rate,
                    "frac_in_round1_units": achieved * 96 / i8_rate,        # what the same rate would need with the 4-byte one-hot code (K = 96)
                    "encoding": "3 bytes per base (rank-minimal ternary/0-1 code), K = 64 for 20-nt guides: two 32-byte MMA K steps per tile "
                                "(the 4-byte one-hot code of round 1 needed three)",
                    "comparisons_per_s": achieved, "kernel": "knn_hamming_tc_kernel<KC> (+ the neighbourhood warm start: radix sort of the queries, warm_window_kernel)",
                    "kernel_ms_per_launch": scan_ms, "kernel_share_of_step": prof["scan_kernel_ms"] / ms_total,
                    "algorithmic_bytes_per_launch": alg_bytes, "hbm_equiv_gbs": alg_bytes / (scan_ms * 1e-3) / 1e9,
                    "traffic": traffic.get("knn_hamming_tc_kernel_dram_bytes_per_launch_" + args.workload)}
    else:
        roofline = {"bound": "int (XU pipe: 1 POPC per comparison; not hbm/tensor)", "achieved": achieved / 1e9, "peak": popc_rate / 1e9,
                    "unit": "Gcomparisons/s", "frac": achieved / popc_rate,
                    "peak_source": "measured live: gm_microbench(POPC), register-resident, whole GPU (16 POPC/clk/SM x 148 SM x SM clock)",
                    "kernel": "knn_hamming_scan_kernel<R>", "kernel_ms_per_launch": scan_ms,
                    "kernel_share_of_step": prof["scan_kernel_ms"] / ms_total,
                    "algorithmic_bytes_per_launch": alg_bytes, "hbm_equiv_gbs": alg_bytes / (scan_ms * 1e-3) / 1e9,
                    "traffic": traffic.get("knn_hamming_scan_kernel_dram_bytes_per_launch")}

    # ---- e2e: host buffers, copies inside the timed region ---------------------------------------------------
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()          # noqa: E731
    h_uniq = pin(uniq.view(np.int64)).view(np.uint64)
    h_q = pin(np.ascontiguousarray(g[lo:hi]).view(np.int64)).view(np.uint64)
    if world == 1:
        h_idx = pin(np.empty((hi - lo, k), np.int32)); h_dist = pin(np.empty((hi - lo, k), np.uint8))

        def step_e2e():
            hix = _capi.Index(h_uniq, GUIDE_LEN, 0)                    # H2D guide table
            hix.knn(h_q, k, out_idx=h_idx, out_dist=h_dist)            # H2D queries, kernels, D2H results
            hix.close()
        e2e_api = "ctypes: gm_index_create + gm_knn + gm_index_free (host buffers, pinned)"
        d2h = int((hi - lo) * k * 5)
    else:
        t_q = torch.from_numpy(h_q.view(np.int64))                     # pinned host tensors (views of the arrays above)
        hg_idx = torch.empty((world * rows_max, k), dtype=torch.int32, pin_memory=True)
        hg_dist = torch.empty((world * rows_max, k), dtype=torch.uint8, pin_memory=True)
        dq = torch.empty(hi - lo, dtype=torch.int64, device=dev)

        def step_e2e():
            hix = _capi.Index(h_uniq, GUIDE_LEN, 0)                    # H2D guide table (replicated on every rank)
            dq.copy_(t_q, non_blocking=True)                           # H2D this rank's query rows
            hix.knn_dev(dq.data_ptr(), hi - lo, k, main.d_idx.data_ptr(), main.d_dist.data_ptr(), stream)
            dist.all_gather_into_tensor(main.g_idx, main.d_idx)        # device rows -> every rank, over NVLink
            dist.all_gather_into_tensor(main.g_dist, main.d_dist)
            hg_idx.copy_(main.g_idx, non_blocking=True)                # ONE D2H of the gathered table into pinned memory
            hg_dist.copy_(main.g_dist, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hix.close()
        e2e_api = ("ctypes: gm_index_create + gm_knn_dev on the rank's query shard (H2D from pinned memory), NCCL all-gather of the "
                   "device rows, one D2H of the full table into pinned memory on every rank")
        d2h = int(world * rows_max * k * 5)

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize(); barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": comparisons * e2e_steps / float(dt.item()), "unit": UNIT,
           "h2d_bytes_per_step": int(NU * 8 + (hi - lo) * 8), "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": e2e_api}
    main.close()
    del main
    torch.cuda.empty_cache()

    extras = not args.no_extras
    # ---- genome wall time through the public Python API, on the workload's genome, at this N ----------------------
    wall = api_genome_wall(recs, k) if extras else None
    del recs
    # ---- BASELINE configs[3]: 12 Mb / 16 records + 100 000 controls (first round) ---------------------------------------
    c4 = None
    if extras:
        from guidemaker_b200.synth import config_genome
        c4 = api_genome_wall(config_genome(WORKLOAD_C4), k, controls=100000)
        c4["workload"] = WORKLOAD_C4

    # ---- BASELINE configs[1] (the round-1 workload) + the other Hamming engine + the Levenshtein kernel ----------------
    c2 = leven = None
    if extras:
        _, g2, u2, info2 = build_workload(use_gpu_scan=True, name=WORKLOAD_C2)
        r2 = Resident(g2, u2)
        for _ in range(3):
            r2.step()
        ms2, p2 = timed(r2, 5)
        other = 1 - args.engine
        _capi.knn_engine(other)
        r2.step(); torch.cuda.synchronize()
        ms_o, p_o = timed(r2, 1)
        _capi.knn_engine(args.engine)
        rate_o = p_o["pairs"] / (p_o["scan_kernel_ms"] * 1e-3)
        c2 = {"workload": WORKLOAD_C2, "queries": info2["queries"], "indexed_guides": info2["indexed_guides"],
              "value": float(info2["comparisons_per_step"]) * 5 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / 5,
              "kernel_comparisons_per_s_per_gpu": p2["pairs"] / (p2["scan_kernel_ms"] * 1e-3),
              "alt_engine": {"engine": "K3a xor/popc (INT pipes)" if other == 0 else "K3b tcgen05 kind::i8 GEMM",
                             "comparisons_per_s_per_gpu": rate_o, "kernel_ms": p_o["scan_kernel_ms"],
                             "frac_of_popc_peak": rate_o / popc_rate if other == 0 else None, "popc_peak_lane_ops_per_s": popc_rate}}
        r2.close()
        # K4: Levenshtein on the same table, a bounded slice of the query rows (the kernel is ~200x slower per pair):
        # the prefix-sharing scan (K4p, default) and the plain scan (engine 0) beside it; same bits
        nq = min(len(g2), 65536 * world)
        r4 = Resident(g2[:nq], u2, metric=1)
        r4.step(); torch.cuda.synchronize()
        ms4, p4 = timed(r4, 1)
        out4 = r4.full_result()
        r4.ix.tune(engine=0)
        r4.step(); torch.cuda.synchronize()
        ms4p, p4p = timed(r4, 1)
        same4 = all(np.array_equal(x, y) for x, y in zip(out4, r4.full_result()))
        lop_rate = _capi.microbench(1)
        rate4 = p4["pairs"] / (p4["scan_kernel_ms"] * 1e-3)
        rate4p = p4p["pairs"] / (p4p["scan_kernel_ms"] * 1e-3)
        leven = {"metric": "Levenshtein comparisons/s (20-nt, Myers bit-parallel, K4p: prefix-sorted table, shared DP states)",
                 "value": float(nq) * len(u2) / (ms4 * 1e-3),
                 "kernel_comparisons_per_s_per_gpu": rate4, "queries": int(nq), "indexed_guides": int(len(u2)),
                 "cell_updates_per_s_per_gpu": rate4 * GUIDE_LEN * GUIDE_LEN,
                 "lop3_peak_lane_ops_per_s": lop_rate, "alu_ops_per_comparison": 7 * GUIDE_LEN,
                 "frac_of_alu_peak": rate4 * 7 * GUIDE_LEN / lop_rate,
                 "plain_kernel": {"kernel_comparisons_per_s_per_gpu": rate4p, "frac_of_alu_peak": rate4p * 7 * GUIDE_LEN / lop_rate,
                                  "same_output": bool(same4)},
                 "note": "7 LOP3 (ALU pipe) + 3 IMAD (FMA pipe) per pair and text step, counted in the SASS; peak = LOP3 lane rate measured "
                         "live (gm_microbench(1)); alu_ops_per_comparison = 7 L is the ALGORITHMIC count of the plain recurrence (ncu: ALU "
                         "pipe 84 %, profiles/r02_ncu_full_knn_leven_scan.csv) -- K4p executes ~L - log4(n) of the L steps per pair, so its "
                         "fraction can exceed 1"}
        r4.close()

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        cpu, _, _ = cpu_bruteforce_rate(uniq, g, 12.0)

    if rank == 0:
        run = {"parallelism": "query rows sharded x%d, guide table replicated, NCCL all-gather of top-k" % world if world > 1 else "single GPU",
               "l2": "256 MiB buffer rewritten between timed iterations (L2 flush)", "queries_per_rank": rows_max}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "i8 (3-byte base code, int32 accumulate)" if args.engine == 1 else "u32", "data": "synthetic", "config": info, "run": run,
                "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(prof["all_kernel_launches"]), "gpu_launches_per_step": launches_per_step,
                "roofline": roofline, "engine": "K3b tcgen05 kind::i8 GEMM" if args.engine == 1 else "K3a xor/popc",
                "cpu_baseline": cpu, "genome_wall": wall, "c4_controls": c4, "c2": c2, "alt_metric": leven,
                "oracle_check_256_rows": checked, "published_reference_bruteforce_cps": PUBLISHED_REF_BRUTEFORCE,
                "reference_recall": "not measurable: nmslib absent from the image and the wheelhouse (the notebook reports 0.990-1.000 on Carsonella)",
                "bench_wall_s": round(time.perf_counter() - t_start, 1)}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--engine", type=int, choices=[0, 1], default=1, help="Hamming pair-scan engine: 1 = K3b tensor (default), 0 = K3a INT pipes")
    ap.add_argument("--workload", default=WORKLOAD, help="synthetic configuration (guidemaker_b200.synth.CONFIGS); default = BASELINE configs[4]")
    ap.add_argument("--no-extras", action="store_true", help="skip the API wall-time / C4 / C2 / Levenshtein legs (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
