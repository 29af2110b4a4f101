"""End-to-end genome run through the public API (find_targets -> find_unique_near_pam -> create_index ->
get_neighbors [-> get_control_seqs]) on a named synthetic configuration of BASELINE.json, with per-stage wall
times.  Works single-process or under torchrun (queries sharded over the ranks).

    python tools/genome_run.py c5_arabidopsis_120Mb [--controls N] [--dtype hamming|leven] [--pam NGG] [--check N]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--pam", default="NGG")
    ap.add_argument("--orientation", default="3prime")
    ap.add_argument("--length", type=int, default=20)
    ap.add_argument("--lsr", type=int, default=10)
    ap.add_argument("--dist", type=int, default=2)
    ap.add_argument("--knum", type=int, default=5)
    ap.add_argument("--dtype", default="hamming")
    ap.add_argument("--controls", type=int, default=0)
    ap.add_argument("--check", type=int, default=0, help="verify this many random query rows against the CPU oracle")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import yaml
    import guidemaker_b200 as gmk
    from guidemaker_b200 import _capi
    from guidemaker_b200.synth import config_genome
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=__import__("datetime").timedelta(seconds=120))
    _capi.init(local_rank)

    cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9},
                    "CONTROL": {"MINIMUM_HMDIST": 1, "CONTROL_SEARCH_MULTIPLE": [10, 100]}}, cfg)   # one round of 10 x n, SURVEY 8d
    cfg.close()
    t = {}
    t0 = time.perf_counter()
    recs = config_genome(args.config)
    t["genome_synthesis_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    df = gmk.PamTarget(args.pam, args.orientation, args.dtype).find_targets(recs, args.length)
    t["find_targets_s"] = time.perf_counter() - t0
    tp = gmk.TargetProcessor(df, lsr=args.lsr, editdist=args.dist, knum=args.knum)
    t0 = time.perf_counter()
    tp.check_restriction_enzymes([])
    tp.find_unique_near_pam()
    t["find_unique_near_pam_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    tp.create_index(cfg.name)
    t["create_index_s"] = time.perf_counter() - t0
    _capi.prof_enable(True)
    _capi.prof_reset()
    t0 = time.perf_counter()
    tp.get_neighbors(cfg.name)
    t["get_neighbors_s"] = time.perf_counter() - t0
    pr = _capi.prof_read()
    t["get_neighbors_kernel_s"] = pr["scan_kernel_ms"] / 1e3
    t["pipeline_wall_s"] = sum(t[k] for k in ("find_targets_s", "find_unique_near_pam_s", "create_index_s", "get_neighbors_s"))
    res = {"config": args.config, "world_size": world, "targets": len(df), "distinct_guides": len(tp.nmslib_index),
           "seed_duplicated": int(tp.targets["isseedduplicated"].sum()), "guides_kept": len(tp.neighbors),
           "comparisons": float(pr["pairs"]) * world, "dtype": args.dtype, "pam": args.pam, "k": args.knum}
    if args.controls:
        t0 = time.perf_counter()
        np.random.seed(40)
        cmin, cmed, cdf = tp.get_control_seqs(recs, configpath=cfg.name, length=args.length, n=args.controls)
        t["get_control_seqs_s"] = time.perf_counter() - t0
        res.update({"controls": args.controls, "control_queries": tp.ncontrolsearched, "control_min": float(cmin), "control_median": float(cmed)})
    if args.check and rank == 0:
        from oracle import oracle as O
        from guidemaker_b200._encode import encode_guides
        g = encode_guides(tp.targets["target"])
        rows = np.random.default_rng(0).integers(0, len(g), size=args.check)
        metric = 0 if args.dtype == "hamming" else 1
        oi, od = O.c_knn(tp.nmslib_index.uniq, g[rows], args.length, metric, args.knum, threads=os.cpu_count())
        gi, gd = tp.nmslib_index.knn_packed(g[rows], args.knum)
        res["oracle_rows_checked"] = int(args.check)
        res["oracle_match"] = bool(np.array_equal(oi, gi) and np.array_equal(od, gd))
    res["timing"] = {k: round(v, 4) for k, v in t.items()}
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    os.unlink(cfg.name)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
