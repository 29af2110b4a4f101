"""cProfile of the public API on a named synthetic configuration (host-side hot spots of find_targets .. get_neighbors).

    GM_TRACE=1 python tools/api_profile.py c5_arabidopsis_120Mb [n_lines]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c5_arabidopsis_120Mb"
    import tempfile
    import yaml
    import guidemaker_b200 as gmk
    from guidemaker_b200 import _capi
    from guidemaker_b200.synth import config_genome
    _capi.init(0)
    recs = config_genome(name)
    cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9}}, cfg)
    cfg.close()
    for rep in range(int(os.environ.get("PROFILE_REPS", "1")) - 1):      # earlier passes: the same run, unprofiled
        t = [time.perf_counter()]
        df = gmk.PamTarget("NGG", "3prime", "hamming").find_targets(recs, 20); t.append(time.perf_counter())
        tp = gmk.TargetProcessor(df, lsr=10, editdist=2, knum=5)
        tp.check_restriction_enzymes([]); tp.find_unique_near_pam(); t.append(time.perf_counter())
        tp.create_index(cfg.name); t.append(time.perf_counter())
        tp.get_neighbors(cfg.name); t.append(time.perf_counter())
        print("pass %d: find_targets %.3f  find_unique %.3f  create_index %.3f  get_neighbors %.3f  total %.3f"
              % (rep, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[4] - t[0]), flush=True)
        del df, tp
    pr = cProfile.Profile()
    t = [time.perf_counter()]
    pr.enable()
    df = gmk.PamTarget("NGG", "3prime", "hamming").find_targets(recs, 20); t.append(time.perf_counter())
    tp = gmk.TargetProcessor(df, lsr=10, editdist=2, knum=5)
    tp.check_restriction_enzymes([]); tp.find_unique_near_pam(); t.append(time.perf_counter())
    tp.create_index(cfg.name); t.append(time.perf_counter())
    tp.get_neighbors(cfg.name); t.append(time.perf_counter())
    pr.disable()
    print("find_targets %.3f  find_unique %.3f  create_index %.3f  get_neighbors %.3f  total %.3f (first call in a fresh process)"
          % (t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[4] - t[0]))
    pstats.Stats(pr).sort_stats("tottime").print_stats(int(sys.argv[2]) if len(sys.argv) > 2 else 30)
    os.unlink(cfg.name)


if __name__ == "__main__":
    main()
