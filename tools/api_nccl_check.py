"""The public API under torchrun with the NCCL backend (query rows sharded, device-side gather and neighbour filter):
every rank must end with the oracle-exact neighbour table and control table.   torchrun --nproc-per-node N tools/api_nccl_check.py"""
import os
import sys
import tempfile

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import guidemaker_b200 as gmk  # noqa: E402
from guidemaker_b200 import _capi  # noqa: E402
from guidemaker_b200._encode import encode_guides  # noqa: E402
from guidemaker_b200.synth import synthetic_genome  # noqa: E402
from oracle import oracle as O  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
_capi.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9}, "CONTROL": {"MINIMUM_HMDIST": 3, "CONTROL_SEARCH_MULTIPLE": [10, 100]}}, cfg)
cfg.close()
ok = True
for total, nrec, dtype, metric, k, editdist in ((300_000, 3, "hamming", 0, 5, 2), (150_001, 2, "leven", 1, 3, 3)):
    recs = synthetic_genome(total, nrec, 0.5, seed=7)
    df = gmk.PamTarget("NGG", "3prime", dtype).find_targets(recs, 20)
    tp = gmk.TargetProcessor(df, lsr=10, editdist=editdist, knum=k)
    tp.check_restriction_enzymes(["GGTCTC"])
    tp.find_unique_near_pam()
    tp.create_index(cfg.name)
    tp.get_neighbors(cfg.name)
    g = encode_guides(tp.targets["target"], 20)
    uniq, _ = O.unique_first_order(g)
    qmask = (~tp.targets["isseedduplicated"].to_numpy()) | (~tp.targets["hasrestrictionsite"].to_numpy().astype(bool))
    qg = g[qmask]
    oi, od = O.c_knn(uniq, qg, 20, metric, k, threads=os.cpu_count())
    first = np.zeros(len(qg), bool)
    first[np.unique(qg, return_index=True)[1]] = True
    sel = (od[:, 1] >= editdist) & first
    nb = tp.neighbors
    good = bool(np.array_equal(nb.codes, qg[sel]) and np.array_equal(nb.index_matrix(), oi[sel]) and np.array_equal(nb.distance_matrix(), od[sel]))
    np.random.seed(100 + rank)                                       # different seeds: rank 0's draw must win
    cmin, cmed, cdf = tp.get_control_seqs(recs, configpath=cfg.name, length=20, n=200)
    true = O.c_min_dist(uniq, encode_guides(cdf["Sequences"].tolist(), 20), 20, metric)
    good = good and [float(x) for x in true] == [float(x) for x in cdf["Hamming distance"]]
    box = [None] * world
    dist.all_gather_object(box, cdf["Sequences"].tolist())
    good = good and all(b == box[0] for b in box)
    ok = ok and good
    print("rank %d/%d %s: %d targets, %d kept -> %s" % (rank, world, dtype, len(g), len(nb), "ok" if good else "MISMATCH"), flush=True)
os.unlink(cfg.name)
dist.barrier()
dist.destroy_process_group()
if not ok:
    raise SystemExit(1)
sys.stdout.write("API_NCCL_OK %d\n" % rank)      # one write: the ranks share the pipe
sys.stdout.flush()
