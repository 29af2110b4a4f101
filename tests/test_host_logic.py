"""CPU-only tests of the host side: C-ABI surface, packing helpers, the Myers recurrence compiled
for the host, the array-backed neighbour mapping, and the 2-rank sharding plumbing (gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import ROOT, _OracleIndex


def test_c_abi_exports_match_header():
    """libgm_b200.so loads (no CUDA call) and exports every function include/gm_b200.h declares."""
    from guidemaker_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "gm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_capi.EXPORTS)
    lib = _capi.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gm_version() >= 100


def test_engine_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device the product raises instead of computing."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from guidemaker_b200 import _capi\n"
            "try:\n    _capi.pam_scan(b'ACGT'*10, 'NGG', False, 20)\n"
            "except _capi.EngineUnavailable as e:\n    print('RAISED', e)\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "RAISED" in out.stdout, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under guidemaker_b200/ may reference it"""
    pkg = os.path.join(ROOT, "guidemaker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# `engine` exists so that the multi-rank plumbing can be unit-tested on CPU boxes with an\n        # injected checker", ""), f


def test_encode_decode_roundtrip():
    from guidemaker_b200._encode import decode_guides, encode_guides, as_byte_matrix, matrix_to_strings
    rng = np.random.default_rng(0)
    for L in (1, 10, 20, 27):
        seqs = ["".join(rng.choice(list("ACGT"), size=L)) for _ in range(50)]
        g = encode_guides(seqs)
        assert g.tolist() == [O.pack(s) for s in seqs]
        assert [x.decode() for x in decode_guides(g, L)] == seqs
        assert matrix_to_strings(as_byte_matrix(seqs)).to_pylist() == seqs
    with pytest.raises(ValueError):
        encode_guides(["ACGN"])
    with pytest.raises(ValueError):
        encode_guides(["ACG", "ACGT"])
    with pytest.raises(ValueError):
        encode_guides(["A" * 28])
    assert len(encode_guides([])) == 0


def test_myers_recurrence_on_host(tmp_path):
    """distance.cuh (the exact code the Levenshtein kernel runs per pair) compiled for the host and
    compared with the textbook DP of the oracle on random and adversarial pairs."""
    src = tmp_path / "myers_check.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "%s/guidemaker_b200/csrc/distance.cuh"
int main(int argc, char** argv) {
    // stdin: L a b (packed guide2bit, decimal) per line -> stdout: hamming leven
    int L; unsigned long long a, b;
    while (scanf("%%d %%llu %%llu", &L, &a, &b) == 3) {
        uint32_t alo = gm::compress_even_bits_hd(a), ahi = gm::compress_even_bits_hd(a >> 1);
        uint32_t blo = gm::compress_even_bits_hd(b), bhi = gm::compress_even_bits_hd(b >> 1);
        printf("%%d %%d\n", gm::hamming_planes(alo, ahi, blo, bhi), gm::myers_planes(alo, ahi, blo, bhi, L));
    }
    return 0;
}
''' % ROOT)
    exe = tmp_path / "myers_check"
    subprocess.run(["g++", "-O2", "-x", "c++", str(src), "-o", str(exe)], check=True, capture_output=True)
    rng = np.random.default_rng(1)
    cases = []
    for L in (1, 2, 10, 19, 20, 23, 27):
        for _ in range(300):
            a = "".join(rng.choice(list("ACGT"), size=L))
            r = rng.random()
            if r < 0.3:
                b = "".join(rng.choice(list("ACGT"), size=L))
            elif r < 0.6:                       # shifted copy: indels make leven << hamming
                sh = int(rng.integers(1, 4))
                b = (a[sh:] + "".join(rng.choice(list("ACGT"), size=sh)))[:L]
            elif r < 0.8:
                b = list(a)
                for p in rng.integers(0, L, size=int(rng.integers(0, 4))):
                    b[p] = "ACGT"[int(rng.integers(4))]
                b = "".join(b)
            else:
                b = a[::-1]
            cases.append((L, a, b))
    inp = "\n".join("%d %d %d" % (L, O.pack(a), O.pack(b)) for L, a, b in cases)
    out = subprocess.run([str(exe)], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    for (L, a, b), line in zip(cases, out):
        h, lv = (int(x) for x in line.split())
        assert h == O.py_hamming(a, b), (a, b)
        assert lv == O.py_leven(a, b), (a, b)


def test_neighbor_map_and_index_facade():
    from guidemaker_b200.neighbors import ExactIndex, NeighborMap
    seqs = ["ACGTACGTAC", "ACGTACGTAA", "TTTTTTTTTT", "ACGAACGTAC"]
    uniq = O.pack_many(seqs)
    ix = ExactIndex(uniq, 10, 0, engine=_OracleIndex(uniq, 10, 0))
    idx, dist = ix.knn_packed(uniq, 3)
    nm = NeighborMap(np.concatenate([uniq[:3], uniq[:1]]), np.concatenate([idx[:3], idx[:1]]), np.concatenate([dist[:3], dist[:1]]), uniq, 10)
    assert list(nm) == seqs[:3] and len(nm) == 3 and list(nm.keys()) == seqs[:3]
    assert nm["ACGTACGTAC"] == {"target": "ACGTACGTAC", "neighbors": {"seqs": ["ACGTACGTAC", "ACGTACGTAA", "ACGAACGTAC"], "dist": [0, 1, 1]}}
    assert "ACGAACGTAC" not in nm and "nonsense" not in nm and 5 not in nm
    with pytest.raises(KeyError):
        nm["ACGAACGTAC"]
    assert nm.get("GGGGGGGGGG") is None
    # nmslib protocol: one-hot or plain strings in, (ids, doubled dists) out
    oh = "1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1 1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1 1 0 0 0 0 1 0 0"
    a = ix.knnQueryBatch([oh], k=2)
    b = ix.knnQueryBatch(["ACGTACGTAC"], k=2)
    assert a[0][0].tolist() == b[0][0].tolist() == [0, 1] and a[0][1].tolist() == [0, 2]
    assert ix.knnQueryBatch(["ACGTACGTAC"], k=10)[0][0].tolist() == [0, 1, 3, 2]      # fewer than k targets -> shorter result
    lev = ExactIndex(uniq, 10, 1, engine=_OracleIndex(uniq, 10, 1))
    assert lev.knnQueryBatch(["ACGTACGTAC"], k=2)[0][1].tolist() == [0, 1]            # leven distances are not doubled


def test_shard_bounds_cover_and_balance():
    from guidemaker_b200.sharding import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000003):
        for ws in (1, 2, 3, 8):
            b = [shard_bounds(n, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from guidemaker_b200.neighbors import ExactIndex
from guidemaker_b200.sharding import sharded_knn, sharded_min_dist, world
from tests.conftest import _OracleIndex
from oracle import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=int(sys.argv[2]))
rng = np.random.default_rng(5)
L = 20
t = rng.integers(0, 1 << 40, size=901, dtype=np.uint64)
q = np.concatenate([t[:333], rng.integers(0, 1 << 40, size=334, dtype=np.uint64)])     # 667 rows: uneven shards
for metric in (0, 1):
    ix = ExactIndex(t, L, metric, engine=_OracleIndex(t, L, metric))
    idx, d = sharded_knn(ix, q, 4)
    oi, od = O.c_knn(t, q, L, metric, 4)
    assert world() == (int(sys.argv[1]), int(sys.argv[2]))
    assert np.array_equal(idx, oi) and np.array_equal(d, od), "sharded result differs from single-rank result"
    assert np.array_equal(sharded_min_dist(ix, q), od[:, 0])
dist.barrier()
dist.destroy_process_group()
print("RANK_OK", sys.argv[1])
'''


@pytest.mark.parametrize("world_size", [2, 3])
def test_sharded_knn_gloo(world_size, tmp_path):
    """one process per rank over gloo: the gathered result equals the single-rank result byte for byte"""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world_size)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world_size)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "RANK_OK %d" % r in o, o


_API_WORKER = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np, yaml, tempfile
import torch.distributed as dist
from unittest import mock
from guidemaker_b200 import _capi, core
from tests.conftest import _OracleIndex, _OracleSession, Rec
rank, ws = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=ws)
_capi.Index, _capi.Session, _capi.init = _OracleIndex, _OracleSession, (lambda device=None: None)
rng = np.random.default_rng(11)
recs = [Rec("c%%d" %% i, "".join(rng.choice(list("ACGT"), size=6000))) for i in range(3)]
cfg = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9}, "CONTROL": {"MINIMUM_HMDIST": 3, "CONTROL_SEARCH_MULTIPLE": [10, 100]}}, cfg)
cfg.close()
df = core.PamTarget("NGG", "3prime", "hamming").find_targets(recs, 20)
tp = core.TargetProcessor(df, lsr=10, editdist=2, knum=3)
tp.find_unique_near_pam(); tp.create_index(cfg.name); tp.get_neighbors(cfg.name)
# every rank must hold the same, complete neighbour table
g, _, _, _, _ = _OracleSession(b"N".join(r.seq.encode() for r in recs), np.cumsum([0] + [len(r) + 1 for r in recs]), "NGG", False, 20).fetch_rows()
from oracle import oracle as O
uniq, _ = O.unique_first_order(g)
qmask = ~tp.targets["isseedduplicated"].to_numpy()
oi, od = O.c_knn(uniq, g[qmask], 20, 0, 3)
keep = od[:, 1] >= 2
assert len(tp.neighbors) == len(set(g[qmask][keep].tolist())), (len(tp.neighbors), int(keep.sum()))
# controls: DIFFERENT numpy seeds per rank -- rank 0's draw must be the one every rank searches and returns
np.random.seed(100 + rank)
cmin, cmed, cdf = tp.get_control_seqs(recs, configpath=cfg.name, length=20, n=20)
seqs = cdf["Sequences"].tolist()
gathered = [None] * ws
dist.all_gather_object(gathered, (seqs, cdf["Hamming distance"].tolist()))
assert all(x == gathered[0] for x in gathered), "ranks returned different control tables"
from guidemaker_b200._encode import encode_guides
true = O.c_min_dist(uniq, encode_guides(seqs, 20), 20, 0)
assert [float(x) for x in true] == cdf["Hamming distance"].tolist(), "control distances do not belong to the returned sequences"
dist.barrier(); dist.destroy_process_group()
print("RANK_OK", rank)
'''


def test_api_under_two_ranks_gloo(tmp_path):
    """find_targets -> get_neighbors -> get_control_seqs under a 2-rank group whose ranks seed numpy differently:
    the query rows are sharded, every rank ends with the full result, and the controls are rank 0's draw."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "api_worker.py"
    script.write_text(_API_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "RANK_OK %d" % r in o, o


def test_seqid_categorical_run_lengths_and_fallback():
    """find_targets builds `seqid` by run-length expansion when the rows are grouped by record (the session's order) and by
    a plain gather otherwise; categories are lexicographic either way, and records without rows stay categories"""
    import pandas as pd
    from guidemaker_b200.core import PamTarget
    ids = ["chrB", "chrA", "empty", "chrC"]
    grouped = np.array([0, 0, 0, 1, 3, 3], np.int32)
    shuffled = np.array([3, 0, 1, 0, 3, 0], np.int32)
    for rec in (grouped, shuffled, np.zeros(0, np.int32)):
        c = PamTarget._seqid_categorical(ids, rec)
        want = pd.Categorical([ids[i] for i in rec], categories=sorted(ids))
        assert list(c.categories) == sorted(ids)
        assert list(c) == list(want) and np.array_equal(c.codes, want.codes)
