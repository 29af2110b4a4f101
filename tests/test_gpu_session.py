"""GPU tests of the device-resident session (gm_session_*) and of the BASELINE.json configurations at full size.

* session rows / text columns / chained stages against the CPU restatement (tests/conftest.py::_OracleSession, which
  is built from oracle/gm_oracle.c) on ragged multi-record genomes;
* configs[2] (TTTV 5prime, 23-nt, Levenshtein) at full size, configs[3] (12 Mb, 16 records + controls),
  configs[4] (120 Mb) through the public API, each against the oracle on >= 512 sampled query rows;
* a bounded run of the randomised K3b = K3a = oracle fuzzer (tools/fuzz_knn.py).
"""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from oracle import oracle as O
from tests.conftest import ROOT, _OracleSession

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from guidemaker_b200 import _capi
    _capi.init(0)
    _capi.knn_engine(1)
    _capi.knn_tune(8, 0, -1)
    return _capi


def _genome(rng, sizes, gc=0.5):
    recs = []
    for n in sizes:
        s = rng.choice(np.frombuffer(b"GCAT", np.uint8), size=n, p=[gc / 2, gc / 2, (1 - gc) / 2, (1 - gc) / 2])
        for st in rng.integers(0, max(n - 8, 1), size=max(n // 2000, 1) if n > 50 else 0):
            s[st:st + rng.integers(1, 8)] = ord("N")
        if n > 100:
            idx = rng.integers(0, n, size=n // 1000)
            s[idx] = s[idx] + 32                                       # lower case: never matches, never in a target
        if n > 3000:                                                   # duplicated segment -> duplicate guides / seeds
            s[n // 2: n // 2 + 700] = s[100:800]
        recs.append(s.tobytes())
    return recs


def _join(recs):
    buf = b"N".join(recs)
    rec_start = np.zeros(len(recs) + 1, np.int64)
    rec_start[1:] = np.cumsum([len(r) + 1 for r in recs])
    return np.frombuffer(buf, np.uint8), rec_start


SESSION_CASES = {"ngg3p20": ("NGG", False, 20), "ngg5p20": ("NGG", True, 20), "tttv5p23": ("TTTV", True, 23),
                 "nngrrt3p21": ("NNGRRT", False, 21), "nag3p27": ("NAG", False, 27), "n5p10": ("NN", True, 10)}


@pytest.mark.parametrize("name", list(SESSION_CASES))
def test_session_rows_and_text_vs_oracle(capi, name):
    """row order, record-relative coordinates, strands, PAM codes, `target` text, 30-nt context and edge flags"""
    pam, five, L = SESSION_CASES[name]
    rng = np.random.default_rng(len(name) + L)
    # ragged records: empty, shorter than a guide, shorter than the context window, boundaries inside a 32-base word
    recs = _genome(rng, [5000, 0, 7, 45, 33, 12001, 64, 1, 3000, 31])
    buf, rec_start = _join(recs)
    s = capi.Session(buf, rec_start, pam, five, L)
    o = _OracleSession(buf, rec_start, pam, five, L)
    assert s.n_rows == o.n_rows and s.n_rows > 100
    for a, b, what in zip(s.fetch_rows(), o.fetch_rows(), ("guides", "start", "pamcode", "rec", "strand")):
        assert np.array_equal(a, b), what
    # exact_pam on the device: histogram of the packed codes, per-row category through a caller-made table
    g, st, p, r, f = s.fetch_rows(want_pamcode=False)
    assert p is None and np.array_equal(g, o.g) and np.array_equal(r, o.rec)
    hist = s.pam_histogram()
    assert hist.dtype == np.uint32 and np.array_equal(hist, np.bincount(o.p, minlength=1 << 16))
    lut = rng.integers(-128, 128, size=1 << 16).astype(np.int8)
    assert np.array_equal(s.pam_categories(lut), lut[o.p])
    t, c, e = s.fetch_text(30)
    ot, oc, oe = o.fetch_text(30)
    assert np.array_equal(t, ot) and np.array_equal(e, oe) and (e.any() or name not in ("ngg3p20", "ngg5p20"))
    assert np.array_equal(c[~e], oc[~oe]) and (c[e] == ord("?")).all()
    s.close()


def test_session_chain_vs_oracle(capi):
    """seed flags, restriction flags, distinct-guide table + row map, masked kNN (both metrics) off one session"""
    rng = np.random.default_rng(77)
    recs = _genome(rng, [12000, 2500, 20000, 17], gc=0.6)
    buf, rec_start = _join(recs)
    for pam, five, L in (("NGG", False, 20), ("TTTV", True, 23)):
        s = capi.Session(buf, rec_start, pam, five, L)
        o = _OracleSession(buf, rec_start, pam, five, L)
        g = o.g
        for lsr in (0, 10, L):
            assert np.array_equal(s.seed_dedup(lsr), o.seed_dedup(lsr)), lsr
        motifs = ["GGTCTC", "GAGACC", "RAATTY"]
        assert np.array_equal(s.restriction(motifs), o.restriction(motifs))
        for metric in (0, 1):
            ix, uniq, r2u = s.build_index(metric)
            ou, _ = O.unique_first_order(g)
            assert np.array_equal(uniq, ou) and ix.n == len(ou)
            assert np.array_equal(uniq[r2u], g)
            for qmask in (np.ones(len(g), bool), rng.random(len(g)) < 0.37, np.arange(len(g)) == 5):
                idx, dist = s.knn(ix, qmask, 4)
                oi, od = O.c_knn(ou, g[qmask], L, metric, 4)
                assert np.array_equal(idx, oi) and np.array_equal(dist, od), (pam, metric)
            ix.close()
        s.close()


def test_session_neighbors_filter_vs_oracle(capi):
    """gm_session_neighbors: the distance filter (core.py:512,518) and the one-entry-per-guide rule on the device"""
    rng = np.random.default_rng(91)
    recs = _genome(rng, [9000, 4000, 7000], gc=0.55)
    buf, rec_start = _join(recs)
    s = capi.Session(buf, rec_start, "NGG", False, 20)
    o = _OracleSession(buf, rec_start, "NGG", False, 20)
    g = o.g
    assert len(set(g.tolist())) < len(g)                                # duplicated guides exist
    ou, _ = O.unique_first_order(g)
    for metric in (0, 1):
        ix, uniq, _ = s.build_index(metric)
        for qmask in (np.ones(len(g), bool), ~o.seed_dedup(10), rng.random(len(g)) < 0.5):
            for k, editdist in ((5, 2), (2, 4), (3, 0)):
                codes, idx, dist, n_short = s.neighbors(ix, qmask, k, editdist)
                oi, od = O.c_knn(ou, g[qmask], 20, metric, k)
                keep = od[:, 1] >= editdist
                qg = g[qmask]
                first = np.zeros(len(qg), bool)
                first[np.unique(qg, return_index=True)[1]] = True       # first query row of every guide
                sel = keep & first
                assert n_short == 0
                assert np.array_equal(codes, qg[sel]) and np.array_equal(idx, oi[sel]) and np.array_equal(dist, od[sel]), (metric, k, editdist)
        ix.close()
    # fewer than two indexed guides: every row is short of a second hit
    s1 = capi.Session(np.frombuffer(b"ACGTACGTACGTACGTACGTAGG" + b"T" * 30, np.uint8), np.array([0, 54]), "NGG", False, 20)
    if s1.n_rows == 1:
        ix1, _, _ = s1.build_index(0)
        assert s1.neighbors(ix1, np.ones(1, bool), 2, 2)[3] == 1
    s.close()


def test_session_bad_arguments(capi):
    buf, rec_start = _join([b"ACGT" * 50])
    with pytest.raises(ValueError):
        capi.Session(buf, rec_start + 1, "NGG", False, 20)            # record table must start at 0
    with pytest.raises(ValueError):
        capi.Session(buf, rec_start, "NGZ", False, 20)                # not an IUPAC letter
    s = capi.Session(buf, rec_start, "NGG", False, 20)
    if s.n_rows:
        ix, _, _ = s.build_index(0)
        with pytest.raises(ValueError):                               # n_q does not match the mask
            capi._check(capi.load_library().gm_session_knn(s._h, ix._h, capi._p(np.ones(s.n_rows, np.uint8)), 1, 2,
                                                           capi._p(np.zeros((1, 2), np.int32)), capi._p(np.zeros((1, 2), np.uint8))), "x")


# ---- BASELINE.json configurations at full size, through the public API ------------------------------------------------
def _cfg(tmp_path, min_hm=7, mult=(10, 100, 1000, 10000)):
    import yaml
    p = tmp_path / "config.yaml"
    p.write_text(yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9},
                                 "CONTROL": {"MINIMUM_HMDIST": min_hm, "CONTROL_SEARCH_MULTIPLE": list(mult)}}))
    return str(p)


def _run_api(recs, pam, orientation, L, dtype, cfg, knum=5, dist=2):
    import guidemaker_b200 as gm
    df = gm.PamTarget(pam, orientation, dtype).find_targets(recs, L)
    tp = gm.TargetProcessor(df, lsr=10, editdist=dist, knum=knum)
    tp.check_restriction_enzymes(["GGTCTC"])
    tp.find_unique_near_pam()
    tp.create_index(cfg)
    tp.get_neighbors(cfg)
    return df, tp


def _check_rows_vs_oracle(tp, L, metric, knum, n_rows, seed):
    """re-run `n_rows` sampled query rows on the CPU oracle against the whole distinct-guide table"""
    from guidemaker_b200._encode import encode_guides
    t = tp.targets
    g = encode_guides(t["target"], L)
    uniq = tp.nmslib_index.uniq
    ou, _ = O.unique_first_order(g)
    assert np.array_equal(uniq, ou), "distinct-guide table differs from first-occurrence order"
    rows = np.random.default_rng(seed).choice(len(g), size=min(n_rows, len(g)), replace=False)
    gi, gd = tp.nmslib_index.knn_packed(g[rows], knum)
    oi, od = O.c_knn(uniq, g[rows], L, metric, knum, threads=os.cpu_count())
    assert np.array_equal(gi, oi) and np.array_equal(gd, od), "kNN differs from the oracle"
    # the neighbour map holds exactly the query rows whose nearest OTHER guide is >= editdist away
    qmask = (~t["isseedduplicated"].to_numpy()) | (~t["hasrestrictionsite"].to_numpy().astype(bool))
    sel = rows[qmask[rows]]
    keep = od[qmask[rows], 1] >= tp.editdist
    for r, k_ in zip(sel[:200], keep[:200]):
        assert (t["target"].iat[int(r)] in tp.neighbors) == bool(k_)
    return g, uniq


def test_config3_tttv_leven_full_scale(capi, tmp_path):
    """BASELINE configs[2]: 6.3 Mb genome, Cas12a TTTV 5prime, 23-nt guides, --dtype leven --dist 2 (all ~5.1e4 targets)"""
    from guidemaker_b200.synth import config_genome
    recs = config_genome("c2_bacterial_6.3Mb")
    df, tp = _run_api(recs, "TTTV", "5prime", 23, "leven", _cfg(tmp_path))
    assert 3.0e4 < len(df) < 8.0e4 and (df["dtype"] == "leven").all()
    _check_rows_vs_oracle(tp, 23, 1, 5, 768, seed=3)
    sg, ss, sp, nf, nr = O.c_pam_scan(recs[0].seq.encode(), "TTTV", True, 23)
    assert len(df) == nf + nr and np.array_equal(df["start"].to_numpy()[:nf], ss[:nf])


def test_config4_multi_record_controls(capi, tmp_path):
    """BASELINE configs[3]: 12 Mb / 16 records, NGG, hamming, n = 100 000 controls: first round (10^6 random queries,
    seed 40) and the default-config round sequence on a smaller n, against the oracle"""
    from guidemaker_b200.synth import config_genome
    from guidemaker_b200._encode import encode_guides
    recs = config_genome("c4_yeast_12Mb")
    cfg1 = _cfg(tmp_path, min_hm=1, mult=(10, 100))
    df, tp = _run_api(recs, "NGG", "3prime", 20, "hamming", cfg1)
    assert df["seqid"].nunique() == 16 and 6.0e5 < len(df) < 1.1e6
    g, uniq = _check_rows_vs_oracle(tp, 20, 0, 5, 512, seed=4)
    # rows are grouped per record, forward block first (core.py:254-284)
    codes = pd.factorize(df["seqid"].astype(str))[0]                 # record order = order of first appearance
    assert (np.diff(codes) >= 0).all() and list(dict.fromkeys(df["seqid"].astype(str))) == [r.id for r in recs]
    for c in (0, 7, 15):
        st = df["strand"].to_numpy()[codes == c]
        assert st[0] and not st[-1] and (np.diff(st.astype(np.int8)) <= 0).all()
    np.random.seed(40)
    cmin, cmed, cdf = tp.get_control_seqs(recs, configpath=cfg1, length=20, n=100000)
    assert tp.ncontrolsearched == 1000000 and len(cdf) == 100000
    d = cdf["Hamming distance"].to_numpy()
    assert (np.diff(d) <= 0).all() and cmin == d.min()
    rows = np.random.default_rng(5).choice(len(cdf), size=600, replace=False)
    true = O.c_min_dist(uniq, encode_guides(cdf["Sequences"].to_numpy()[rows].tolist(), 20), 20, 0, threads=os.cpu_count())
    assert np.array_equal(true.astype(np.float64), d[rows])
    # the reference's default thresholds (MINIMUM_HMDIST 7, multiples 10..10000) on a small n: several rounds
    np.random.seed(41)
    cmin, cmed, cdf = tp.get_control_seqs(recs, configpath=_cfg(tmp_path, 5, (10, 100, 1000, 10000)), length=20, n=50)
    assert cmin >= 5 and tp.ncontrolsearched in (500, 5000, 50000)
    true = O.c_min_dist(uniq, encode_guides(cdf["Sequences"].tolist(), 20), 20, 0, threads=os.cpu_count())
    assert np.array_equal(true.astype(np.float64), cdf["Hamming distance"].to_numpy())


def test_config5_arabidopsis_scale_sample(capi, tmp_path):
    """BASELINE configs[4]: 120 Mb / 5 records, ~7.8e6 NGG targets, exact all-vs-all Hamming kNN on one GPU; 512 sampled
    rows against the oracle, size-independent properties on all rows"""
    from guidemaker_b200.synth import config_genome
    recs = config_genome("c5_arabidopsis_120Mb")
    df, tp = _run_api(recs, "NGG", "3prime", 20, "hamming", _cfg(tmp_path))
    assert 7.0e6 < len(df) < 8.5e6 and df["seqid"].nunique() == 5
    g, uniq = _check_rows_vs_oracle(tp, 20, 0, 5, 512, seed=5)
    nb = tp.neighbors
    idx, dist = nb.index_matrix(), nb.distance_matrix()
    assert (dist[:, 0] == 0).all() and (dist[:, 1] >= 2).all()
    key = dist.astype(np.int64) * (1 << 32) + idx
    assert (np.diff(key, axis=1) > 0).all() and idx.min() >= 0 and idx.max() < len(uniq)
    assert np.array_equal(uniq[idx[:, 0]], nb.codes)
    # the scan at full size: per-record oracle scan of one record
    sg, ss, sp, nf, nr = O.c_pam_scan(recs[2].seq.encode(), "NGG", False, 20)
    sub = df.loc[df["seqid"] == recs[2].id]
    assert len(sub) == nf + nr and np.array_equal(sub["start"].to_numpy(), ss)
    from guidemaker_b200._encode import encode_guides
    assert np.array_equal(encode_guides(sub["target"], 20), sg)
    assert np.array_equal(tp.targets["isseedduplicated"].to_numpy(), O.c_seed_dedup(g, 20, 10, False))


def test_fuzz_knn_bounded():
    """tools/fuzz_knn.py for ~20 s: random table sizes, query counts, k and L; K3b = K3a = oracle on every case"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_knn.py"), "20", "12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all identical to the oracle" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_knn_sharded_through_the_c_abi():
    """gm_comm_* + gm_knn_sharded (NCCL bound lazily inside the library): tools/comm_check.py under torchrun with as many
    ranks as there are GPUs (at most 2; one rank still runs ncclCommInitRank / ncclAllGather).  Every rank must get the
    full, oracle-exact table."""
    import torch
    world = max(1, min(2, torch.cuda.device_count()))
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "comm_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and all("COMM_OK %d" % i in r.stdout for i in range(world)), r.stdout[-3000:] + r.stderr[-3000:]


def test_public_api_under_nccl_ranks():
    """tools/api_nccl_check.py under torchrun with 2 ranks (skipped on a single-GPU box): sharded search, device-side gather
    and neighbour filter, rank 0's controls -- every rank ends with the oracle-exact tables"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "api_nccl_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "API_NCCL_OK 0" in r.stdout and "API_NCCL_OK 1" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
