"""Pins the CPU oracle (oracle/) against the reference: its own known-answer tests
(/root/reference/tests/test_core.py) and fixtures produced by running the reference's real
core.py (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

CASES = {"ngg3p20": ("NGG", False, 20, 10), "ngg5p20": ("NGG", True, 20, 10),
         "tttv5p23": ("TTTV", True, 23, 10), "nngrrt3p21": ("NNGRRT", False, 21, 12)}


def oracle_rows(seq, pam, five, L):
    g, s, p, nf, nr = O.c_pam_scan(seq.encode("latin-1"), pam, five, L)
    tgt = np.array([O.unpack(v, L) for v in g], dtype="S")
    pams = np.array([O.unpack(v, len(pam)) for v in p], dtype="S")
    strand = np.arange(len(g)) < nf
    return tgt, pams, s, strand, g


# ---- reference known-answer vectors (tests/test_core.py) ------------------------------------------

def test_ref_kat_find_targets_5p(inline_ref):            # test_core.py:41-46
    seq = inline_ref["t5p/seq"][0].decode()
    tgt, *_ = oracle_rows(seq, "NGG", True, 6)
    assert tgt[0] == b"ATGCAC" and tgt[1] == b"TAACAA"


def test_ref_kat_find_targets_3p(inline_ref):            # test_core.py:52-57 (sequence ends with ']')
    seq = inline_ref["t3p/seq"][0].decode()
    assert seq.endswith("]")
    tgt, *_ = oracle_rows(seq, "NGG", False, 6)
    assert tgt[0] == b"ATGATC" and tgt[1] == b"ATTAGA"


def test_ref_kat_fullgenome(carsonella):                 # test_core.py:59-65
    tgt, *_ = oracle_rows(carsonella[1], "NGG", True, 20)
    assert tgt[0] == b"AAATGGTACGTTATGTGTTA"


def test_ref_kat_keep_first():                           # test_core.py:67-102: rows 0,1 same guide -> 2 rows kept
    g = O.pack_many(["AAATGGTACGTTATGTGTTA", "AAATGGTACGTTATGTGTTA", "AACAGTAAAATGGTTTAATG"])
    assert O.c_seed_dedup(g, 20, 10, False).tolist() == [False, True, False]


def test_ref_kat_hamming_12():                           # test_core.py:116-126
    g = O.pack_many(["AAATGGTACGTTATGTGTTA", "AACAGTAAAATGGTTTAATG"])
    idx, dist = O.c_knn(g, g, 20, 0, 2)
    assert dist[0].tolist() == [0, 12] and idx[0].tolist() == [0, 1]


def test_ref_kat_levin_dist(inline_ref):                 # test_core.py:319-347
    seq = inline_ref["lev/seq"][0].decode()
    tgt, _, _, _, g = oracle_rows(seq, "NGG", False, 20)
    uniq, _ = O.unique_first_order(g)
    qi = [O.unpack(v, 20) for v in uniq].index("CTAGTCACTAGCTGACAGCA")
    _, dl = O.c_knn(uniq, uniq, 20, 1, 3)
    _, dh = O.c_knn(uniq, uniq, 20, 0, 3)
    assert dl[qi].tolist() == [0, 1, 2]
    assert dh[qi].tolist() == [0, 1, 16]


def test_md5_guide_name():                               # core.py:626-627, :905-906 (survey anchor)
    import hashlib
    assert hashlib.md5(b"AAATGGTACGTTATGTGTTA").hexdigest() == "927efc4c37b7736b1ade8656a598f9f5"


# ---- fixtures produced by the reference's real core.py ----------------------------------------------

@pytest.mark.parametrize("name", list(CASES))
def test_scan_and_dedup_vs_reference_carsonella(name, carsonella, carsonella_ref):
    pam, five, L, lsr = CASES[name]
    r = carsonella_ref
    tgt, pams, start, strand, g = oracle_rows(carsonella[1], pam, five, L)
    assert np.array_equal(tgt, r[name + "/target"])
    assert np.array_equal(pams, r[name + "/exact_pam"])
    assert np.array_equal(start, r[name + "/start"])
    assert np.array_equal(start + L, r[name + "/stop"])
    assert np.array_equal(strand, r[name + "/strand"])
    assert np.array_equal(O.c_seed_dedup(g, L, lsr, five), r[name + "/isseedduplicated"])


def _scan_records(ids, seqs, pam, five, L):
    parts = [oracle_rows(s, pam, five, L) for s in seqs]
    keep = [i for i, p in enumerate(parts) if len(p[0])]
    cat = lambda j: np.concatenate([parts[i][j] for i in keep])
    seqid = np.concatenate([np.full(len(parts[i][0]), ids[i], dtype="S16") for i in keep])
    return cat(0), cat(1), cat(2), cat(3), cat(4), seqid


@pytest.mark.parametrize("name", ["ngg3p", "ngg5p", "tttv", "nnagaaw", "yg10", "nggnorest"])
def test_full_path_vs_reference_synthetic(name, synthetic_ref):
    r = synthetic_ref
    pam, orient, dtype = (x.decode() for x in r[name + "/meta"])
    L, lsr, dist, knum = (int(x) for x in r[name + "/params"])
    five = orient == "5prime"
    ids = [x.decode() for x in r["rec_ids"]]; seqs = [x.decode() for x in r["rec_seqs"]]
    tgt, pams, start, strand, g, seqid = _scan_records(ids, seqs, pam, five, L)
    assert np.array_equal(tgt, r[name + "/target"])
    assert np.array_equal(pams, r[name + "/exact_pam"])
    assert np.array_equal(start, r[name + "/start"])
    assert np.array_equal(strand, r[name + "/strand"])
    assert np.array_equal(seqid, r[name + "/seqid"].astype("S16"))
    dup = O.c_seed_dedup(g, L, lsr, five)
    assert np.array_equal(dup, r[name + "/isseedduplicated"])
    # get_neighbors (core.py:495-523): query mask, exact kNN over distinct guides, dist[1] threshold
    if name + "/hasrestrictionsite" in r.files:
        qmask = (~dup) | (~r[name + "/hasrestrictionsite"])
    else:
        qmask = ~dup                                    # NaN == False is False
    uniq, _ = O.unique_first_order(g)
    metric = 0 if dtype == "hamming" else 1
    _, d = O.c_knn(uniq, g[qmask], L, metric, knum)
    keep = d[:, 1] >= dist
    got = {}
    for t, row in zip(tgt[qmask][keep], d[keep]):
        got[t] = [int(x) for x in row if x != 255]
    ref = {k: v.tolist() for k, v in zip(r[name + "/nb_keys"], r[name + "/nb_dist"])}
    assert list(got.keys()) == list(ref.keys())
    assert got == ref


@pytest.mark.parametrize("name", ["ngg3p20", "ngg5p20"])
def test_neighbors_vs_reference_carsonella(name, carsonella, carsonella_ref):
    pam, five, L, lsr = CASES[name]
    r = carsonella_ref
    tgt, _, _, _, g = oracle_rows(carsonella[1], pam, five, L)
    dup = O.c_seed_dedup(g, L, lsr, five)
    qmask = (~dup) | (~r[name + "/hasrestrictionsite"])
    uniq, _ = O.unique_first_order(g)
    idx, d = O.c_knn(uniq, g[qmask], L, 0, 3)
    keep = d[:, 1] >= 2
    got = {t: row.tolist() for t, row in zip(tgt[qmask][keep], d[keep])}
    ref = {k: v.tolist() for k, v in zip(r[name + "/nb_keys"], r[name + "/nb_dist"])}
    assert got == ref
    if name == "ngg3p20":                                # survey-derived anchors
        assert tgt[0] == b"TTTTCAAGAATAACACCATT"
        assert idx[0].tolist() == [0, 1121, 1144] and d[0].tolist() == [0, 6, 6]
        _, dall = O.c_knn(uniq, g, L, 0, 2)               # over ALL 3813 guides: 3811 have no neighbour closer than 2
        assert int((dall[:, 1] >= 2).sum()) == 3811 and int(dup.sum()) == 207


def test_control_distances_vs_reference(controls_ref):
    r = controls_ref
    for name, metric in (("ham", 0), ("lev", 1)):
        targets = O.pack_many([t.decode() for t in r[name + "/targets"]])
        uniq, _ = O.unique_first_order(targets)
        q = O.pack_many([s.decode() for s in r[name + "/seqs"]])
        d = O.c_min_dist(uniq, q, 20, metric)
        assert np.array_equal(d.astype(np.float64), r[name + "/dist"])


# ---- C oracle vs the literal Python restatement on random small inputs ------------------------------------

@pytest.mark.parametrize("seed", range(6))
def test_c_scan_equals_regex_restatement(seed):
    rng = np.random.default_rng(seed)
    letters = list("ACGTMRWSYKVHDBXN")
    for _ in range(12):
        P = int(rng.integers(2, 9)); L = int(rng.integers(10, 28)); five = bool(rng.integers(2))
        pam = "".join(rng.choice(letters, size=P, p=[0.12] * 4 + [0.52 / 12] * 12))
        n = int(rng.integers(0, 400))
        seq = "".join(rng.choice(list("ACGTNacgt"), size=n, p=[0.24, 0.24, 0.24, 0.24, 0.02, 0.005, 0.005, 0.005, 0.005]))
        rows = O.py_find_targets(seq, pam, five, L)
        tgt, pams, start, strand, _ = oracle_rows(seq, pam, five, L)
        assert [x[0].encode() for x in rows] == tgt.tolist()
        assert [x[1].encode() for x in rows] == pams.tolist()
        assert [x[2] for x in rows] == start.tolist()
        assert [x[4] for x in rows] == strand.tolist()


@pytest.mark.parametrize("seed", range(4))
def test_c_dedup_and_knn_equal_python(seed):
    rng = np.random.default_rng(100 + seed)
    L = int(rng.integers(10, 28)); lsr = int(rng.integers(0, L + 1)); five = bool(rng.integers(2))
    base = ["".join(rng.choice(list("ACGT"), size=L)) for _ in range(40)]
    seqs = [base[i] for i in rng.integers(0, 40, size=90)]
    for i in range(0, 90, 3):                      # near-duplicates -> ties and small distances
        s = list(seqs[i]); s[int(rng.integers(L))] = "ACGT"[int(rng.integers(4))]; seqs[i] = "".join(s)
    g = O.pack_many(seqs)
    assert np.array_equal(O.c_seed_dedup(g, L, lsr, five), O.py_seed_dedup(seqs, lsr, five))
    uniq, inv = O.unique_first_order(g)
    useqs = [O.unpack(v, L) for v in uniq]
    assert useqs == list(dict.fromkeys(seqs)) and [useqs[i] for i in inv] == seqs
    for metric in (0, 1):
        k = int(rng.integers(1, 8))
        idx, dist = O.c_knn(uniq, g[:25], L, metric, k)
        pi, pd_ = O.py_knn(useqs, seqs[:25], metric, k)
        assert idx.tolist() == pi and dist.tolist() == pd_
        assert O.c_min_dist(uniq, g[:25], L, metric).tolist() == [r[0] for r in pd_]


# ---- restriction-site flag (core.py:354-377) ------------------------------------------------------------------------
def test_restriction_flags_vs_reference_fixtures(carsonella, carsonella_ref, synthetic_ref):
    """hasrestrictionsite as written by the reference's check_restriction_enzymes (tests/golden/make_golden.py:
    Carsonella with ['NRAGCA'], the synthetic genome with ['GGTCTC', 'NGGTAB'])"""
    def motifs(enz):
        out = []
        for r in set(enz):
            out += [r.upper(), O.reverse_complement(r.upper())]
        return out
    n = 0
    for name, (pam, five, L, lsr) in CASES.items():
        key = name + "/hasrestrictionsite"
        if key in carsonella_ref.files:
            tgt = [t.decode() for t in carsonella_ref[name + "/target"]]
            assert np.array_equal(O.c_restriction(O.pack_many(tgt), L, motifs(["NRAGCA"])), carsonella_ref[key]), name
            assert np.array_equal(O.py_restriction(tgt, ["NRAGCA"]), carsonella_ref[key]), name
            n += 1
    r = synthetic_ref
    for key in [k for k in r.files if k.endswith("/hasrestrictionsite")]:
        name = key[: -len("/hasrestrictionsite")]
        tgt = [t.decode() for t in r[name + "/target"]]
        L = len(tgt[0])
        ref = r[key]
        if ref.dtype != bool or not ref.any():
            continue                                   # cases run without restriction enzymes
        assert np.array_equal(O.c_restriction(O.pack_many(tgt), L, motifs(["GGTCTC", "NGGTAB"])), ref), name
        n += 1
    assert n >= 2


def test_restriction_c_equals_literal_restatement():
    rng = np.random.default_rng(5)
    for L in (8, 20, 27):
        seqs = ["".join(rng.choice(list("ACGT"), L)) for _ in range(1500)]
        g = O.pack_many(seqs)
        for enz in (["NRAGCA"], ["GGTCTC", "NGGTAB"], ["A"], [""], ["ACGTACGTACGTACGTACGTACGTACGTACGT"], [], ["TTTV", "GAATTC", "CCWGG"],
                    ["RYKM", "SWBD"], ["ggtctc"]):
            motifs = []
            for r in set(enz):
                motifs += [r.upper(), O.reverse_complement(r.upper())]
            assert np.array_equal(O.c_restriction(g, L, motifs), O.py_restriction(seqs, enz)), (L, enz)


def test_tuned_cpu_knn_equals_checker():
    """oracle/gm_oracle.c gmo_knn_hamming_fast (the timed CPU arm: AVX-512 VPOPCNTD + query blocking) returns exactly
    what the scalar checker gmo_knn returns, ties included, on ragged sizes"""
    rng = np.random.default_rng(9)
    for n_t, n_q, L, k in ((1, 3, 20, 2), (17, 70, 5, 5), (4097, 129, 20, 1), (50003, 777, 27, 32), (8192, 64, 1, 3)):
        t = rng.integers(0, 1 << (2 * L), size=n_t, dtype=np.uint64)
        q = np.concatenate([t[: min(n_t, n_q // 2)], rng.integers(0, 1 << (2 * L), size=n_q - min(n_t, n_q // 2), dtype=np.uint64)])
        a, b = O.c_knn(t, q, L, 0, k), O.c_knn_hamming_fast(t, q, L, k)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (n_t, n_q, L, k)
