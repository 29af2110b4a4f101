"""GPU parity tests: every kernel of libgm_b200.so, called through the C ABI (ctypes), against the
CPU oracle on the same seeded inputs and against the committed reference fixtures.  Bit-exact."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi_mod():
    from guidemaker_b200 import _capi
    _capi.init(0)
    yield _capi
    _capi.knn_tune(8, 0, -1)
    _capi.knn_engine(1)


@pytest.fixture
def capi(capi_mod):
    """the INT-pipe engine (K3a XOR/POPC) for the Hamming tests of this section; K3b has its own section below"""
    capi_mod.knn_engine(0)
    yield capi_mod
    capi_mod.knn_engine(1)
    capi_mod.knn_tune(8, 0, -1)


def rand_guides(rng, n, L, n_base=None):
    """random guides with planted duplicates and near-duplicates (ties, small distances)"""
    n_base = n_base or max(n // 2, 1)
    base = rng.integers(0, 4, size=(n_base, L), dtype=np.uint64)
    rows = base[rng.integers(0, n_base, size=n)]
    mut = rng.random(n) < 0.5
    pos = rng.integers(0, L, size=n)
    rows[mut, pos[mut]] = rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint64)
    g = np.zeros(n, dtype=np.uint64)
    for i in range(L):
        g |= rows[:, i] << np.uint64(2 * i)
    return g


def rand_genome(rng, n, gc=0.5, n_frac=0.002, lower_frac=0.001):
    s = rng.choice(np.frombuffer(b"GCAT", np.uint8), size=n, p=[gc / 2, gc / 2, (1 - gc) / 2, (1 - gc) / 2])
    k = int(n * n_frac)
    if k:
        for st in rng.integers(0, max(n - 8, 1), size=max(k // 4, 1)):
            s[st:st + rng.integers(1, 8)] = ord("N")
    k = int(n * lower_frac)
    if k:
        idx = rng.integers(0, n, size=k)
        s[idx] = s[idx] + 32
    return s.tobytes()


# ---- K1 ---------------------------------------------------------------------------------------------------
CASES = {"ngg3p20": ("NGG", False, 20), "ngg5p20": ("NGG", True, 20), "tttv5p23": ("TTTV", True, 23), "nngrrt3p21": ("NNGRRT", False, 21)}


@pytest.mark.parametrize("name", list(CASES))
def test_scan_carsonella_vs_oracle_and_reference(capi, name, carsonella, carsonella_ref):
    pam, five, L = CASES[name]
    seq = carsonella[1].encode()
    g, s, p, nf, nr = capi.pam_scan(seq, pam, five, L)
    og, os_, op, onf, onr = O.c_pam_scan(seq, pam, five, L)
    assert (nf, nr) == (onf, onr)
    assert np.array_equal(g, og) and np.array_equal(s, os_) and np.array_equal(p, op)
    r = carsonella_ref
    assert np.array_equal(np.array([O.unpack(v, L) for v in g], dtype="S"), r[name + "/target"])
    assert np.array_equal(s, r[name + "/start"])


@pytest.mark.parametrize("seed", range(8))
def test_scan_random_iupac(capi, seed):
    rng = np.random.default_rng(seed)
    letters = list("ACGTMRWSYKVHDBXN")
    for _ in range(6):
        P = int(rng.integers(1, 9)); L = int(rng.integers(1, 28)); five = bool(rng.integers(2))
        pam = "".join(rng.choice(letters, size=P, p=[0.12] * 4 + [0.52 / 12] * 12))
        n = int(rng.choice([0, 1, 5, 31, 32, 33, 63, 64, 65, 200, 1000, 8191, 8192, 8193, 50000]))
        seq = rand_genome(rng, n, gc=float(rng.uniform(0.2, 0.8)), n_frac=0.01) if n else b""
        got = capi.pam_scan(seq, pam, five, L)
        exp = O.c_pam_scan(seq, pam, five, L)
        assert got[3:] == exp[3:], (pam, five, L, n)
        for a, b in zip(got[:3], exp[:3]):
            assert np.array_equal(a, b), (pam, five, L, n)


def test_scan_hits_at_record_ends(capi):
    L = 20
    core = "ACGTACGTACGTACGTACGT"
    for seq in ("CCA" + core + "TGG", core + "AGG", "CCT" + core, "GG" + core[:18] + "CC", "NGG" + core + "NGG" + "N" * 40 + "CCN" + core):
        for five in (False, True):
            got = capi.pam_scan(seq.encode(), "NGG", five, L)
            exp = O.c_pam_scan(seq.encode(), "NGG", five, L)
            assert got[3:] == exp[3:]
            for a, b in zip(got[:3], exp[:3]):
                assert np.array_equal(a, b)


def test_gather_windows_vs_numpy(capi):
    """the target_seq30 gather (gm_gather_windows): forward and reverse-complemented windows, IUPAC letters, lower
    case and N kept as Bio.Seq would, windows outside the buffer -> '?'"""
    rng = np.random.default_rng(3)
    seq = rng.choice(np.frombuffer(b"ACGTNRYKMSWBDHVXacgtnry-*", np.uint8), size=50021)
    ws = rng.integers(-40, len(seq) + 10, size=30011)
    ws[:6] = [0, len(seq) - 30, len(seq) - 29, -1, 1, len(seq)]
    rc = rng.random(len(ws)) < 0.5
    for width in (30, 1, 7):
        got = capi.gather_windows(seq, ws, rc, width)
        assert np.array_equal(got, O.np_gather_windows(seq, ws, rc, width)), width
    assert capi.gather_windows(seq, ws[:0], rc[:0], 30).shape == (0, 30)
    assert (capi.gather_windows(seq[:0], ws[:5], rc[:5], 30) == ord("?")).all()


def test_scan_rejects_bad_pam(capi):
    with pytest.raises(ValueError):
        capi.pam_scan(b"ACGT" * 10, "NGZ", False, 20)
    with pytest.raises(ValueError):
        capi.pam_scan(b"ACGT" * 10, "NGG", False, 28)


# ---- K2 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(4))
def test_seed_dedup_vs_oracle(capi, seed):
    rng = np.random.default_rng(10 + seed)
    for n in (1, 2, 1000, 100003):
        L = int(rng.integers(10, 28)); lsr = int(rng.integers(0, L + 1)); five = bool(rng.integers(2))
        g = rand_guides(rng, n, L, n_base=max(n // 8, 1))
        assert np.array_equal(capi.seed_dedup(g, L, lsr, five), O.c_seed_dedup(g, L, lsr, five))
        assert np.array_equal(capi.first_occurrence(g), O.c_first_occurrence(g))
    assert len(capi.seed_dedup(np.zeros(0, np.uint64), 20, 10, False)) == 0


def test_dedup_all_equal_and_all_distinct(capi):
    g = np.full(5000, 12345, np.uint64)
    d = capi.seed_dedup(g, 20, 10, True)
    assert not d[0] and d[1:].all()
    g = np.arange(5000, dtype=np.uint64) << np.uint64(20)
    assert not capi.seed_dedup(g, 20, 0, True).any()
    assert capi.seed_dedup(g, 20, 10, True)[1:].all()          # first 10 bases all 'A' -> every later row is a duplicate


# ---- K3a / K5 ----------------------------------------------------------------------------------------------
def check_knn(capi, targets, queries, L, metric, k):
    ix = capi.Index(targets, L, metric)
    idx, dist = ix.knn(queries, k)
    oi, od = O.c_knn(targets, queries, L, metric, k)
    assert np.array_equal(dist, od)
    assert np.array_equal(idx, oi)
    md = ix.min_dist(queries)
    assert np.array_equal(md, od[:, 0])
    ix.close()


@pytest.mark.parametrize("k", [1, 2, 3, 5, 20, 32])
def test_knn_hamming_carsonella_all_vs_all(capi, k, carsonella):
    g, *_ = O.c_pam_scan(carsonella[1].encode(), "NGG", False, 20)
    uniq, _ = O.unique_first_order(g)
    check_knn(capi, uniq, g, 20, 0, k)


@pytest.mark.parametrize("L", [1, 10, 16, 17, 20, 27])
def test_knn_hamming_lengths_and_ragged_sizes(capi, L):
    rng = np.random.default_rng(L)
    for n_t, n_q in ((1, 1), (3, 7), (1023, 1025), (1024, 4096), (1025, 100), (5000, 3333)):
        t, _ = O.unique_first_order(rand_guides(rng, n_t, L))
        q = rand_guides(rng, n_q, L)
        check_knn(capi, t, q, L, 0, 5)


def test_knn_fewer_targets_than_k(capi):
    rng = np.random.default_rng(3)
    t, _ = O.unique_first_order(rand_guides(rng, 4, 20))
    ix = capi.Index(t, 20, 0)
    idx, dist = ix.knn(t, 8)
    assert (idx[:, len(t):] == -1).all() and (dist[:, len(t):] == 255).all()
    oi, od = O.c_knn(t, t, 20, 0, 8)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


def test_knn_ties_resolved_by_index(capi):
    # every target at distance exactly 1 from the query: ties everywhere -> lowest indices win
    L = 20
    q = O.pack("A" * L)
    t = []
    for pos in range(L):
        for b in "CGT":
            s = ["A"] * L; s[pos] = b; t.append("".join(s))
    t = O.pack_many(t)
    ix = capi.Index(t, L, 0)
    idx, dist = ix.knn(np.array([q], np.uint64), 10)
    assert idx[0].tolist() == list(range(10)) and (dist[0] == 1).all()
    check_knn(capi, np.tile(t, 1)[::-1].copy(), np.array([q] * 300, np.uint64), L, 0, 7)


@pytest.mark.parametrize("tune", [(4, 0, -1), (8, 1, 0), (8, 3, 0), (4, 7, 2048), (8, 64, 1024), (8, 0, 4096)])
def test_knn_hamming_tuning_variants_agree(capi, tune):
    """queries per thread, target splits and warm start must not change a single bit"""
    rng = np.random.default_rng(99)
    t, _ = O.unique_first_order(rand_guides(rng, 70000, 20, n_base=60000))
    q = rand_guides(rng, 3000, 20)
    capi.knn_tune(*tune)
    try:
        check_knn(capi, t, q, 20, 0, 6)
    finally:
        capi.knn_tune(8, 0, -1)


# ---- K3b: tcgen05 one-hot GEMM engine --------------------------------------------------------------------------
@pytest.fixture
def tc_engine(capi_mod):
    capi_mod.knn_engine(1)
    yield capi_mod
    capi_mod.knn_engine(1)
    capi_mod.knn_tune(8, 0, -1)


@pytest.mark.parametrize("L", [1, 4, 10, 16, 17, 20, 24, 27])
def test_knn_tc_lengths_and_ragged_sizes(tc_engine, L):
    rng = np.random.default_rng(700 + L)
    for n_t, n_q in ((1, 1), (3, 7), (255, 257), (1023, 1025), (1024, 300), (5000, 3333)):
        t, _ = O.unique_first_order(rand_guides(rng, n_t, L))
        q = rand_guides(rng, n_q, L)
        check_knn(tc_engine, t, q, L, 0, 5)


@pytest.mark.parametrize("k", [1, 3, 20])
def test_knn_tc_carsonella_all_vs_all(tc_engine, k, carsonella):
    g, *_ = O.c_pam_scan(carsonella[1].encode(), "NGG", False, 20)
    uniq, _ = O.unique_first_order(g)
    check_knn(tc_engine, uniq, g, 20, 0, k)


@pytest.mark.parametrize("tune", [(8, 1, 0), (8, 3, 0), (4, 7, 2048), (8, 64, 1024), (8, 0, 4096)])
def test_knn_tc_splits_and_warm_start_agree(tc_engine, tune):
    rng = np.random.default_rng(98)
    t, _ = O.unique_first_order(rand_guides(rng, 70000, 20, n_base=60000))
    q = rand_guides(rng, 3000, 20)
    tc_engine.knn_tune(*tune)
    check_knn(tc_engine, t, q, 20, 0, 6)


def test_knn_tc_ties_and_duplicates(tc_engine):
    L = 20
    q = O.pack("A" * L)
    t = []
    for pos in range(L):
        for b in "CGT":
            s = ["A"] * L; s[pos] = b; t.append("".join(s))
    t = O.pack_many(t)
    check_knn(tc_engine, t, np.array([q] * 300, np.uint64), L, 0, 7)
    check_knn(tc_engine, t[::-1].copy(), t, L, 0, 10)
    same = np.full(3000, O.pack("ACGTACGTACGTACGTACGT"), np.uint64)        # all-equal queries against a tiny table
    check_knn(tc_engine, t[:5], same, L, 0, 8)


def test_knn_tc_equals_popc_engine_at_scale(capi_mod):
    """both engines, 200k x 200k: byte-identical outputs"""
    capi = capi_mod
    rng = np.random.default_rng(11)
    t, _ = O.unique_first_order(rand_guides(rng, 200000, 20, n_base=190000))
    q = rand_guides(rng, 200000, 20)
    ix = capi.Index(t, 20, 0)
    capi.knn_engine(0)
    a = ix.knn(q, 5)
    capi.knn_engine(1)
    b = ix.knn(q, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rows = rng.integers(0, len(q), size=200)
    oi, od = O.c_knn(t, q[rows], 20, 0, 5)
    assert np.array_equal(b[0][rows], oi) and np.array_equal(b[1][rows], od)


def test_knn_tc_tail_wave_launch(capi_mod):
    """160 query tiles on 148 SMs: the 12 tiles behind the full wave run as a second launch cut into target splits
    (knn.cu, two-tier launch).  Same bits as K3a, and as the oracle on a sample."""
    capi = capi_mod
    rng = np.random.default_rng(12)
    t, _ = O.unique_first_order(rand_guides(rng, 150000, 20, n_base=140000))
    q = rand_guides(rng, 160 * 512 - 77, 20)
    ix = capi.Index(t, 20, 0)
    capi.knn_tune(8, 0, -1)
    capi.knn_engine(0)
    a = ix.knn(q, 4)
    capi.knn_engine(1)
    b = ix.knn(q, 4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rows = np.concatenate([rng.integers(0, len(q), size=100), np.arange(len(q) - 100, len(q))])   # incl. the tail tiles
    oi, od = O.c_knn(t, q[rows], 20, 0, 4)
    assert np.array_equal(b[0][rows], oi) and np.array_equal(b[1][rows], od)
    d1 = ix.min_dist(q[-5000:])
    assert np.array_equal(d1, b[1][-5000:, 0])


@pytest.mark.parametrize("L,k,splits", [(9, 1, 0), (9, 5, 0), (11, 32, 0), (20, 5, 5), (27, 3, 2), (2, 4, 0)])
def test_knn_tc_neighbourhood_warm_start(tc_engine, L, k, splits):
    """tables large enough for K3b's default warm start (warm.cu: bounds from the guides around the query's rank in
    sorted copies of the table), short guides so that ties at the k-th distance are everywhere; dense queries take the
    shared-memory staging path, sparse ones read their windows from L2; also under explicit target splits"""
    rng = np.random.default_rng(900 + L)
    if 4 ** L < 400000:
        t = rng.permutation(4 ** L).astype(np.uint64)[: min(4 ** L, 120000)]
        if len(t) < 40000:                                  # L = 2: too small for the window path, must still be exact
            t = rng.permutation(t)
    else:
        t, _ = O.unique_first_order(rand_guides(rng, 120000, L, n_base=100000))
    tc_engine.knn_tune(8, splits, -1)
    for q in (np.concatenate([t[:20000], rand_guides(rng, 20000, L)]), rand_guides(rng, 500, L)):
        ix = tc_engine.Index(t, L, 0)
        idx, dist = ix.knn(q, k)
        rows = rng.integers(0, len(q), size=min(len(q), 400))
        oi, od = O.c_knn(t, q[rows], L, 0, k)
        assert np.array_equal(idx[rows], oi) and np.array_equal(dist[rows], od)
        ix.tune(engine=0)                                   # all rows against the other engine
        ref = ix.knn(q, k)
        assert np.array_equal(idx, ref[0]) and np.array_equal(dist, ref[1])
        ix.close()


# ---- K6 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [1, 7, 10, 20, 27])
def test_restriction_scan_vs_oracle(capi_mod, L):
    """IUPAC motif search on packed guides (check_restriction_enzymes, core.py:354-377) against the C oracle, and
    against the literal regex restatement on a sample"""
    capi = capi_mod
    rng = np.random.default_rng(100 + L)
    g = rand_guides(rng, 20011, L)
    letters = list(O.IUPAC)
    cases = [[], [""], ["A"], ["N"], ["GGTCTC", "NGGTAB"], ["NRAGCA"], ["ACGT" * 8], ["T" * L], ["N" * (L + 1)]]
    for _ in range(12):
        cases.append(["".join(rng.choice(letters, size=int(rng.integers(1, 9)))) for _ in range(int(rng.integers(1, 5)))])
    cases.append(["".join(rng.choice(letters, size=int(rng.integers(2, 7)))) for _ in range(40)])     # > 64 motifs: two batches
    for enz in cases:
        motifs = []
        for r in set(enz):
            motifs += [r, O.reverse_complement(r)]
        got = capi.restriction_scan(g, L, motifs)
        assert np.array_equal(got, O.c_restriction(g, L, motifs)), enz
        if sum(4 ** sum(c in "NX" for c in m) * 3 ** sum(c in "VHDB" for c in m) for m in motifs) < 5000:
            sample = g[:400]
            assert np.array_equal(got[:400], O.py_restriction([O.unpack(v, L) for v in sample], enz)), enz
    assert len(capi.restriction_scan(g[:0], L, ["GAATTC"])) == 0
    with pytest.raises(KeyError):
        capi.restriction_scan(g, L, ["GAZTTC"])


# ---- K4 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [10, 20, 23, 27])
def test_knn_leven_vs_oracle(capi, L):
    rng = np.random.default_rng(40 + L)
    t, _ = O.unique_first_order(rand_guides(rng, 3000, L))
    # queries: targets with an insertion/deletion (shifted) so edit distance < hamming distance
    q = rand_guides(rng, 700, L)
    sh = t[rng.integers(0, len(t), size=300)]
    mask = np.uint64((1 << (2 * L)) - 1)
    q = np.concatenate([q, (sh << np.uint64(2)) & mask, sh >> np.uint64(2)])
    check_knn(capi, t, q, L, 1, 4)


def test_knn_leven_inline_reference_vector(capi, inline_ref):
    """tests/test_core.py:319-347 -- leven [0,1,2], hamming [0,1,16]"""
    seq = inline_ref["lev/seq"][0]
    g, *_ = capi.pam_scan(seq, "NGG", False, 20)
    uniq = g[np.sort(np.unique(g, return_index=True)[1])]
    qi = [O.unpack(v, 20) for v in uniq].index("CTAGTCACTAGCTGACAGCA")
    assert capi.Index(uniq, 20, 1).knn(uniq, 3)[1][qi].tolist() == [0, 1, 2]
    assert capi.Index(uniq, 20, 0).knn(uniq, 3)[1][qi].tolist() == [0, 1, 16]


def test_knn_leven_warm_and_splits(capi):
    rng = np.random.default_rng(5)
    t, _ = O.unique_first_order(rand_guides(rng, 40000, 23, n_base=30000))
    q = rand_guides(rng, 600, 23)
    for tune in ((8, 5, 2048), (8, 1, 0)):
        capi.knn_tune(*tune)
        try:
            check_knn(capi, t, q, 23, 1, 3)
        finally:
            capi.knn_tune(8, 0, -1)


@pytest.mark.parametrize("L,n,k,tune", [(20, 50000, 4, (8, 0, -1)), (8, 30000, 3, (8, 0, -1)), (27, 20000, 5, (8, 0, -1)),
                                         (23, 40000, 2, (8, 5, 2048)), (20, 30000, 6, (4, 3, 0)), (9, 4096, 1, (8, 0, -1))])
def test_knn_leven_prefix_sharing(capi_mod, L, n, k, tune):
    """K4p (engine 1, the default for Levenshtein): the scan over the prefix-sorted copy of the table that resumes Myers'
    recurrence from the state shared with the previous target; targets arrive out of index order, so ties at the k-th
    distance exercise the full-key lists.  Against the oracle and against the plain kernel (engine 0)."""
    rng = np.random.default_rng(60 + L + k)
    if 4 ** L < 200000:
        t = rng.permutation(4 ** L).astype(np.uint64)[:n]
    else:
        t, _ = O.unique_first_order(rand_guides(rng, n, L, n_base=n // 2))
    mask = np.uint64((1 << (2 * L)) - 1)
    sh = t[rng.integers(0, len(t), size=200)]
    q = np.concatenate([rand_guides(rng, 400, L), (sh << np.uint64(2)) & mask, sh >> np.uint64(2), t[:100]])
    capi_mod.knn_engine(1)
    capi_mod.knn_tune(*tune)
    try:
        ix = capi_mod.Index(t, L, 1)
        idx, dist = ix.knn(q, k)
        oi, od = O.c_knn(t, q, L, 1, k)
        assert np.array_equal(dist, od)
        assert np.array_equal(idx, oi)
        assert np.array_equal(ix.min_dist(q), od[:, 0])
        ix.tune(engine=0)
        ref = ix.knn(q, k)
        assert np.array_equal(idx, ref[0]) and np.array_equal(dist, ref[1])
        ix.close()
    finally:
        capi_mod.knn_tune(8, 0, -1)


# ---- errors ---------------------------------------------------------------------------------------------------
def test_argument_errors(capi):
    with pytest.raises(ValueError):
        capi.Index(np.zeros(0, np.uint64), 20, 0)
    with pytest.raises(ValueError):
        capi.Index(np.zeros(4, np.uint64), 28, 0)
    ix = capi.Index(np.arange(4, dtype=np.uint64), 20, 0)
    with pytest.raises(ValueError):
        ix.knn(np.zeros(3, np.uint64), 33)
    assert ix.knn(np.zeros(0, np.uint64), 3)[0].shape == (0, 3)


# ---- BASELINE-size property checks ---------------------------------------------------------------------------
@pytest.mark.parametrize("engine", [1, 0])
def test_knn_hamming_config2_scale_properties(capi_mod, engine):
    """configs[1] scale (6.3 Mb, 66 % GC, NGG 3prime, L=20): sample-checked against the oracle plus
    size-independent properties over all rows."""
    capi = capi_mod
    capi.knn_engine(engine)
    rng = np.random.default_rng(2)
    seq = rand_genome(rng, 6_300_000, gc=0.66, n_frac=0.0005, lower_frac=0)
    g, s, p, nf, nr = capi.pam_scan(seq, "NGG", False, 20)
    assert 1.0e6 < len(g) < 1.8e6
    first = capi.first_occurrence(g)
    is_first = first == np.arange(len(g))
    uniq = np.ascontiguousarray(g[is_first])
    row2uniq = (np.cumsum(is_first) - 1)[first]
    ix = capi.Index(uniq, 20, 0)
    k = 5
    idx, dist = ix.knn(g, k)
    # (1) the self hit comes first, at distance 0 and at the query's own index
    assert (dist[:, 0] == 0).all() and np.array_equal(idx[:, 0], row2uniq.astype(np.int32))
    # (2) rows are sorted by (distance, index), indices distinct and in range
    key = dist.astype(np.int64) * (1 << 32) + idx
    assert (np.diff(key, axis=1) > 0).all() and idx.min() >= 0 and idx.max() < len(uniq)
    # (3) reported distances are the true distances of the reported pairs
    def ham(a, b):
        x = a ^ b
        x = (x | (x >> np.uint64(1))) & np.uint64(0x5555555555555555)
        return np.array([bin(int(v)).count("1") for v in x])
    rows = rng.integers(0, len(g), size=2000)
    for j in range(k):
        assert np.array_equal(ham(g[rows], uniq[idx[rows, j]]), dist[rows, j])
    # (4) oracle on a random sample of queries against the whole table
    rows = rng.integers(0, len(g), size=512)
    oi, od = O.c_knn(uniq, g[rows], 20, 0, k)
    assert np.array_equal(idx[rows], oi) and np.array_equal(dist[rows], od)
    # (5) min-distance query == first column
    assert np.array_equal(ix.min_dist(g[:100000]), dist[:100000, 0])
    # (6) scan + dedupe agree with the oracle at full size
    og, os_, op, onf, onr = O.c_pam_scan(seq, "NGG", False, 20)
    assert np.array_equal(g, og) and np.array_equal(s, os_) and (nf, nr) == (onf, onr)
    assert np.array_equal(capi.seed_dedup(g, 20, 10, False), O.c_seed_dedup(g, 20, 10, False))


def test_per_handle_engine_and_tuning(capi_mod):
    """two indices with different engines / tunings in one process do not interfere (gm_index_tune), and both return the
    same bits as the process-default configuration"""
    rng = np.random.default_rng(31)
    t, _ = O.unique_first_order(rand_guides(rng, 60000, 20))
    q = rand_guides(rng, 3000, 20)
    capi_mod.knn_engine(1)
    a, b = capi_mod.Index(t, 20, 0), capi_mod.Index(t, 20, 0)
    a.tune(engine=0, queries_per_thread=4, splits=3, warm_sample=0)      # K3a, explicit splits, no warm start
    b.tune(engine=1)                                                     # K3b
    ra, rb = a.knn(q, 5), b.knn(q, 5)
    ra2 = a.knn(q, 5)                                                    # a's settings survive b's call
    oi, od = O.c_knn(t, q, 20, 0, 5)
    for r in (ra, rb, ra2):
        assert np.array_equal(r[0], oi) and np.array_equal(r[1], od)
    with pytest.raises(ValueError):
        a.tune(engine=2)
    a.close(); b.close()
