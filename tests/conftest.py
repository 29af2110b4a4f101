import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def read_fasta_gz(path):
    """-> list of (id, sequence)"""
    out, name, chunks = [], None, []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(chunks)))
                name, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.strip())
    if name is not None:
        out.append((name, "".join(chunks)))
    return out


@pytest.fixture(scope="session")
def carsonella():
    recs = read_fasta_gz(os.path.join(GOLDEN, "carsonella.fa.gz"))
    assert len(recs) == 1 and len(recs[0][1]) == 159662
    return recs[0]


@pytest.fixture(scope="session")
def carsonella_ref():
    return load_npz("carsonella_ref.npz")


@pytest.fixture(scope="session")
def inline_ref():
    return load_npz("inline_ref.npz")


@pytest.fixture(scope="session")
def synthetic_ref():
    return load_npz("synthetic_ref.npz")


@pytest.fixture(scope="session")
def controls_ref():
    return load_npz("controls_ref.npz")


class Rec:
    """Duck-typed SeqRecord (Biopython is not in this image): .id, str(.seq), len()."""

    def __init__(self, id, seq):
        self.id, self.seq = id, seq

    def __len__(self):
        return len(self.seq)


class _OracleIndex:
    """Checker engine with the interface of guidemaker_b200._capi.Index (tests only)."""

    def __init__(self, uniq2bit, L, metric, **kw):
        from oracle import oracle as O
        self._O, self.u, self.L, self.metric, self.n = O, np.ascontiguousarray(uniq2bit, np.uint64), L, metric, len(uniq2bit)
        if self.n < 1:
            raise ValueError("empty guide table")

    def knn(self, q, k):
        return self._O.c_knn(self.u, q, self.L, self.metric, k)

    def min_dist(self, q):
        return self._O.c_min_dist(self.u, q, self.L, self.metric)

    def close(self):
        pass


class _OracleSession:
    """Checker stand-in for guidemaker_b200._capi.Session (tests only): the same row order, coordinates, text columns
    and chained stages, restated with the CPU oracle's primitives and plain numpy."""

    def __init__(self, buf, rec_start, pam, five_prime, L):
        from oracle import oracle as O
        if any(c not in O.IUPAC for c in pam) or not (1 <= L <= 27):
            raise ValueError("bad PAM / L")
        self._O, self.L, self.P, self.five_prime = O, int(L), len(pam), bool(five_prime)
        self.buf = np.ascontiguousarray(np.frombuffer(buf, np.uint8) if not isinstance(buf, np.ndarray) else buf)
        self.rec_start = np.asarray(rec_start, np.int64)
        g, gstart, p, nf, nr = O.c_pam_scan(self.buf.tobytes(), pam, five_prime, L)
        strand = np.zeros(nf + nr, dtype=bool)
        strand[:nf] = True
        rec = np.searchsorted(self.rec_start, gstart.astype(np.int64), side="right") - 1
        order = np.argsort(rec, kind="stable")               # per record: forward block, then reverse block
        self.g, self.p, self.strand, self.rec = g[order], p[order], strand[order], rec[order].astype(np.int32)
        self.gstart = gstart[order].astype(np.int64)
        self.start = (self.gstart - self.rec_start[self.rec]).astype(np.uint32)
        self.n_rows = len(self.g)

    def fetch_rows(self, want_pamcode=True):
        return self.g, self.start, (self.p if want_pamcode else None), self.rec, self.strand

    def pam_histogram(self):
        return np.bincount(self.p, minlength=1 << 16).astype(np.uint32)

    def pam_categories(self, lut):
        return np.asarray(lut, np.int8)[self.p]

    def fetch_text(self, width=30):
        from guidemaker_b200._encode import decode_matrix
        st, L, P, five = self.start.astype(np.int64), self.L, self.P, self.five_prime
        ms = np.where(self.strand, st - P, st + L) if five else np.where(self.strand, st + L, st - P)
        a = np.where(self.strand == five, ms - 3, ms + P - (width - 3))
        lens = (self.rec_start[1:] - self.rec_start[:-1] - 1)[self.rec]
        interior = (a >= 0) & (a + width <= lens)
        ctx = self._O.np_gather_windows(self.buf, np.where(interior, a + self.rec_start[self.rec], -1), ~self.strand, width)
        return decode_matrix(self.g, L), ctx, ~interior

    def seed_dedup(self, lsr):
        return self._O.c_seed_dedup(self.g, self.L, lsr, self.five_prime)

    def restriction(self, motifs):
        return self._O.c_restriction(self.g, self.L, list(motifs))

    def build_index(self, metric):
        first = self._O.c_first_occurrence(self.g)
        is_first = first == np.arange(len(self.g))
        uniq = np.ascontiguousarray(self.g[is_first])
        return _OracleIndex(uniq, self.L, metric), uniq, ((np.cumsum(is_first) - 1)[first]).astype(np.int32)

    def knn(self, index, qmask, k):
        return index.knn(np.ascontiguousarray(self.g[np.asarray(qmask).astype(bool)]), k)

    def close(self):
        pass


@pytest.fixture
def oracle_engine(monkeypatch):
    """Route the product's C-ABI calls to the CPU oracle so the HOST logic (frame assembly, masks,
    ordering, sharding) can be tested without a GPU.  Never used by the product itself."""
    from guidemaker_b200 import _capi
    from oracle import oracle as O

    def pam_scan(seq, pam, five_prime, L):
        if any(c not in O.IUPAC for c in pam) or not (1 <= L <= 27):
            raise ValueError("bad PAM / L")
        return O.c_pam_scan(bytes(seq), pam, five_prime, L)

    monkeypatch.setattr(_capi, "pam_scan", pam_scan)
    monkeypatch.setattr(_capi, "seed_dedup", lambda g, L, lsr, five: O.c_seed_dedup(g, L, lsr, five))
    monkeypatch.setattr(_capi, "first_occurrence", lambda g: O.c_first_occurrence(g))
    monkeypatch.setattr(_capi, "gather_windows", lambda seq, ws, rc, w: O.np_gather_windows(seq, ws, rc, w))
    monkeypatch.setattr(_capi, "restriction_scan", lambda g, L, motifs: O.c_restriction(g, L, motifs))
    monkeypatch.setattr(_capi, "Index", _OracleIndex)
    monkeypatch.setattr(_capi, "Session", _OracleSession)
    monkeypatch.setattr(_capi, "init", lambda device=None: None)
    return "oracle"


def _gpu_visible() -> bool:
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    """On a box without a GPU the `gpu` tests are skipped (not errors); on a GPU box nothing is skipped, so a missing or
    unloadable libgm_b200.so still fails loudly there."""
    if _gpu_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (gpu tests run on the B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cuda_engine():
    from guidemaker_b200 import _capi
    _capi.init(0)
    return "cuda"


@pytest.fixture
def config_yaml(tmp_path):
    """Same content as the reference's guidemaker/data/config_default.yaml."""
    import yaml
    p = tmp_path / "config_default.yaml"
    p.write_text(yaml.safe_dump({"NMSLIB": {"M": 16, "efc": 10, "post": 1, "ef": 9},
                                 "CONTROL": {"MINIMUM_HMDIST": 7, "CONTROL_SEARCH_MULTIPLE": [10, 100, 1000, 10000]},
                                 "MINIMUM_PROPORTION": 0.5}))
    return str(p)
