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
    monkeypatch.setattr(_capi, "init", lambda device=None: None)
    return "oracle"


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
