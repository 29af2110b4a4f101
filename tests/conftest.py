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
