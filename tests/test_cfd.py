"""CFD off-target scores (second half): the scalar definition against the reference's known answer
and its weight table, the GPU kernel against the scalar definition (bit-exact doubles), and ``cfd_score(df)``."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from guidemaker_b200 import cfd
from guidemaker_b200._encode import encode_guides


def test_calc_cfd_reference_known_answer():
    """tests/test_core.py:265-267"""
    result = cfd.calc_cfd("GCATGCACAGCTAGCATGCATGCAGCT", "GCATGCACAGCTAGCATGCATGCAGCG")
    assert abs(result - 0.176470588) < 0.0001
    assert cfd.calc_cfd("ACGTACGTACGTACGTACGT", "ACGTACGTACGTACGTACGT") == 1.0
    mm, pam = cfd.get_mm_pam_scores()
    assert len(mm) == 240 and mm["rU:dT,12"] == 0.8 and mm["rG:dA,14"] == 0.26666666699999997
    with pytest.raises(AssertionError):
        cfd.calc_cfd("ACGT", "ACG")


def test_weight_table_equals_reference_copy():
    ref = "/root/reference/guidemaker/data/cfd_data.json"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box")
    assert json.load(open(ref))["mm"] == cfd.get_mm_pam_scores()[0]


def _rand_seqs(rng, n, L):
    return ["".join(rng.choice(list("ACGT"), size=L)) for _ in range(n)]


@pytest.mark.gpu
@pytest.mark.parametrize("L", [10, 19, 20, 23, 27])
def test_cfd_kernel_equals_scalar_definition(cuda_engine, L):
    rng = np.random.default_rng(L)
    wt = _rand_seqs(rng, 300, L)
    off = []
    for w in wt:
        row = []
        for j in range(6):
            s = list(w)
            for p in rng.integers(0, L, size=j):            # 0..5 mismatches
                s[p] = rng.choice(list("ACGT"))
            row.append("".join(s))
        off.append(row)
    got = cfd.cfd_scores_packed(encode_guides(wt, L), np.array([encode_guides(r, L) for r in off]), L)
    mm, _ = cfd.get_mm_pam_scores()
    want = np.array([[cfd.calc_cfd(w, o, mm_scores=mm) for o in row] for w, row in zip(wt, off)])
    assert np.array_equal(got, want)                        # same doubles: same products in the same order


@pytest.mark.gpu
def test_cfd_score_frame(cuda_engine):
    """core.py:1129-1148: list of str(score) per similar guide, and their maximum"""
    rng = np.random.default_rng(1)
    g = _rand_seqs(rng, 40, 20)
    sims = [";".join([x] + _rand_seqs(rng, 3, 20)) for x in g]
    df = pd.DataFrame({"Guide sequence": g, "Similar guides": sims})
    out = cfd.cfd_score(df.copy())
    mm, _ = cfd.get_mm_pam_scores()
    for i in range(len(g)):
        want = [str(cfd.calc_cfd(g[i], s, mm_scores=mm)) for s in sims[i].split(";")]
        assert out["CFD Similar Guides"].iloc[i] == want
        assert out["Max CFD"].iloc[i] == max(float(x) for x in want) == 1.0       # the self hit scores 1
    # ragged rows (fewer neighbours than k) and the scalar fallback give the same answer
    df2 = pd.DataFrame({"Guide sequence": g[:3], "Similar guides": [g[0], g[1] + ";" + g[2], ";".join(g[:3])]})
    out2 = cfd.cfd_score(df2.copy())
    assert [len(x) for x in out2["CFD Similar Guides"]] == [1, 2, 3]
    assert out2["CFD Similar Guides"].iloc[1] == [str(cfd.calc_cfd(g[1], s, mm_scores=mm)) for s in (g[1], g[2])]
