"""CFD off-target scores for the guide table: ``guidemaker.core.cfd_score`` (core.py:1129-1148) and
``guidemaker.cfd_score_calculator`` (cfd_score_calculator.py:29-85).

``calc_cfd`` is the scalar definition (kept for callers and as the checker's twin); ``cfd_score(df)`` scores the whole
table at once on the GPU (``gm_cfd_scores``: double precision, the reference's multiplication order, so every score --
and its ``str()`` in the 'CFD Similar Guides' column -- equals the reference's)."""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

import numpy as np

from . import _capi
from ._encode import encode_guides

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cfd_mm_scores.json")
_TABLE = None


def mm_table() -> np.ndarray:
    """[rna base A,C,G,U][dna base A,C,G,T][position - 1] mismatch weights (Doench et al. 2016)"""
    global _TABLE
    if _TABLE is None:
        _TABLE = np.ascontiguousarray(np.array(json.load(open(_DATA))["mm"], dtype=np.float64))
        assert _TABLE.shape == (4, 4, 20)
    return _TABLE


def get_mm_pam_scores() -> Tuple[Dict, Dict]:
    """the mismatch weights in the reference's dict form ('rA:dC,7' -> weight); PAM weights are not used by GuideMaker
    (cfd_score_calculator.py:4-7) and are returned empty"""
    t = mm_table()
    mm = {}
    for ri, r in enumerate("ACGU"):
        for di, d in enumerate("ACGT"):
            if "ACGT"[3 - di] == ("T" if r == "U" else r):        # a match, not a mismatch pair
                continue
            for p in range(20):
                mm["r%s:d%s,%d" % (r, d, p + 1)] = float(t[ri, di, p])
    return mm, {}


def calc_cfd(wt: str, off: str, mm_scores=None) -> float:
    """CFD score of one guide / off-target pair (cfd_score_calculator.py:62-85)"""
    assert len(wt) == len(off), "The lengths wt and off differ: wt = {}, off = {}".format(str(len(wt)), str(len(off)))
    guidelen = len(wt)
    if mm_scores is None:
        mm_scores, _ = get_mm_pam_scores()
    score = 1.
    off = off.upper().replace('T', 'U')
    wt = wt.upper().replace('T', 'U')
    basecomp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'U': 'A'}
    for i, sl in enumerate(off):
        if (guidelen - 20 - i) <= 0:
            if wt[i] != sl:
                score *= mm_scores['r' + wt[i] + ':d' + basecomp[sl] + ',' + str(20 + i + 1 - guidelen)]
    return score


def cfd_scores_packed(wt2bit: np.ndarray, off2bit: np.ndarray, L: int) -> np.ndarray:
    """(n,) guides x (n, k) off-targets (guide2bit) -> (n, k) float64 scores, on the GPU"""
    _capi.init()
    wt = np.ascontiguousarray(wt2bit, np.uint64)
    off = np.ascontiguousarray(off2bit, np.uint64)
    n, k = off.shape
    out = np.empty((n, k), np.float64)
    t = mm_table()
    _capi._check(_capi.load_library().gm_cfd_scores(_capi._p(wt), _capi._p(off), n, k, int(L), _capi._p(t), _capi._p(out)), "gm_cfd_scores")
    return out


def cfd_score(df):
    """Adds 'CFD Similar Guides' (list of str(score), one per similar guide) and 'Max CFD' (core.py:1129-1148)."""
    guides = df['Guide sequence'].astype(str).tolist()
    sims = [s.split(';') for s in df['Similar guides'].astype(str)]
    if len(guides) == 0:
        df['CFD Similar Guides'] = []
        df['Max CFD'] = []
        return df
    L = len(guides[0])
    k = max(len(s) for s in sims)
    uniform = all(len(g) == L for g in guides) and all(len(x) == L for s in sims for x in s) and L <= _capi.MAX_L
    if uniform:
        wt = encode_guides(guides, L)
        off = np.repeat(wt[:, None], k, axis=1)                    # padding: the guide itself (score 1.0, dropped below)
        flat = encode_guides([x for s in sims for x in s], L)
        counts = np.array([len(s) for s in sims])
        rows = np.repeat(np.arange(len(sims)), counts)
        cols = np.arange(len(flat)) - np.repeat(np.cumsum(counts) - counts, counts)
        off[rows, cols] = flat
        scores = cfd_scores_packed(wt, off, L)
        cfd_lists = [[str(float(x)) for x in scores[i, :c]] for i, c in enumerate(counts)]
    else:                                                          # ragged or non-ACGT input: the scalar definition
        mm, _ = get_mm_pam_scores()
        cfd_lists = [[str(calc_cfd(g, x, mm_scores=mm)) for x in s] for g, s in zip(guides, sims)]
    df['CFD Similar Guides'] = cfd_lists
    df['Max CFD'] = [max(float(x) for x in lst) for lst in cfd_lists]
    return df
