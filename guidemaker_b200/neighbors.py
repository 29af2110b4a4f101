"""Array-backed replacements for two reference objects on the hot path:

* ``ExactIndex``  -- stands where ``TargetProcessor.nmslib_index`` (an nmslib HNSW index,
  /root/reference/guidemaker/core.py:451-467) stood.  It keeps nmslib's query protocol
  (``knnQueryBatch`` / ``setQueryTimeParams``) so external callers keep working, but the search is
  the exact brute force of ``libgm_b200.so``.
* ``NeighborMap`` -- stands where the ``self.neighbors`` dict-of-dicts built by the Python loop of
  core.py:504-523 stood; same ``neighbors[seq]["neighbors"]["dist"|"seqs"]`` access pattern, but
  backed by the arrays that come off the GPU (no per-guide Python objects until asked for).
"""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np

from . import _capi
from ._encode import decode_guides, encode_guides

_ONEHOT = {"1 0 0 0": "A", "0 1 0 0": "C", "0 0 1 0": "G", "0 0 0 1": "T"}


def _from_one_hot(s: str) -> str:
    """Inverse of TargetProcessor._one_hot_encode (core.py:379-386)."""
    toks = s.split(" ")
    return "".join(_ONEHOT[" ".join(toks[i:i + 4])] for i in range(0, len(toks), 4))


class ExactIndex:
    """Exact kNN index over the distinct guides (first-occurrence order)."""

    def __init__(self, uniq2bit: np.ndarray, L: int, metric: int, engine=None):
        self.uniq = np.ascontiguousarray(uniq2bit, np.uint64)
        self.L, self.metric = int(L), int(metric)
        # `engine` exists so that the multi-rank plumbing can be unit-tested on CPU boxes with an
        # injected checker; the product never passes it and always gets the CUDA engine.
        self._engine = engine if engine is not None else _capi.Index(self.uniq, self.L, self.metric)

    def __len__(self):
        return len(self.uniq)

    # ---- packed fast path used by TargetProcessor
    def knn_packed(self, q2bit: np.ndarray, k: int):
        """-> (idx int32 [q,k], dist uint8 [q,k]); true mismatch/edit counts; pads idx=-1, dist=255."""
        return self._engine.knn(np.ascontiguousarray(q2bit, np.uint64), int(k))

    def min_dist_packed(self, q2bit: np.ndarray) -> np.ndarray:
        return self._engine.min_dist(np.ascontiguousarray(q2bit, np.uint64))

    # ---- nmslib protocol (core.py:501-503, :603)
    def setQueryTimeParams(self, params=None):      # HNSW efSearch: an exact search has no tunables
        return None

    def knnQueryBatch(self, queries, k: int = 10, num_threads: int = 0):
        """list[str] -> list[(ids int32[<=k], dists int32[<=k])], ascending (distance, id).

        For the hamming index queries may be one-hot strings (nmslib bit_hamming input, as
        produced by ``_one_hot_encode``) or plain DNA; distances are reported DOUBLED, as nmslib
        reports them over the one-hot encoding (callers halve them, core.py:512-514, :613)."""
        queries = list(queries)
        if self.metric == _capi.METRIC_HAMMING:
            queries = [_from_one_hot(s) if " " in s else s for s in queries]
        q = encode_guides(queries, self.L) if queries else np.zeros(0, np.uint64)
        idx, dist = self.knn_packed(q, min(int(k), _capi.MAX_K))
        scale = 2 if self.metric == _capi.METRIC_HAMMING else 1
        out = []
        for i in range(len(q)):
            m = idx[i] >= 0
            out.append((idx[i][m].astype(np.int32), dist[i][m].astype(np.int32) * scale))
        return out


class NeighborMap(Mapping):
    """``neighbors[seq] -> {"target": seq, "neighbors": {"seqs": [...], "dist": [...]}}``."""

    def __init__(self, qcodes: np.ndarray, idx: np.ndarray, dist: np.ndarray, uniq: np.ndarray, L: int, group=None, rows=None, final=False):
        """qcodes/idx/dist: query rows in row order; `rows` (ascending indices, default all) selects the kept ones.
        A dict keyed by the guide string keeps the first row of every distinct guide, in order of first appearance
        (later rows carry identical values).  `group[i]` (one per kept row) = any integer id < len(uniq) that is equal
        for equal guides (e.g. the row's index in the distinct-guide table): with it the dedupe is a linear scatter;
        without it, one stable sort.  The big arrays are gathered once, with the composed index."""
        if final:                                            # already filtered and one row per guide (gm_session_neighbors)
            self.codes, self.idx, self.dist = np.ascontiguousarray(qcodes), idx, dist
            self.uniq, self.L = uniq, int(L)
            self._sorted = self._pos = None
            return
        if rows is None:
            rows = np.arange(len(qcodes), dtype=np.int64)
        n = len(rows)
        all_rows = n == len(qcodes)
        if group is not None and n:
            # first kept row of every distinct guide, linear: rows are written in DESCENDING order, so the last
            # assignment to a slot -- the smallest row -- wins; only slots that are read back are ever written
            group = np.asarray(group)
            it = np.int32 if n < (1 << 31) else np.int64
            rev = np.arange(n - 1, -1, -1, dtype=it)
            slot = np.empty(int(group.max()) + 1 if len(uniq) == 0 else len(uniq), dtype=it)
            slot[group[::-1]] = rev
            kept = np.flatnonzero(slot[group] == np.arange(n, dtype=it))
        else:
            codes = qcodes if all_rows else qcodes[rows]
            order = np.argsort(codes, kind="stable")
            srt = codes[order]
            head = np.ones(n, dtype=bool)
            head[1:] = srt[1:] != srt[:-1]
            keep_mask = np.zeros(n, dtype=bool)
            keep_mask[order[head]] = True                    # stable sort: smallest row of each group
            kept = np.flatnonzero(keep_mask)
        final = kept if all_rows else rows[kept]
        whole = len(final) == len(qcodes)                    # nothing dropped: no copy at all
        self.codes = np.ascontiguousarray(qcodes if whole else qcodes[final])
        self.idx = idx if whole else idx[final]
        self.dist = dist if whole else dist[final]
        self.uniq, self.L = uniq, int(L)
        self._sorted = self._pos = None                      # lookup index, built on first use

    def _lookup_index(self):
        if self._sorted is None:
            self._pos = np.argsort(self.codes, kind="stable")
            self._sorted = self.codes[self._pos]
        return self._sorted, self._pos

    def _row(self, seq) -> int:
        if not isinstance(seq, str) or len(seq) != self.L:
            return -1
        try:
            code = encode_guides([seq], self.L)[0]
        except ValueError:
            return -1
        srt, pos = self._lookup_index()
        j = int(np.searchsorted(srt, code))
        if j < len(srt) and srt[j] == code:
            return int(pos[j])
        return -1

    def __getitem__(self, seq):
        r = self._row(seq)
        if r < 0:
            raise KeyError(seq)
        m = self.idx[r] >= 0
        seqs = [s.decode() for s in decode_guides(self.uniq[self.idx[r][m]], self.L)]
        return {"target": seq, "neighbors": {"seqs": seqs, "dist": [int(x) for x in self.dist[r][m]]}}

    def __contains__(self, seq):
        return self._row(seq) >= 0

    def __iter__(self):
        return (s.decode() for s in decode_guides(self.codes, self.L))

    def __len__(self):
        return len(self.codes)

    def __repr__(self):
        return "NeighborMap(%d guides, k=%d)" % (len(self), self.idx.shape[1] if self.idx.ndim == 2 else 0)

    # ---- vectorised accessors for table assembly (no per-row Python)
    def key_array(self) -> np.ndarray:
        """kept guide strings as a numpy S{L} array, in key order"""
        return decode_guides(self.codes, self.L)

    def distance_matrix(self) -> np.ndarray:
        return self.dist

    def index_matrix(self) -> np.ndarray:
        return self.idx
