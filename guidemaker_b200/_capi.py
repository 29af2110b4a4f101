"""ctypes binding of libgm_b200.so (include/gm_b200.h).  Thin: numpy arrays in, numpy arrays out.

There is deliberately no CPU fallback here: if the shared library is missing or no sm_100 device
is present every call raises (``EngineUnavailable`` / ``RuntimeError``).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._build import LIBPATH

METRIC_HAMMING, METRIC_LEVEN = 0, 1
MAX_L, MAX_K, MAX_PAM = 27, 32, 8
MAX_BASES = 0xFFFFFF00          # gm_scan_create: uint32 coordinates
MAX_INDEX = 1 << 27              # gm_index_create: 27-bit index in the (distance, index) key

_c_i64p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p


class EngineUnavailable(RuntimeError):
    """libgm_b200.so could not be loaded or no B200-class device is usable."""


_LIB = None

_SIGS = {
    "gm_init": [ctypes.c_int],
    "gm_version": [],
    "gm_trim": [],
    "gm_device_info": [ctypes.POINTER(ctypes.c_int)] * 3 + [_c_i64p],
    "gm_scan_create": [_vp, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                       ctypes.POINTER(_vp), _c_i64p, _c_i64p],
    "gm_scan_fetch": [_vp, _vp, _vp, _vp],
    "gm_scan_free": [_vp],
    "gm_gather_windows": [_vp, ctypes.c_int64, _vp, _vp, ctypes.c_int64, ctypes.c_int, _vp],
    "gm_seed_dedup": [_vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp],
    "gm_first_occurrence": [_vp, ctypes.c_int64, _vp],
    "gm_seed_dedup_dev": [_vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp],
    "gm_first_occurrence_dev": [_vp, ctypes.c_int64, _vp, _vp],
    "gm_scan_device_ptrs": [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), _c_i64p],
    "gm_session_create": [_vp, ctypes.c_int64, _vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                          ctypes.POINTER(_vp), _c_i64p],
    "gm_session_info": [_vp, _c_i64p] + [ctypes.POINTER(ctypes.c_int)] * 4,
    "gm_session_fetch_rows": [_vp, _vp, _vp, _vp, _vp, _vp],
    "gm_session_fetch_text": [_vp, _vp, _vp, ctypes.c_int, _vp],
    "gm_session_pam_histogram": [_vp, _vp],
    "gm_session_pam_categories": [_vp, _vp, _vp],
    "gm_session_seed_dedup": [_vp, ctypes.c_int, _vp],
    "gm_session_restriction": [_vp, _vp, _vp, ctypes.c_int, _vp],
    "gm_session_index": [_vp, ctypes.c_int, ctypes.POINTER(_vp), _vp, _vp, _c_i64p],
    "gm_session_knn": [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp],
    "gm_session_knn_dev": [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp],
    "gm_session_neighbors": [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _c_i64p, _c_i64p],
    "gm_session_filter_dev": [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _c_i64p, _c_i64p],
    "gm_session_fetch_neighbors": [_vp, _vp, _vp, _vp],
    "gm_session_free": [_vp],
    "gm_cfd_scores": [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp],
    "gm_restriction_scan": [_vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp],
    "gm_restriction_scan_dev": [_vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp],
    "gm_index_create": [_vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)],
    "gm_index_create_dev": [_vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp), _vp],
    "gm_index_info": [_vp, _c_i64p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)],
    "gm_index_free": [_vp],
    "gm_knn": [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp],
    "gm_knn_dev": [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp],
    "gm_min_dist": [_vp, _vp, ctypes.c_int64, _vp],
    "gm_min_dist_dev": [_vp, _vp, ctypes.c_int64, _vp, _vp],
    "gm_comm_unique_id": [_vp],
    "gm_comm_create": [_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)],
    "gm_comm_free": [_vp],
    "gm_knn_sharded": [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp],
    "gm_prof_enable": [ctypes.c_int],
    "gm_prof_reset": [],
    "gm_prof_read": [ctypes.POINTER(ctypes.c_double), _c_i64p, ctypes.POINTER(ctypes.c_double), _c_i64p],
    "gm_knn_tune": [ctypes.c_int, ctypes.c_int, ctypes.c_int],
    "gm_knn_engine": [ctypes.c_int],
    "gm_index_tune": [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int],
    "gm_microbench": [ctypes.c_int, ctypes.POINTER(ctypes.c_double)],
}
EXPORTS = tuple(_SIGS) + ("gm_last_error",)


def load_library(path: str = LIBPATH) -> ctypes.CDLL:
    """dlopen the engine and declare every prototype of include/gm_b200.h (no CUDA call is made)."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("GM_B200_LIB", path)        # kernel experiments: an alternative build of the same library
        if not os.path.exists(path):
            raise EngineUnavailable(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). guidemaker_b200 has no CPU fallback.")
        try:
            lib = ctypes.CDLL(path)
        except OSError as e:  # e.g. libcudart not found
            raise EngineUnavailable(f"cannot load {path}: {e}") from e
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        lib.gm_last_error.argtypes = []
        lib.gm_last_error.restype = ctypes.c_char_p
        _LIB = lib
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().gm_last_error().decode(errors="replace")
        if rc == -4:
            raise EngineUnavailable(f"{what}: {msg}")
        if rc == -2:
            raise ValueError(f"{what}: {msg}")
        if rc == -3:
            raise MemoryError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


_initialised = False


def init(device: int | None = None) -> None:
    """Bind the CUDA device (default: LOCAL_RANK or 0).  Raises EngineUnavailable without a B200."""
    global _initialised
    if _initialised and device is None:
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    _check(load_library().gm_init(device), "gm_init")
    _initialised = True


def device_info() -> dict:
    init()
    sm, ma, mi, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    _check(load_library().gm_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(mem)), "gm_device_info")
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "mem_bytes": mem.value}


# ---- K1 ---------------------------------------------------------------------------------------------
def pam_scan(seq: bytes | np.ndarray, pam: str, five_prime: bool, L: int):
    """-> (guide2bit u64[n], start u32[n], pamcode u16[n], n_fwd, n_rev); forward rows first."""
    init()
    lib = load_library()
    buf = np.frombuffer(seq, dtype=np.uint8) if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, np.uint8)
    h, nf, nr = _vp(), ctypes.c_int64(), ctypes.c_int64()
    _check(lib.gm_scan_create(_p(buf) if len(buf) else None, len(buf), pam.encode("ascii", "replace"), len(pam), int(bool(five_prime)),
                              int(L), ctypes.byref(h), ctypes.byref(nf), ctypes.byref(nr)), "gm_scan_create")
    try:
        n = nf.value + nr.value
        g = np.empty(n, np.uint64); s = np.empty(n, np.uint32); p = np.empty(n, np.uint16)
        if n:
            _check(lib.gm_scan_fetch(h, _p(g), _p(s), _p(p)), "gm_scan_fetch")
    finally:
        lib.gm_scan_free(h)
    return g, s, p, nf.value, nr.value


def gather_windows(seq: np.ndarray, win_start: np.ndarray, revcomp: np.ndarray, width: int) -> np.ndarray:
    """(n_rows, width) uint8: window i of the ASCII buffer, reverse-complemented where flagged; windows outside the
    buffer come back as '?' (the caller patches the few rows near record ends)."""
    init()
    seq = np.ascontiguousarray(seq, np.uint8)
    win_start = np.ascontiguousarray(win_start, np.int64)
    revcomp = np.ascontiguousarray(revcomp, np.uint8)
    out = np.empty((len(win_start), int(width)), np.uint8)
    _check(load_library().gm_gather_windows(_p(seq) if len(seq) else None, len(seq), _p(win_start), _p(revcomp), len(win_start),
                                            int(width), _p(out)), "gm_gather_windows")
    return out


def _motif_tables(motifs):
    motifs = list(motifs)
    sets = np.zeros((max(len(motifs), 1), MAX_MOTIF), np.uint8)
    lens = np.zeros(max(len(motifs), 1), np.int32)
    for t, m in enumerate(motifs):
        lens[t] = len(m)
        for j, ch in enumerate(m[:MAX_MOTIF]):
            sets[t, j] = IUPAC_SETS[ch]
    return sets, lens, len(motifs)


class Session:
    """One genome scan whose rows (and the genome) stay in HBM; the later stages of the hot path run off this handle
    (include/gm_b200.h, "session").  Records are joined by one invalid byte; ``rec_start[r]`` is the offset of record r
    in ``buf`` and ``rec_start[-1] == len(buf) + 1``."""

    def __init__(self, buf: np.ndarray, rec_start: np.ndarray, pam: str, five_prime: bool, L: int):
        init()
        self._h = _vp()
        buf = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, np.uint8)
        rec_start = np.ascontiguousarray(rec_start, np.int64)
        n = ctypes.c_int64()
        _check(load_library().gm_session_create(_p(buf) if len(buf) else None, len(buf), _p(rec_start), len(rec_start) - 1,
                                                pam.encode("ascii", "replace"), len(pam), int(bool(five_prime)), int(L),
                                                ctypes.byref(self._h), ctypes.byref(n)), "gm_session_create")
        self.n_rows, self.L, self.P, self.five_prime = n.value, int(L), len(pam), bool(five_prime)

    def fetch_rows(self, want_pamcode: bool = True):
        """-> guide2bit u64[n], start u32[n] (record-relative), pamcode u16[n] (None if not wanted), rec i32[n], strand bool[n]"""
        n = self.n_rows
        g = np.empty(n, np.uint64); s = np.empty(n, np.uint32); r = np.empty(n, np.int32); f = np.empty(n, np.uint8)
        p = np.empty(n, np.uint16) if want_pamcode else None
        if n:
            _check(load_library().gm_session_fetch_rows(self._h, _p(g), _p(s), _p(p) if want_pamcode else None, _p(r), _p(f)),
                   "gm_session_fetch_rows")
        return g, s, p, r, f.view(np.bool_)

    def pam_histogram(self) -> np.ndarray:
        """how often each of the 65536 possible packed PAM codes occurs among the rows"""
        h = np.zeros(1 << 16, np.uint32)
        _check(load_library().gm_session_pam_histogram(self._h, _p(h)), "gm_session_pam_histogram")
        return h

    def pam_categories(self, lut: np.ndarray) -> np.ndarray:
        """int8 category code of every row: lut[pamcode[row]] evaluated on the device"""
        lut = np.ascontiguousarray(lut, np.int8)
        assert lut.shape == (1 << 16,)
        out = np.empty(self.n_rows, np.int8)
        _check(load_library().gm_session_pam_categories(self._h, _p(lut), _p(out) if self.n_rows else None), "gm_session_pam_categories")
        return out

    def fetch_text(self, width: int = 30):
        """-> target ASCII (n, L), context ASCII (n, width), edge bool[n] (context window leaves its record: row is '?')"""
        n = self.n_rows
        t = np.empty((n, self.L), np.uint8); c = np.empty((n, int(width)), np.uint8); e = np.empty(n, np.uint8)
        if n:
            _check(load_library().gm_session_fetch_text(self._h, _p(t), _p(c), int(width), _p(e)), "gm_session_fetch_text")
        return t, c, e.view(np.bool_)

    def seed_dedup(self, lsr: int) -> np.ndarray:
        out = np.zeros(self.n_rows, np.uint8)
        _check(load_library().gm_session_seed_dedup(self._h, int(lsr), _p(out)), "gm_session_seed_dedup")
        return out.view(np.bool_)

    def restriction(self, motifs) -> np.ndarray:
        sets, lens, nm = _motif_tables(motifs)
        out = np.zeros(self.n_rows, np.uint8)
        _check(load_library().gm_session_restriction(self._h, _p(sets), _p(lens), nm, _p(out)), "gm_session_restriction")
        return out.view(np.bool_)

    def build_index(self, metric: int):
        """-> (Index over the distinct guides, uniq u64[n_u] in first-occurrence order, row2uniq i32[n])"""
        h, nu = _vp(), ctypes.c_int64()
        uniq = np.empty(self.n_rows, np.uint64); r2u = np.empty(self.n_rows, np.int32)
        _check(load_library().gm_session_index(self._h, int(metric), ctypes.byref(h), _p(uniq), _p(r2u), ctypes.byref(nu)), "gm_session_index")
        return Index.from_handle(h, nu.value, self.L, metric), uniq[: nu.value], r2u

    def knn(self, index: "Index", qmask: np.ndarray, k: int):
        qmask = np.ascontiguousarray(qmask, np.uint8 if qmask.dtype != np.bool_ else np.bool_).view(np.uint8)
        nq = int(np.count_nonzero(qmask))
        idx = np.empty((nq, k), np.int32); dist = np.empty((nq, k), np.uint8)
        _check(load_library().gm_session_knn(self._h, index._h, _p(qmask), nq, int(k), _p(idx), _p(dist)), "gm_session_knn")
        return idx, dist

    def neighbors(self, index: "Index", qmask: np.ndarray, k: int, editdist: int):
        """get_neighbors on the device: -> (codes u64[m], idx i32[m,k], dist u8[m,k], n_short) for the kept query rows
        (nearest other guide >= editdist away, first query row of its guide), in row order"""
        qmask = np.ascontiguousarray(qmask, np.uint8 if qmask.dtype != np.bool_ else np.bool_).view(np.uint8)
        nq = int(np.count_nonzero(qmask))
        kept, short = ctypes.c_int64(), ctypes.c_int64()
        _check(load_library().gm_session_neighbors(self._h, index._h, _p(qmask), nq, int(k), int(editdist), ctypes.byref(kept), ctypes.byref(short)),
               "gm_session_neighbors")
        m = kept.value
        codes = np.empty(m, np.uint64); idx = np.empty((m, k), np.int32); dist = np.empty((m, k), np.uint8)
        if m:
            _check(load_library().gm_session_fetch_neighbors(self._h, _p(codes), _p(idx), _p(dist)), "gm_session_fetch_neighbors")
        return codes, idx, dist, short.value

    def filter_dev(self, qmask: np.ndarray, k: int, editdist: int, d_idx: int, d_dist: int, stream: int = 0):
        """the selection of `neighbors` for (idx, dist) rows already on the device (all masked rows, row order)"""
        qmask = np.ascontiguousarray(qmask, np.uint8 if qmask.dtype != np.bool_ else np.bool_).view(np.uint8)
        nq = int(np.count_nonzero(qmask))
        kept, short = ctypes.c_int64(), ctypes.c_int64()
        _check(load_library().gm_session_filter_dev(self._h, _p(qmask), nq, int(k), int(editdist), _vp(d_idx), _vp(d_dist), _vp(stream),
                                                    ctypes.byref(kept), ctypes.byref(short)), "gm_session_filter_dev")
        m = kept.value
        codes = np.empty(m, np.uint64); idx = np.empty((m, k), np.int32); dist = np.empty((m, k), np.uint8)
        if m:
            _check(load_library().gm_session_fetch_neighbors(self._h, _p(codes), _p(idx), _p(dist)), "gm_session_fetch_neighbors")
        return codes, idx, dist, short.value

    def knn_dev(self, index: "Index", qmask: np.ndarray, k: int, d_idx: int, d_dist: int, stream: int = 0) -> int:
        """kNN of the masked rows into DEVICE buffers (no synchronisation); returns the number of query rows"""
        qmask = np.ascontiguousarray(qmask, np.uint8 if qmask.dtype != np.bool_ else np.bool_).view(np.uint8)
        nq = int(np.count_nonzero(qmask))
        _check(load_library().gm_session_knn_dev(self._h, index._h, _p(qmask), nq, int(k), _vp(d_idx), _vp(d_dist), _vp(stream)),
               "gm_session_knn_dev")
        return nq

    def close(self):
        if self._h:
            load_library().gm_session_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- K2 ---------------------------------------------------------------------------------------------
def seed_dedup(guides: np.ndarray, L: int, lsr: int, five_prime: bool) -> np.ndarray:
    init()
    guides = np.ascontiguousarray(guides, np.uint64)
    out = np.zeros(len(guides), np.uint8)
    _check(load_library().gm_seed_dedup(_p(guides), len(guides), int(L), int(lsr), int(bool(five_prime)), _p(out)), "gm_seed_dedup")
    return out.view(np.bool_)


def first_occurrence(keys: np.ndarray) -> np.ndarray:
    init()
    keys = np.ascontiguousarray(keys, np.uint64)
    out = np.zeros(len(keys), np.int64)
    _check(load_library().gm_first_occurrence(_p(keys), len(keys), _p(out)), "gm_first_occurrence")
    return out


# IUPAC letter -> accepted bases (bit 0 = A, 1 = C, 2 = G, 3 = T); X and N accept everything (core.py:1107-1110)
IUPAC_SETS = {"A": 1, "C": 2, "G": 4, "T": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12, "V": 7, "H": 11, "D": 13,
              "B": 14, "X": 15, "N": 15}
MAX_MOTIF = 32


def restriction_scan(guides: np.ndarray, L: int, motifs) -> np.ndarray:
    """bool[n]: guide contains one of the IUPAC motifs (upper-case strings) at some offset (K6)."""
    init()
    guides = np.ascontiguousarray(guides, np.uint64)
    sets, lens, nm = _motif_tables(motifs)
    out = np.zeros(len(guides), np.uint8)
    _check(load_library().gm_restriction_scan(_p(guides), len(guides), int(L), _p(sets), _p(lens), nm, _p(out)),
           "gm_restriction_scan")
    return out.view(np.bool_)


# ---- K3/K4/K5 ---------------------------------------------------------------------------------------
class Index:
    """Owns one gm index handle (the distinct-guide table resident in HBM)."""

    def __init__(self, uniq2bit, L: int, metric: int, device_ptr: int | None = None, n: int | None = None, stream: int = 0):
        init()
        self._h = _vp()
        self.L, self.metric = int(L), int(metric)
        lib = load_library()
        if device_ptr is None:
            uniq2bit = np.ascontiguousarray(uniq2bit, np.uint64)
            self.n = len(uniq2bit)
            _check(lib.gm_index_create(_p(uniq2bit) if self.n else None, self.n, self.L, self.metric, ctypes.byref(self._h)), "gm_index_create")
        else:
            self.n = int(n)
            _check(lib.gm_index_create_dev(_vp(device_ptr), self.n, self.L, self.metric, ctypes.byref(self._h), _vp(stream)), "gm_index_create_dev")

    @classmethod
    def from_handle(cls, handle, n: int, L: int, metric: int) -> "Index":
        """adopt an index handle created inside the library (gm_session_index)"""
        self = cls.__new__(cls)
        self._h, self.n, self.L, self.metric = handle, int(n), int(L), int(metric)
        return self

    def knn(self, q2bit: np.ndarray, k: int, out_idx: np.ndarray | None = None, out_dist: np.ndarray | None = None):
        """host buffers in, host buffers out (synchronous); out_* may be caller-owned (e.g. pinned) arrays"""
        q2bit = np.ascontiguousarray(q2bit, np.uint64)
        q = len(q2bit)
        idx = np.empty((q, k), np.int32) if out_idx is None else out_idx
        dist = np.empty((q, k), np.uint8) if out_dist is None else out_dist
        assert idx.shape == (q, k) and idx.dtype == np.int32 and idx.flags.c_contiguous
        assert dist.shape == (q, k) and dist.dtype == np.uint8 and dist.flags.c_contiguous
        _check(load_library().gm_knn(self._h, _p(q2bit) if q else None, q, int(k), _p(idx), _p(dist)), "gm_knn")
        return idx, dist

    def min_dist(self, q2bit: np.ndarray) -> np.ndarray:
        q2bit = np.ascontiguousarray(q2bit, np.uint64)
        dist = np.empty(len(q2bit), np.uint8)
        _check(load_library().gm_min_dist(self._h, _p(q2bit) if len(q2bit) else None, len(q2bit), _p(dist)), "gm_min_dist")
        return dist

    def knn_dev(self, d_q: int, q: int, k: int, d_idx: int, d_dist: int, stream: int = 0) -> None:
        _check(load_library().gm_knn_dev(self._h, _vp(d_q), int(q), int(k), _vp(d_idx), _vp(d_dist), _vp(stream)), "gm_knn_dev")

    def min_dist_dev(self, d_q: int, q: int, d_dist: int, stream: int = 0) -> None:
        _check(load_library().gm_min_dist_dev(self._h, _vp(d_q), int(q), _vp(d_dist), _vp(stream)), "gm_min_dist_dev")

    def tune(self, engine: int = -1, queries_per_thread: int = -1, splits: int = -1, warm_sample: int = -2) -> None:
        """per-handle engine / tuning (defaults follow knn_engine / knn_tune)"""
        _check(load_library().gm_index_tune(self._h, int(engine), int(queries_per_thread), int(splits), int(warm_sample)), "gm_index_tune")

    def close(self):
        if self._h:
            load_library().gm_index_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- multi-GPU through the C ABI (NCCL inside the library) ---------------------------------------------
class Comm:
    """NCCL communicator owned by the library: rank 0 calls ``Comm.unique_id()``, the host carries the 128 bytes to the
    other ranks, every rank constructs ``Comm(id, rank, world)``."""

    @staticmethod
    def unique_id() -> bytes:
        buf = np.zeros(128, np.uint8)
        _check(load_library().gm_comm_unique_id(_p(buf)), "gm_comm_unique_id")
        return buf.tobytes()

    def __init__(self, unique_id: bytes, rank: int, world: int):
        init()
        self._h = _vp()
        self.rank, self.world = int(rank), int(world)
        buf = np.frombuffer(unique_id, np.uint8).copy()
        _check(load_library().gm_comm_create(_p(buf), self.rank, self.world, ctypes.byref(self._h)), "gm_comm_create")

    def knn(self, index: "Index", q2bit: np.ndarray, k: int):
        """gm_knn_sharded: every rank passes the same rows; every rank gets all rows back"""
        q2bit = np.ascontiguousarray(q2bit, np.uint64)
        idx = np.empty((len(q2bit), k), np.int32); dist = np.empty((len(q2bit), k), np.uint8)
        _check(load_library().gm_knn_sharded(index._h, self._h, _p(q2bit) if len(q2bit) else None, len(q2bit), int(k), _p(idx), _p(dist)),
               "gm_knn_sharded")
        return idx, dist

    def close(self):
        if self._h:
            load_library().gm_comm_free(self._h)
            self._h = _vp()


# ---- measurement hooks ------------------------------------------------------------------------------
def prof_enable(on: bool = True):
    _check(load_library().gm_prof_enable(int(on)), "gm_prof_enable")


def prof_reset():
    _check(load_library().gm_prof_reset(), "gm_prof_reset")


def prof_read() -> dict:
    ms, n, pairs, alln = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_int64()
    _check(load_library().gm_prof_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(pairs), ctypes.byref(alln)), "gm_prof_read")
    return {"scan_kernel_ms": ms.value, "scan_kernel_launches": n.value, "pairs": pairs.value, "all_kernel_launches": alln.value}


def knn_tune(queries_per_thread: int = 0, splits: int = 0, warm_sample: int = -1):
    _check(load_library().gm_knn_tune(queries_per_thread, splits, warm_sample), "gm_knn_tune")


def knn_engine(engine: int):
    """Hamming: 0 = K3a XOR/POPC (INT pipes), 1 = K3b tcgen05 int8 GEMM (tensor pipe); Levenshtein: 0 = plain scan,
    1 = prefix-sorted table with shared DP states (K4p)"""
    _check(load_library().gm_knn_engine(int(engine)), "gm_knn_engine")


def microbench(what: int) -> float:
    init()
    v = ctypes.c_double()
    _check(load_library().gm_microbench(int(what), ctypes.byref(v)), "gm_microbench")
    return v.value
