"""CPU ORACLE for the GuideMaker off-target hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``guidemaker_b200`` never does; it fails loudly when its CUDA library is missing.

Two layers:

* ``c_*``  -- ctypes bindings to ``oracle/gm_oracle.c`` (plain-C restatement, used at sizes
  up to ~10^4 x 10^6 pairs and as the timed CPU baseline);
* ``py_*`` -- slow, literal Python restatements used only on tiny inputs to cross-check the C
  code: the PAM scan follows the reference line by line with ``regex.finditer(...,
  overlapped=True)`` (core.py:142-246), the duplicate flag uses ``pandas.Series.duplicated``
  (core.py:416), the distances are character loops.

Parity pin: ``tests/test_oracle_golden.py`` (reference known-answer vectors from
/root/reference/tests/test_core.py and fixtures made by running the reference's real core.py,
``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
BASES = "ACGT"

# IUPAC table of core.py:118-121 / :1103-1120
IUPAC = {
    "A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG",
    "Y": "CT", "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC",
}
_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn")


def build(force: bool = False) -> str:
    """Compile oracle/gm_oracle.c (gcc, see oracle/Makefile)."""
    so = os.path.join(_HERE, "libgm_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "gm_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = build()
        try:
            L = ctypes.CDLL(so)
        except OSError:  # e.g. built for another CPU -> rebuild here
            so = build(force=True)
            L = ctypes.CDLL(so)
        u8p, u16p, u32p, u64p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint16, ctypes.c_uint32, ctypes.c_uint64))
        i32p, i64p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)
        L.gmo_pam_scan.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, i64p, i64p]
        L.gmo_seed_dedup.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.gmo_first_occurrence.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.gmo_knn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.gmo_min_dist.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_int]
        L.gmo_restriction.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p]
        L.gmo_knn_hamming_fast.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.gmo_knn_hamming_fast.restype = ctypes.c_int
        L.gmo_has_avx512_popcnt.restype = ctypes.c_int
        for f in (L.gmo_pam_scan, L.gmo_seed_dedup, L.gmo_first_occurrence, L.gmo_knn, L.gmo_min_dist, L.gmo_num_threads,
                  L.gmo_restriction):
            f.restype = ctypes.c_int
        del u8p, u16p, u32p, u64p, i32p
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ packing helpers

def pack(seq: str) -> int:
    """guide string -> guide2bit (base i at bits 2i..2i+1)."""
    v = 0
    for i, ch in enumerate(seq):
        v |= CODE[ch] << (2 * i)
    return v


def unpack(v: int, L: int) -> str:
    return "".join(BASES[(int(v) >> (2 * i)) & 3] for i in range(L))


def pack_many(seqs) -> np.ndarray:
    return np.array([pack(s) for s in seqs], dtype=np.uint64)


def reverse_complement(s: str) -> str:
    """Bio.Seq.reverse_complement as used at core.py:95-106 (IUPAC aware, other bytes kept)."""
    return s.translate(_COMP)[::-1]


# ------------------------------------------------------------------ C oracle wrappers

def c_pam_scan(seq: bytes, pam: str, five_prime: bool, L: int):
    """-> (guides u64[n], start u32[n], pamcode u16[n], n_fwd, n_rev); forward rows first."""
    buf = np.frombuffer(seq, dtype=np.uint8)
    nf, nr = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib().gmo_pam_scan(_ptr(buf) if len(buf) else None, len(buf), pam.encode(), len(pam), int(five_prime), L,
                            None, None, None, 0, ctypes.byref(nf), ctypes.byref(nr))
    if rc != 0:
        raise ValueError(f"gmo_pam_scan rc={rc}")
    n = nf.value + nr.value
    g = np.zeros(n, np.uint64); s = np.zeros(n, np.uint32); p = np.zeros(n, np.uint16)
    rc = lib().gmo_pam_scan(_ptr(buf) if len(buf) else None, len(buf), pam.encode(), len(pam), int(five_prime), L,
                            _ptr(g), _ptr(s), _ptr(p), n, ctypes.byref(nf), ctypes.byref(nr))
    if rc != 0:
        raise ValueError(f"gmo_pam_scan rc={rc}")
    return g, s, p, nf.value, nr.value


def c_seed_dedup(guides: np.ndarray, L: int, lsr: int, five_prime: bool) -> np.ndarray:
    guides = np.ascontiguousarray(guides, np.uint64)
    out = np.zeros(len(guides), np.uint8)
    rc = lib().gmo_seed_dedup(_ptr(guides), len(guides), L, lsr, int(five_prime), _ptr(out))
    if rc != 0:
        raise RuntimeError(rc)
    return out.astype(bool)


def c_first_occurrence(guides: np.ndarray) -> np.ndarray:
    guides = np.ascontiguousarray(guides, np.uint64)
    out = np.zeros(len(guides), np.int64)
    rc = lib().gmo_first_occurrence(_ptr(guides), len(guides), _ptr(out))
    if rc != 0:
        raise RuntimeError(rc)
    return out


def unique_first_order(guides: np.ndarray):
    """distinct guides in first-occurrence order, and row -> unique index map."""
    fr = c_first_occurrence(guides)
    is_first = fr == np.arange(len(guides))
    uniq = np.ascontiguousarray(guides[is_first])
    rank = np.cumsum(is_first) - 1
    return uniq, rank[fr]


def c_knn(targets, queries, L: int, metric: int, k: int, threads: int = 0):
    targets = np.ascontiguousarray(targets, np.uint64); queries = np.ascontiguousarray(queries, np.uint64)
    idx = np.zeros((len(queries), k), np.int32); dist = np.zeros((len(queries), k), np.uint8)
    rc = lib().gmo_knn(_ptr(targets), len(targets), _ptr(queries), len(queries), L, metric, k, _ptr(idx), _ptr(dist), threads)
    if rc != 0:
        raise RuntimeError(rc)
    return idx, dist


def c_knn_hamming_fast(targets, queries, L: int, k: int, threads: int = 0):
    """gmo_knn(metric 0) with a tuned schedule (AVX-512 VPOPCNTD when available, query blocking): the CPU arm of bench.py.
    Identical output; gmo_knn remains the checker."""
    t = np.ascontiguousarray(targets, np.uint64)
    q = np.ascontiguousarray(queries, np.uint64)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.uint8)
    rc = lib().gmo_knn_hamming_fast(_ptr(t), len(t), _ptr(q), len(q), L, k, _ptr(idx), _ptr(dist), threads)
    if rc:
        raise RuntimeError("gmo_knn_hamming_fast failed: %d" % rc)
    return idx, dist


def has_avx512_popcnt() -> bool:
    return bool(lib().gmo_has_avx512_popcnt())


def c_min_dist(targets, queries, L: int, metric: int, threads: int = 0):
    targets = np.ascontiguousarray(targets, np.uint64); queries = np.ascontiguousarray(queries, np.uint64)
    dist = np.zeros(len(queries), np.uint8)
    rc = lib().gmo_min_dist(_ptr(targets), len(targets), _ptr(queries), len(queries), L, metric, _ptr(dist), threads)
    if rc != 0:
        raise RuntimeError(rc)
    return dist


def np_gather_windows(seq, win_start, revcomp, width: int) -> np.ndarray:
    """target_seq30 windows (core.py:156,184,210-211,237) as plain numpy: slice, reverse-complement with Bio.Seq's
    IUPAC table where flagged, '?' for windows outside the buffer."""
    seq = np.ascontiguousarray(seq, np.uint8)
    win_start = np.asarray(win_start, np.int64)
    revcomp = np.asarray(revcomp).astype(bool)
    lut = np.arange(256, dtype=np.uint8)
    lut[np.frombuffer(b"ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", np.uint8)] = np.frombuffer(b"TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn", np.uint8)
    out = np.full((len(win_start), width), ord("?"), np.uint8)
    ok = (win_start >= 0) & (win_start + width <= len(seq))
    if ok.any():
        idx = win_start[ok][:, None] + np.arange(width)[None, :]
        rows = seq[idx]
        rc = revcomp[ok]
        rows[rc] = lut[rows[rc][:, ::-1]]
        out[ok] = rows
    return out


def c_restriction(guides, L: int, motifs) -> np.ndarray:
    """bool[n]: guide contains one of the IUPAC motifs (the caller passes sites AND reverse complements)."""
    guides = np.ascontiguousarray(guides, np.uint64)
    motifs = list(motifs)
    if any(len(m) > 32 for m in motifs):
        motifs = [m for m in motifs if len(m) <= 32]          # longer than any guide: can never occur
    buf = bytearray(32 * max(len(motifs), 1))
    lens = np.zeros(max(len(motifs), 1), np.int32)
    for t, m in enumerate(motifs):
        buf[32 * t: 32 * t + len(m)] = m.encode()
        lens[t] = len(m)
    out = np.zeros(len(guides), np.uint8)
    rc = lib().gmo_restriction(_ptr(guides), len(guides), L, bytes(buf), _ptr(lens), len(motifs), _ptr(out))
    if rc != 0:
        raise ValueError(f"gmo_restriction rc={rc}")
    return out.astype(bool)


def num_threads() -> int:
    return lib().gmo_num_threads()


# ------------------------------------------------------------------ literal Python restatements (tiny inputs)

def py_pam2re(pam: str) -> str:
    """core.py:108-122 (character classes written without the harmless literal '|')."""
    return "".join(IUPAC[b] if len(IUPAC[b]) == 1 else "[" + IUPAC[b] + "]" for b in pam)


def py_find_targets(seq: str, pam: str, five_prime: bool, L: int):
    """Rows of find_targets for one record, as tuples
    (target, exact_pam, start, stop, strand, pam_orientation, target_seq30) -- core.py:142-246."""
    import regex

    def ok(t):
        return len(t) == L and all(c in "ATCG" for c in t)

    rows = []
    fre, rre = py_pam2re(pam), py_pam2re(reverse_complement(pam))
    if five_prime:
        for m in regex.finditer(fre, seq, overlapped=True):                       # core.py:154-166
            t = seq[m.end(): m.end() + L]
            if ok(t):
                rows.append((t, m.group(0), m.end(), m.end() + L, True, True, seq[m.start() - 3: m.start() + 27]))
        for m in regex.finditer(rre, seq, overlapped=True):                       # core.py:207-220
            t = reverse_complement(seq[m.start() - L: m.start()])
            if ok(t):
                rows.append((t, reverse_complement(m.group(0)), m.start() - L, m.start(), False, True,
                             reverse_complement(seq[m.end() - 27: m.end() + 3])))
    else:
        for m in regex.finditer(fre, seq, overlapped=True):                       # core.py:182-193
            t = seq[m.start() - L: m.start()]
            if ok(t):
                rows.append((t, m.group(0), m.start() - L, m.start(), True, False, seq[m.end() - 27: m.end() + 3]))
        for m in regex.finditer(rre, seq, overlapped=True):                       # core.py:234-246
            t = reverse_complement(seq[m.end(): m.end() + L])
            if ok(t):
                rows.append((t, reverse_complement(m.group(0)), m.end(), m.end() + L, False, False,
                             reverse_complement(seq[m.start() - 3: m.start() + 27])))
    return rows


def py_restriction(targets, enzymes) -> np.ndarray:
    """check_restriction_enzymes, line by line (core.py:365-377): expand every site and its reverse complement
    with itertools.product over the IUPAC table, join with '|', regex-search every target."""
    import re
    from itertools import product
    element_to_exclude = []
    for record in set(enzymes):
        for rec in (record.upper(), reverse_complement(record.upper())):
            element_to_exclude.append(["".join(i) for i in product(*[IUPAC[j] for j in rec])])
    element_to_exclude = sum(element_to_exclude, [])
    if not element_to_exclude:
        return np.zeros(len(targets), bool)
    pat = re.compile("|".join(element_to_exclude))
    return np.array([pat.search(t) is not None for t in targets], dtype=bool)


def py_seed(t: str, lsr: int, five_prime: bool) -> str:
    """core.py:402-412."""
    if lsr == 0:
        return t
    return t[0:lsr] if five_prime else t[len(t) - lsr:]


def py_seed_dedup(targets, lsr: int, five_prime: bool) -> np.ndarray:
    """core.py:415-416."""
    import pandas as pd
    return pd.Series([py_seed(t, lsr, five_prime) for t in targets]).duplicated().to_numpy()


def py_hamming(a: str, b: str) -> int:
    return sum(x != y for x, y in zip(a, b))


def py_leven(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j - 1] + (ca != cb), prev[j] + 1, cur[j - 1] + 1))
        prev = cur
    return prev[-1]


def py_knn(targets, queries, metric: int, k: int):
    f = py_hamming if metric == 0 else py_leven
    idx, dist = [], []
    for q in queries:
        d = sorted((f(q, t), i) for i, t in enumerate(targets))[:k]
        idx.append([i for _, i in d]); dist.append([x for x, _ in d])
    return idx, dist
