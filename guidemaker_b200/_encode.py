"""Vectorised host-side conversion between guide strings and the packed ``guide2bit`` format
(base i in bits [2i,2i+1], A=0 C=1 G=2 T=3).  No per-row Python."""
from __future__ import annotations

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_LUT = np.full(256, 255, dtype=np.uint8)
_LUT[_ASCII] = np.arange(4, dtype=np.uint8)


def as_byte_matrix(seqs, L: int | None = None) -> np.ndarray:
    """list/array/Series of equal-length ASCII strings -> (N, L) uint8 matrix."""
    if isinstance(seqs, np.ndarray) and seqs.dtype == np.uint8 and seqs.ndim == 2:
        return seqs
    try:
        import pandas as pd
        if isinstance(seqs, (pd.Series, pd.Index)):
            seqs = seqs.to_numpy(dtype=object, na_value="")
    except ImportError:  # pragma: no cover
        pass
    arr = np.asarray(seqs)
    if arr.dtype.kind != "S":
        arr = np.asarray(arr, dtype=object)
        lens = np.fromiter((len(s) for s in arr), dtype=np.int64, count=len(arr))
        width = int(L if L is not None else (lens.max() if len(lens) else 0))
        if len(arr) and (lens != width).any():
            raise ValueError("all guide sequences must have the same length (%d)" % width)
        arr = arr.astype("S%d" % max(width, 1)) if len(arr) else np.empty(0, dtype="S%d" % max(width, 1))
    width = arr.dtype.itemsize
    if L is not None and len(arr) and width != L:
        raise ValueError("all guide sequences must have the same length (%d)" % L)
    return np.ascontiguousarray(arr).view(np.uint8).reshape(len(arr), width)


def _spread32(v: np.ndarray) -> np.ndarray:
    """bit i of a uint32 -> bit 2i of a uint64 (vectorised)"""
    x = v.astype(np.uint64)
    x = (x | (x << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
    x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    x = (x | (x << np.uint64(2))) & np.uint64(0x3333333333333333)
    x = (x | (x << np.uint64(1))) & np.uint64(0x5555555555555555)
    return x


def encode_matrix(mat: np.ndarray) -> np.ndarray:
    """(N, L) uint8 ASCII -> uint64 guide2bit.  Raises on anything but upper-case A/C/G/T."""
    n, L = mat.shape
    if L > 27:
        raise ValueError("guides longer than 27 nt are not supported (got %d)" % L)
    if n == 0 or L == 0:
        return np.zeros(n, dtype=np.uint64)
    seen = np.flatnonzero(np.bincount(mat.reshape(-1), minlength=256))
    if (_LUT[seen] > 3).any():
        raise ValueError("guide sequences may contain only A, C, G, T")
    x = (mat >> 1) & 3                                    # A:0 C:1 G:3 T:2
    codes = x ^ (x >> 1)                                  # A:0 C:1 G:2 T:3
    planes = []
    for bit in (0, 1):                                   # bit planes via packbits, then interleave
        b = np.packbits((codes >> bit) & 1, axis=1, bitorder="little")
        w = np.zeros((n, 4), dtype=np.uint8)
        w[:, : b.shape[1]] = b
        planes.append(w.view("<u4").reshape(n))
    return _spread32(planes[0]) | (_spread32(planes[1]) << np.uint64(1))


def encode_guides(seqs, L: int | None = None) -> np.ndarray:
    return encode_matrix(as_byte_matrix(seqs, L))


# byte of a guide2bit word (4 bases) -> the 4 ASCII letters, little endian (base order = byte order)
_LUT4 = np.zeros(256, dtype="<u4")
for _b in range(256):
    _LUT4[_b] = sum(int(_ASCII[(_b >> (2 * _j)) & 3]) << (8 * _j) for _j in range(4))


def decode_matrix(g: np.ndarray, L: int) -> np.ndarray:
    """uint64 guide2bit -> (N, L) uint8 ASCII matrix (one table look-up per 4 bases)."""
    g = np.ascontiguousarray(g, dtype="<u8")
    nb = (L + 3) // 4
    quads = _LUT4[g.view(np.uint8).reshape(len(g), 8)[:, :nb]]          # (N, nb) uint32 = 4 letters each
    return np.ascontiguousarray(quads.view(np.uint8).reshape(len(g), 4 * nb)[:, :L])


def decode_guides(g: np.ndarray, L: int) -> np.ndarray:
    """uint64 guide2bit -> numpy array of dtype S{L}."""
    if L == 0:
        return np.zeros(len(g), dtype="S1")
    return decode_matrix(g, L).view("S%d" % L).reshape(len(g))


def matrix_to_strings(mat: np.ndarray):
    """(N, W) uint8 -> pandas-ready string array without creating Python objects (Arrow buffers)."""
    import pyarrow as pa
    n, w = mat.shape
    offsets = pa.py_buffer(np.arange(0, (n + 1) * w, w, dtype=np.int64) if w else np.zeros(n + 1, np.int64))
    data = pa.py_buffer(np.ascontiguousarray(mat))
    return pa.LargeStringArray.from_buffers(n, offsets, data)
