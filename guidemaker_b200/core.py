"""Hot-path surface of ``guidemaker.core`` on the B200 engine.

Same class names, constructor/method signatures, attribute names, DataFrame schema and error
behaviour as /root/reference/guidemaker/core.py for the off-target path:

    PamTarget.__init__ / find_targets                          core.py:49-69, :83-292
    TargetProcessor.__init__ / check_restriction_enzymes /
        find_unique_near_pam / create_index / get_neighbors /
        export_bed / get_control_seqs                          core.py:304-633
    extend_ambiguous_dna                                       core.py:1093-1124

All arithmetic (PAM scan, seed duplicate flags, distinct-guide table, exact kNN, control
min-distance) runs in ``libgm_b200.so``; this module only moves arrays in and out of pandas
without per-row Python.  There is no CPU fallback: without the CUDA engine every hot method raises.

Deliberate, documented differences from the reference:
  Q2/Q3  the index holds the distinct guides in FIRST-OCCURRENCE order (the reference uses the
         process-random order of ``list(set(...))``) and ``neighbors[seq]["neighbors"]["seqs"]``
         are the true neighbour sequences (the reference maps ids through the wrong table);
  Q14    the search is exact, so distances can only be <= the reference's approximate HNSW ones;
  ``self.neighbors`` is a read-only Mapping backed by arrays, not a dict of dicts;
  ``exact_pam`` / ``seqid`` are always categorical (the reference's concat silently degrades them to
         strings when per-record categories differ).
"""
from __future__ import annotations

import hashlib
import logging
import statistics
import zlib
from copy import deepcopy
from itertools import product
from typing import List

import numpy as np
import pandas as pd
import yaml

from . import _capi
from ._encode import as_byte_matrix, decode_matrix, encode_matrix, matrix_to_strings
from .neighbors import ExactIndex, NeighborMap
from .sharding import (broadcast_rank0, nccl_group, sharded_knn, sharded_min_dist, sharded_session_knn,
                       sharded_session_neighbors, world)

logger = logging.getLogger(__name__)

_IUPAC_LETTERS = ['A', 'C', 'G', 'T', 'M', 'R', 'W', 'S', 'Y', 'K', 'V', 'H', 'D', 'B', 'X', 'N']

# Bio.Seq.reverse_complement as used at core.py:95-106: IUPAC aware, other bytes unchanged
_COMP_FROM = b"ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn"
_COMP_TO = b"TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn"
_COMP_LUT = np.arange(256, dtype=np.uint8)
_COMP_LUT[np.frombuffer(_COMP_FROM, np.uint8)] = np.frombuffer(_COMP_TO, np.uint8)
_COMP_TABLE = bytes.maketrans(_COMP_FROM, _COMP_TO)


def _reverse_complement(s: str) -> str:
    return s.translate(str.maketrans(_COMP_FROM.decode(), _COMP_TO.decode()))[::-1]


def _str_series(mat: np.ndarray, index=None) -> pd.Series:
    """(N, W) ASCII matrix -> pandas str Series built from Arrow buffers (no Python objects)."""
    if mat.shape[1] == 0:
        return pd.Series([""] * len(mat), dtype="str", index=index)
    return pd.Series(pd.array(matrix_to_strings(mat), dtype="str"), index=index)


def _series_matrix(s: pd.Series) -> np.ndarray:
    """str Series of equal-length guides -> (N, L) uint8, zero-copy from Arrow when possible."""
    try:
        import pyarrow as pa
        arr = pa.array(s, type=pa.large_string()) if not hasattr(s.array, "_pa_array") else s.array._pa_array
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        if arr.null_count == 0 and len(arr):
            if pa.types.is_string(arr.type):
                arr = arr.cast(pa.large_string())
            bufs = arr.buffers()
            off = np.frombuffer(bufs[1], dtype=np.int64)[arr.offset: arr.offset + len(arr) + 1]
            width = int(off[1] - off[0])
            if width > 0 and np.array_equal(off, off[0] + np.arange(len(arr) + 1, dtype=np.int64) * width):
                data = np.frombuffer(bufs[2], dtype=np.uint8)[off[0]: off[-1]]
                return data.reshape(len(arr), width)
    except Exception:  # noqa: BLE001 -- any surprise falls back to the generic (slower) route
        pass
    return as_byte_matrix(s)


class _PackedStash:
    """Packed guides riding along in ``DataFrame.attrs``.  pandas deep-copies ``attrs`` into every Series/frame derived
    from the frame (``__finalize__``); the stash is immutable, so copies share it instead of duplicating 8 bytes per
    row on every column access."""
    __slots__ = ("crc", "shape", "guides", "session", "mat", "addr")

    def __init__(self, crc, shape, guides, session=None, mat=None):
        self.crc, self.shape, self.guides, self.session = crc, tuple(shape), guides, session
        self.guides.setflags(write=False)
        # The `target` column is an Arrow array built zero-copy over `mat`: Arrow buffers are immutable, so a column
        # whose data buffer still lives at this address IS this column -- an O(1) identity check; the CRC (lazy) is
        # only needed for frames whose column was rebuilt.
        self.mat = mat
        self.addr = None if mat is None else mat.ctypes.data
        if mat is not None:
            mat.setflags(write=False)

    def matches_column(self, col) -> bool:
        """True iff `col` (a pandas str Series) is backed by the very Arrow buffer this stash was made for"""
        try:
            arr = col.array._pa_array
            if arr.num_chunks != 1:
                return False
            c = arr.chunk(0)
            return (self.addr is not None and len(c) == self.shape[0] and c.offset == 0 and c.null_count == 0
                    and c.buffers()[2].address == self.addr and c.buffers()[2].size == self.shape[0] * self.shape[1])
        except Exception:  # noqa: BLE001
            return False

    def crc32(self):
        if self.crc is None:
            self.crc = zlib.crc32(self.mat)
        return self.crc

    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self

    def __eq__(self, other):                           # pandas compares attrs when combining objects
        return self is other

    def __hash__(self):
        return id(self)


class PamTarget:
    """A Protospacer Adjacent Motif (PAM) and its targets (core.py:39-292)."""

    def __init__(self, pam: str, pam_orientation: str, dtype: str) -> None:
        for letter in pam.upper():
            assert letter in _IUPAC_LETTERS
        assert pam_orientation in ["3prime", "5prime"]
        self.pam: str = pam.upper()
        self.pam_orientation: str = pam_orientation
        self.dtype: str = dtype

    def __str__(self) -> str:
        return "A PAM object: {self.pam}".format(self=self)

    def find_targets(self, seq_record_iter: object, target_len: int) -> pd.DataFrame:
        """All targets next to a PAM match on both strands (core.py:83-292).

        Records are duck-typed: ``.id`` and ``str(.seq)``.  Row order, coordinates and the 12-column
        schema are the reference's: per record all forward hits (ascending), then all reverse hits.
        """
        five = self.pam_orientation == "5prime"
        P, L = len(self.pam), int(target_len)
        ids, seqs, raw = [], [], []
        for record in seq_record_iter:
            ids.append(record.id)
            seq = record.seq
            if isinstance(getattr(seq, "_raw", None), bytes):      # fastaio records already hold ASCII bytes
                seqs.append(seq)                                   # (str() of it only if a row needs literal slicing)
                raw.append(seq._raw)
            else:
                s = str(seq)
                seqs.append(s)
                raw.append(s.encode("latin-1", "replace"))
        # one launch over the whole genome: records joined by an invalid base, which can neither
        # match a PAM position nor sit inside a target, so no hit straddles two records
        lens = np.array([len(b) for b in raw], dtype=np.int64)
        rec_start = np.zeros(len(raw) + 1, dtype=np.int64)
        if len(raw):
            rec_start[1:] = np.cumsum(lens + 1)
        buf = b"N".join(raw)
        if len(buf) == 0 and not raw:
            return pd.concat([])                     # the reference raises ValueError on no records
        if len(buf) > _capi.MAX_BASES:
            raise ValueError("genome of %d bases exceeds the engine's uint32 coordinate range (%d); scan it in parts"
                             % (len(buf), _capi.MAX_BASES))
        # The scan keeps genome and rows in HBM (a "session"): rows come back already in the reference's order (per
        # record forward hits, then reverse hits) with record-relative coordinates, and the two text columns are produced
        # on the device from the resident data.  The later stages (seed flags, distinct guides, index, kNN) run off
        # the same handle, which rides along in df.attrs.
        sess = _capi.Session(np.frombuffer(buf, np.uint8), rec_start, self.pam, five, L)
        n = sess.n_rows
        if n == 0:
            return pd.concat([])                     # zero hits -> ValueError (core.py:286-287)
        # exact_pam: the device counts the distinct packed PAM codes and, given the code -> category table, returns one int8
        # category per row; the uint16 code column only crosses PCIe when there are too many distinct PAMs for int8
        exact_pam = self._pam_categorical_dev(sess, P)
        guides, start, pamcode, rec, strand = sess.fetch_rows(want_pamcode=exact_pam is None)
        if exact_pam is None:
            exact_pam = self._pam_categorical(pamcode, P)
        target_mat, ctx, edge = sess.fetch_text(30)
        seq30 = self._target_seq30(ctx, edge, seqs, rec, start, strand, five, P, L)
        df = pd.DataFrame({
            "target": _str_series(target_mat),
            "exact_pam": exact_pam,
            "start": start,
            "stop": start + np.uint32(L),
            "strand": strand,
            "pam_orientation": np.full(n, five, dtype=bool),
            "target_seq30": seq30,
            "seqid": self._seqid_categorical(ids, rec),
        })
        # seedseq=NaN -> 'str', hasrestrictionsite=NaN, isseedduplicated=NaN -> bool True, dtype -> category
        # (core.py:288-291), built directly instead of through astype
        import pyarrow as pa
        df["seedseq"] = pd.array(pa.nulls(n, pa.large_string()), dtype="str")
        df["hasrestrictionsite"] = np.nan
        df["isseedduplicated"] = True
        df["dtype"] = pd.Categorical.from_codes(np.zeros(n, dtype=np.int8), categories=[self.dtype])
        # The packed guides and the session ride along so that TargetProcessor need not re-encode / re-upload n strings;
        # they are used only if the CRC of the `target` column's bytes still matches (any edit, filter or reorder of the
        # frame voids them).
        df.attrs["_gm_packed"] = _PackedStash(None, target_mat.shape, guides, sess, mat=target_mat)
        return df

    @staticmethod
    def _seqid_categorical(ids, rec: np.ndarray) -> pd.Categorical:
        """seqid as ``Series(strings).astype('category')`` would build it: categories in LEXICOGRAPHIC order, so that
        ``export_bed``'s sort by chrom orders contigs as the reference does (its per-record concat degrades the column to
        plain strings, which sort lexicographically) and not in FASTA record order."""
        if len(set(ids)) != len(ids):
            return pd.Categorical(np.asarray(ids, dtype=object)[rec])
        order = sorted(range(len(ids)), key=lambda i: ids[i])
        inv = np.empty(len(ids), dtype=np.int8 if len(ids) < 128 else np.int32)
        inv[order] = np.arange(len(ids))
        # the rows come grouped by record in ascending record order (the session's row order), so the code column is a
        # run-length expansion: record boundaries by binary search instead of a gather over all rows
        cuts = np.searchsorted(rec, np.arange(len(ids) + 1))
        runs = np.flatnonzero(cuts[1:] > cuts[:-1])   # records that have rows; each run must start and end with its record
        if cuts[0] == 0 and cuts[-1] == len(rec) and np.array_equal(rec[cuts[runs]], runs) and np.array_equal(rec[cuts[runs + 1] - 1], runs):
            codes = np.repeat(inv, np.diff(cuts))
        else:
            codes = inv[rec]                          # not grouped (a caller-made `rec`): plain gather
        return pd.Categorical.from_codes(codes, categories=pd.Index([ids[i] for i in order]), validate=False)

    @staticmethod
    def _pam_categories(present: np.ndarray, P: int):
        """(category strings in sorted order, rank of each present code in that order)"""
        cats = [b.decode() for b in decode_matrix(present.astype(np.uint64), P).view("S%d" % P).reshape(-1)] if P else [""] * len(present)
        order = np.argsort(np.array(cats, dtype=object), kind="stable")
        return [cats[i] for i in order], order

    @classmethod
    def _pam_categorical_dev(cls, sess, P: int):
        """exact_pam from the session: histogram of the packed codes and the per-row category on the device (None when
        the genome shows 128 or more distinct PAMs: the caller then builds the column from the code column)"""
        present = np.flatnonzero(sess.pam_histogram())
        if len(present) >= 128:
            return None
        cats, order = cls._pam_categories(present, P)
        lut = np.zeros(1 << 16, dtype=np.int8)
        lut[present[order]] = np.arange(len(present), dtype=np.int8)
        return pd.Categorical.from_codes(sess.pam_categories(lut), categories=pd.Index(cats, dtype="str"), validate=False)

    @staticmethod
    def _pam_categorical(pamcode: np.ndarray, P: int) -> pd.Categorical:
        """exact_pam as pd.Categorical(strings) would build it -- categories = the distinct PAM strings in sorted
        order -- from the packed codes by table look-up (no factorisation of n strings)."""
        present = np.flatnonzero(np.bincount(pamcode, minlength=1 << 16))
        cats = [b.decode() for b in decode_matrix(present.astype(np.uint64), P).view("S%d" % P).reshape(-1)] if P else [""] * len(present)
        order = np.argsort(np.array(cats, dtype=object), kind="stable")
        lut = np.zeros(1 << 16, dtype=np.int8 if len(present) < 128 else np.int32)
        lut[present[order]] = np.arange(len(present))
        return pd.Categorical.from_codes(lut[pamcode], categories=pd.Index([cats[i] for i in order], dtype="str"))

    @staticmethod
    def _target_seq30(ctx, edge, seqs, rec, start, strand, five, P, L) -> pd.Series:
        """The 30-nt context column (core.py:156,184,210-211,237): a Python slice of the record around the match,
        reverse-complemented for reverse hits, NOT validated.  `ctx` holds the device-gathered windows; rows flagged in
        `edge` (window leaves the record) get the reference's literal slice semantics -- the piece may be shorter than 30
        or empty."""
        n = len(start)
        rows = np.flatnonzero(edge)
        if len(rows) == 0:
            return _str_series(ctx)
        # The few exceptional rows are spliced in and the column is assembled as ONE Arrow string array with per-row
        # lengths (no detour through Python objects for the other rows).
        import pyarrow as pa
        width = np.full(n, 30, dtype=np.int64)
        flat = ctx.reshape(-1)
        parts, prev = [], 0
        for i in rows:
            st, fwd = int(start[i]), bool(strand[i])
            # match start/end on the forward text, from the target window
            ms = (st - P if fwd else st + L) if five else (st + L if fwd else st - P)
            a = ms - 3 if fwd == five else ms + P - 27       # 5p fwd / 3p rev slice [ms-3, ms+27); others [me-27, me+3)
            piece = str(seqs[rec[i]])[a: a + 30]
            piece = (piece if fwd else _reverse_complement(piece)).encode("latin-1", "replace")
            width[i] = len(piece)
            parts.append(flat[prev * 30: i * 30])     # the untouched rows before this one, as a view
            parts.append(np.frombuffer(piece, np.uint8))
            prev = i + 1
        parts.append(flat[prev * 30:])
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(width, out=offsets[1:])
        arr = pa.LargeStringArray.from_buffers(n, pa.py_buffer(offsets), pa.py_buffer(np.concatenate(parts)))
        return pd.Series(pd.array(arr, dtype="str"))


class TargetProcessor:
    """A set of guide RNA targets (core.py:295-633)."""

    def __init__(self, targets: pd.DataFrame, lsr: int, editdist: int = 2, knum: int = 2) -> None:
        self.targets = targets
        self.lsr: int = lsr
        self.editdist: int = editdist
        self.knum: int = knum
        self.nmslib_index: object = None
        self.neighbors: dict = {}
        self.closest_neighbor_df: pd.DataFrame = None
        self.ncontrolsearched: int = None
        self.gc_percent: float = None
        self.genomesize: float = None
        self.pam_orientation: bool = targets['pam_orientation'].iat[0]

    def __str__(self) -> None:
        info = "TargetList: contains a set of {} potential PAM targets".format(len(self.targets))
        return info

    def __len__(self) -> int:
        return len(self.targets)

    # ---- helpers -------------------------------------------------------------------------------------
    def _is_hamming(self) -> bool:
        return self.targets['dtype'].iat[0] == "hamming"       # anything else means Levenshtein (core.py:448,458)

    def _guide_matrix(self) -> np.ndarray:
        mat = _series_matrix(self.targets['target'])
        if mat.shape[1] > _capi.MAX_L:
            raise ValueError("guides longer than %d nt are not supported" % _capi.MAX_L)
        return mat

    def _packed(self):
        """packed guides of self.targets['target'], cached while that column object is unchanged"""
        col = self.targets['target']
        key = (id(self.targets), id(col.array), len(col))
        cache = getattr(self, "_packed_cache", None)
        if cache is None or cache[0] != key:
            stash = self.targets.attrs.get("_gm_packed") if isinstance(self.targets.attrs, dict) else None
            if isinstance(stash, _PackedStash) and stash.matches_column(col):
                mat, guides, sess = stash.mat, stash.guides, stash.session      # the very column find_targets built
            else:
                mat = self._guide_matrix()
                if (isinstance(stash, _PackedStash) and stash.mat is not None and tuple(stash.shape) == mat.shape
                        and stash.crc32() == zlib.crc32(np.ascontiguousarray(mat))):
                    guides, sess = stash.guides, stash.session                  # rebuilt column, same strings
                else:
                    guides, sess = encode_matrix(mat), None
            cache = (key, guides, mat.shape[1], col.array, sess, mat)           # keep the array alive: ids stay unique
            self._packed_cache = cache
        return cache[1], cache[2]

    def _matrix(self) -> np.ndarray:
        """(N, L) ASCII matrix of the `target` column (cached with the packed guides)"""
        self._packed()
        return self._packed_cache[5]

    def _session(self):
        """the device-resident scan these rows came from (``find_targets``), or None if the frame was edited since"""
        self._packed()
        return self._packed_cache[4]

    # ---- reference API -------------------------------------------------------------------------------
    def check_restriction_enzymes(self, restriction_enzyme_list: list = []) -> None:
        """Flag guides containing a restriction site or its reverse complement (core.py:354-377).

        The reference expands every IUPAC site into all concrete strings and regex-searches their alternation;
        the same predicate is evaluated on the packed guides by the K6 kernel (``gm_restriction_scan``)."""
        motifs = []
        for record in set(restriction_enzyme_list):
            for letter in record.upper():
                assert letter in _IUPAC_LETTERS
            motifs.append(record.upper())
            motifs.append(_reverse_complement(record.upper()))
        if len(motifs) > 0:
            guides, L = self._packed()
            sess = self._session()
            self.targets['hasrestrictionsite'] = sess.restriction(motifs) if sess is not None else _capi.restriction_scan(guides, L, motifs)
        else:
            self.targets['hasrestrictionsite'] = False

    def _one_hot_encode(self, seq_list: List[object]) -> List[str]:
        """nmslib bit_hamming text format (core.py:379-386); kept for callers of the facade."""
        charmap = {'A': '1 0 0 0', 'C': '0 1 0 0', 'G': '0 0 1 0', 'T': '0 0 0 1'}
        return [" ".join(charmap[letter] for letter in seq) for seq in seq_list]

    def find_unique_near_pam(self) -> None:
        """seedseq = PAM-proximal ``lsr`` nt; isseedduplicated = keep-first duplicate flag (core.py:388-416)."""
        guides, _ = self._packed()
        sess = self._session()
        mat = self._matrix()
        # The reference deep-copies the frame (core.py:414) so that the caller's frame keeps its columns; only two
        # columns are REPLACED below, so a shallow copy gives the same isolation without copying ~0.5 GB of strings.
        self.targets = self.targets.copy(deep=False)
        L = mat.shape[1]
        five = bool(self.pam_orientation)
        lsr = int(self.lsr)
        if lsr == 0 or (five and lsr >= L):
            seed, lsr_eff = mat, 0
        elif five:
            seed, lsr_eff = mat[:, 0:lsr], lsr
        else:
            cut = L - lsr                              # tseq[(len(tseq) - lsr):], Python slice semantics
            if cut < 0:
                cut = max(L + cut, 0)
            seed, lsr_eff = mat[:, cut:], L - cut
        self.targets['seedseq'] = _str_series(np.ascontiguousarray(seed), index=self.targets.index)
        lsr_key = lsr_eff if lsr_eff < L else 0
        self.targets['isseedduplicated'] = sess.seed_dedup(lsr_key) if sess is not None else _capi.seed_dedup(guides, L, lsr_key, five)
        col = self.targets['target']
        self._packed_cache = ((id(self.targets), id(col.array), len(col)), guides, L, col.array, sess, mat)

    def create_index(self, configpath: str, num_threads=2):
        """Upload the distinct guides to the GPU (replaces the HNSW build, core.py:418-467).

        ``NMSLIB.{M,efc,post}`` are read (a missing key raises as in the reference) and ignored: the
        search is exact.  ``num_threads`` is accepted and ignored."""
        with open(configpath) as cf:
            config = yaml.safe_load(cf)
        M, efC, post = config['NMSLIB']['M'], config['NMSLIB']['efc'], config['NMSLIB']['post']  # noqa: F841
        guides, L = self._packed()
        sess = self._session()
        metric = _capi.METRIC_HAMMING if self._is_hamming() else _capi.METRIC_LEVEN
        if sess is not None:                                   # distinct guides, index and row map built on the device
            engine, uniq, row2uniq = sess.build_index(metric)
            self.nmslib_index = ExactIndex(uniq, L, metric, engine=engine)
            self._index_session = sess
        else:
            first_row = _capi.first_occurrence(guides)
            is_first = first_row == np.arange(len(guides))
            uniq = np.ascontiguousarray(guides[is_first])
            if len(uniq) >= _capi.MAX_INDEX:
                raise ValueError("%d distinct guides exceed the engine's limit of 2^27 - 1 per index" % len(uniq))
            self.nmslib_index = ExactIndex(uniq, L, metric)
            row2uniq = (np.cumsum(is_first, dtype=np.int64) - 1)[first_row]
            self._index_session = None
        # row -> index of its guide in the distinct-guide table (valid while the packed cache is)
        self._row2uniq = (self._packed_cache[0], row2uniq)

    def get_neighbors(self, configpath, num_threads=2) -> None:
        """k nearest guides of every query row; keep a query iff its nearest OTHER guide is at least
        ``editdist`` away (core.py:471-523).  Writes ``self.neighbors``."""
        with open(configpath) as cf:
            config = yaml.safe_load(cf)
        ef = config['NMSLIB']['ef']  # noqa: F841  (HNSW efSearch; no-op for the exact search)
        t = self.targets
        qmask = ((t['isseedduplicated'] == False) | (t['hasrestrictionsite'] == False)).to_numpy(dtype=bool)  # noqa: E712
        guides, L = self._packed()
        sess = self._session()
        q = guides if qmask.all() else np.ascontiguousarray(guides[qmask])
        index = self.nmslib_index
        index.setQueryTimeParams({'efSearch': ef})
        if len(q) == 0:
            self.neighbors = NeighborMap(q, np.zeros((0, self.knum), np.int32), np.zeros((0, self.knum), np.uint8), index.uniq, L)
            return
        on_device = sess is not None and hasattr(index._engine, "_h")
        if on_device and hasattr(sess, "neighbors") and (world()[1] == 1 or nccl_group()):
            # search, distance filter and one-entry-per-guide rule all on the device; only kept rows come back (several
            # ranks: shards searched and all-gathered on the devices first, every rank filters the gathered table)
            if world()[1] == 1:
                codes, idx, dist, n_short = sess.neighbors(index._engine, qmask, int(self.knum), int(self.editdist))
            else:
                codes, idx, dist, n_short = sharded_session_neighbors(sess, index._engine, qmask, int(self.knum), int(self.editdist))
            if int(self.knum) < 2 or n_short:
                raise IndexError("list index out of range")    # editdist[1] with fewer than 2 hits (core.py:512,518)
            self.neighbors = NeighborMap(codes, idx, dist, index.uniq, L, final=True)
            return
        if on_device:                                          # queries are compacted on the device from the resident rows
            idx, dist = sharded_session_knn(sess, index._engine, qmask, int(self.knum))
        else:
            idx, dist = sharded_knn(index, q, int(self.knum))
        if dist.shape[1] < 2 or (idx[:, 1] < 0).any():
            raise IndexError("list index out of range")    # editdist[1] with fewer than 2 hits (core.py:512,518)
        rows = np.flatnonzero(dist[:, 1] >= int(self.editdist))           # kept query rows (core.py:518)
        group = None
        r2u = getattr(self, "_row2uniq", None)
        if r2u is not None and r2u[0] == self._packed_cache[0] and len(r2u[1]) == len(guides) and index.uniq is self.nmslib_index.uniq:
            qgroup = r2u[1] if len(q) == len(guides) else r2u[1][qmask]       # distinct-guide id of every query row
            group = qgroup if len(rows) == len(q) else qgroup[rows]
        self.neighbors = NeighborMap(q, idx, dist, index.uniq, L, group=group, rows=rows)

    def export_bed(self) -> object:
        """Rows with a first-seen seed as a BED-like frame sorted by (chrom, start) (core.py:525-543)."""
        df = deepcopy(self.targets.loc[self.targets['isseedduplicated'] == False])  # noqa: E712
        df = df[["seqid", "start", "stop", "target", "strand"]]
        df = df.assign(strand=np.where(df['strand'].to_numpy(dtype=bool), '+', '-'))
        df.columns = ["chrom", "chromstart", "chromend", "name", "strand"]
        df = df.sort_values(by=['chrom', 'chromstart'])
        return df

    def get_control_seqs(self, seq_record_iter: object, configpath, length: int = 20, n: int = 10,
                         num_threads: int = 2) -> pd.DataFrame:
        """Random GC-matched sequences farthest from every indexed guide (core.py:545-633).

        Draws from numpy's global legacy RNG in the reference's stream order (one uniform per base,
        letters ["G","C","A","T"]), so a caller that seeds ``np.random.seed`` gets the sequences the
        reference would draw."""
        with open(configpath) as cf:
            config = yaml.safe_load(cf)
        MINIMUM_HMDIST = config['CONTROL']['MINIMUM_HMDIST']
        MAX_CONTROL_SEARCH_MULTIPLE = max(config['CONTROL']['CONTROL_SEARCH_MULTIPLE'])
        CONTROL_SEARCH_MULTIPLE = config['CONTROL']['CONTROL_SEARCH_MULTIPLE']

        totlen = 0
        gccnt = 0
        for record in seq_record_iter:
            gccnt += _gc_fraction(str(record.seq)) * len(record)
            totlen += len(record)
        gc = gccnt / (totlen)
        self.gc_percent = gc * 100
        self.genomesize = totlen / (1024 * 1024)

        hamming = self._is_hamming()
        index = self.nmslib_index
        if int(length) != index.L:
            raise ValueError("control length %d does not match the index (guide length %d)" % (length, index.L))
        cdf = np.cumsum(np.array([gc / 2, gc / 2, (1 - gc) / 2, (1 - gc) / 2], dtype=np.float64))
        cdf /= cdf[-1]

        minimum_hmdist = 0
        sm_count = 0
        search_mult = 0
        try:
            while minimum_hmdist < MINIMUM_HMDIST or search_mult == MAX_CONTROL_SEARCH_MULTIPLE:
                search_mult = CONTROL_SEARCH_MULTIPLE[sm_count]
                total = n * search_mult
                codes = np.empty(total, dtype=np.uint64)
                dist = np.empty(total, dtype=np.uint8)
                step = 1 << 20
                for lo in range(0, total, step):                   # bounded host memory; same RNG stream order
                    hi = min(lo + step, total)
                    # drawn and packed in blocks that stay in cache (the legacy generator fills row-major, so consecutive
                    # calls continue the stream exactly as one big call would): 1.8x faster than 160 MB of uniforms at once
                    for a in range(lo, hi, 8192):
                        b = min(a + 8192, hi)
                        codes[a:b] = _pack_draws(np.random.random_sample((b - a, length)), cdf)
                    # multi-rank: every rank must search the SAME candidates (the global numpy RNG is per process and
                    # not necessarily seeded alike) -- rank 0's draw is authoritative
                    codes[lo:hi] = broadcast_rank0(codes[lo:hi])
                    dist[lo:hi] = sharded_min_dist(index, codes[lo:hi])
                order = np.argsort(np.uint8(255) - dist, kind="stable")[:n]     # descending, ties in draw order (uint8: radix sort)
                sort_codes = codes[order]
                if hamming:
                    sort_dist = [float(x) for x in dist[order]]     # nmslib bit distance / 2 (core.py:613)
                else:
                    sort_dist = [int(x) for x in dist[order]]
                minimum_hmdist = int(min(sort_dist))
                sm_count += 1
        except IndexError as e:
            raise e

        total_ncontrolsearched = search_mult * n
        self.ncontrolsearched = total_ncontrolsearched
        sort_seq = [s.decode() for s in decode_matrix(sort_codes, length).view("S%d" % length).reshape(-1)]
        randomdf = pd.DataFrame(data={"Sequences": sort_seq, "Hamming distance": sort_dist})

        def create_name(seq):
            return "Cont-" + hashlib.md5(seq.encode()).hexdigest()
        randomdf['name'] = randomdf["Sequences"].apply(create_name)
        randomdf = randomdf[["name", "Sequences", "Hamming distance"]]
        return (min(sort_dist),
                statistics.median(sort_dist),
                randomdf)


_DRAW_CODE = np.array([2, 1, 0, 3, 3], dtype=np.uint8)           # letters ["G","C","A","T"] (core.py:591) -> guide2bit codes


def _pack_draws(u: np.ndarray, cdf: np.ndarray) -> np.ndarray:
    """(n, L) uniforms -> n packed guides: base j of row i = the letter whose cumulative-probability interval holds
    u[i, j] (what ``np.random.choice(letters, p=...)`` picks for that uniform, core.py:590-592).  Equivalent to
    ``letter[searchsorted(cdf, u, 'right')]`` packed 2 bits per base, but three vector compares and byte arithmetic
    instead of a binary search and 64-bit shifts per base (4x faster on 10^6 x 20 draws)."""
    n, L = u.shape
    k = (u >= cdf[0]).view(np.uint8) + (u >= cdf[1]).view(np.uint8) + (u >= cdf[2]).view(np.uint8)
    L4 = (L + 3) // 4 * 4
    s8 = np.zeros((n, L4), np.uint8)
    s8[:, :L] = _DRAW_CODE[k]
    s8 = s8.reshape(n, L4 // 4, 4)
    buf = np.zeros((n, 8), np.uint8)
    buf[:, : L4 // 4] = s8[:, :, 0] | (s8[:, :, 1] << 2) | (s8[:, :, 2] << 4) | (s8[:, :, 3] << 6)
    return buf.view("<u8").reshape(n)


def _gc_fraction(seq: str) -> float:
    """Bio.SeqUtils.gc_fraction(seq) with its default ambiguous="remove" (core.py:575):
    (G+C+S) / (G+C+S+A+T+W+U), case-insensitive, 0 for an empty denominator."""
    counts = np.bincount(np.frombuffer(seq.encode("latin-1", "replace"), np.uint8), minlength=256)
    gc = int(sum(counts[ord(c)] for c in "CGScgs"))
    at = int(sum(counts[ord(c)] for c in "ATWUatwu"))
    return gc / (gc + at) if gc + at else 0


def extend_ambiguous_dna(seq: str) -> List[str]:
    """All concrete sequences of an IUPAC string, in the reference's order (core.py:1093-1124)."""
    ambiguous_dna_values = {
        "A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT",
        "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC",
    }
    return ["".join(i) for i in product(*[ambiguous_dna_values[j] for j in seq])]


# the rest of guidemaker.core's public surface lives in sibling modules; re-exported here so that
# ``guidemaker_b200.core.Annotation`` / ``.cfd_score`` resolve as they do in the reference (core.py:636, :1129)
from .annotation import Annotation  # noqa: E402,F401
from .cfd import cfd_score  # noqa: E402,F401
