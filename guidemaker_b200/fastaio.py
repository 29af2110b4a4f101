"""Genome ingest for the hot path: FASTA / GenBank (optionally gzipped) -> upper-cased records.

Stands in for ``guidemaker.core.get_fastas`` + ``Bio.SeqIO.parse`` (core.py:1065-1090, cli.py:161-168): the reference
parses the input with Biopython, upper-cases every record and round-trips it through a temporary FASTA file.  Here
the records are produced directly; only what the hot path needs is parsed (record id + sequence)."""
from __future__ import annotations

import gzip
from typing import Iterable, Iterator, List

from .synth import Record


def is_gzip(filename: str) -> bool:
    """core.py:29-36"""
    with open(filename, "rb") as f:
        return f.read(2) == b"\x1f\x8b"


def _open(path: str):
    return gzip.open(path, "rt") if is_gzip(path) else open(path, "r")


_UPPER = bytes(c - 32 if 97 <= c <= 122 else c for c in range(256))      # str.upper() on ASCII
_WS = b" \t\r\n\x0b\x0c"


class _Seq:
    """Sequence held as upper-case ASCII bytes; ``str()`` decodes on demand (duck-types Bio.Seq for this path).
    ``PamTarget.find_targets`` takes ``_raw`` as is, so a 120 Mb genome is never round-tripped through ``str``."""
    __slots__ = ("_raw", "_str")

    def __init__(self, raw: bytes):
        self._raw, self._str = raw, None

    def __str__(self):
        if self._str is None:
            self._str = self._raw.decode("latin-1")
        return self._str

    def __len__(self):
        return len(self._raw)

    def __eq__(self, other):
        return str(self) == (str(other) if isinstance(other, _Seq) else other)

    def __hash__(self):
        return hash(self._raw)

    def __getitem__(self, key):
        return str(self)[key]

    def upper(self):
        return self


def _read_all(path: str) -> bytes:
    with (gzip.open(path, "rb") if is_gzip(path) else open(path, "rb")) as f:
        return f.read()


def read_fasta(path: str) -> Iterator[Record]:
    """Bulk FASTA reader: header lines are located with ``bytes.find`` (memchr speed), every record body is cleaned by
    ONE ``bytes.translate`` (delete white space + upper-case) -- no Python loop over ~2 million lines.
    Measured on a 120 Mb / 5-record file: 0.45 s plain (was 1.26 s line by line); gzipped input adds the
    single-stream inflate (~1.1 s with zlib), which bounds it."""
    data = _read_all(path)
    starts = []                                            # offsets of the '>' that open a header line
    pos = 0 if data.startswith(b">") else -1
    if pos < 0:
        pos = data.find(b"\n>")
        pos = pos + 1 if pos >= 0 else -1
    while pos >= 0:
        starts.append(pos)
        nxt = data.find(b"\n>", pos + 1)
        pos = nxt + 1 if nxt >= 0 else -1
    starts.append(len(data))
    for a, b in zip(starts[:-1], starts[1:]):
        eol = data.find(b"\n", a, b)
        if eol < 0:
            eol = b
        fields = data[a + 1: eol].split()
        yield Record(fields[0].decode("latin-1") if fields else "", _Seq(data[eol + 1: b].translate(_UPPER, _WS)))


def read_genbank(path: str) -> Iterator[Record]:
    """Sequence-only GenBank reader: id = VERSION (as Biopython's record.id), else ACCESSION, else LOCUS name.  The
    ORIGIN block (numbered, space-separated lower-case lines) is cleaned by ONE ``bytes.translate`` per record."""
    data = _read_all(path)
    for block in data.split(b"\n//"):
        head, sep, origin = block.partition(b"\nORIGIN")
        if not sep:
            continue
        locus = accession = version = None
        stop = head.find(b"\nFEATURES")
        for line in (head if stop < 0 else head[:stop]).split(b"\n"):
            if line.startswith(b"LOCUS"):
                parts = line.split()
                locus = parts[1] if len(parts) > 1 else None
            elif line.startswith(b"ACCESSION") and accession is None:
                parts = line.split()
                accession = parts[1] if len(parts) > 1 else None
            elif line.startswith(b"VERSION") and version is None:
                parts = line.split()
                version = parts[1] if len(parts) > 1 else None
        body = origin.partition(b"\n")[2]                 # drop the rest of the ORIGIN line itself
        rid = version or accession or locus or b""
        yield Record(rid.decode("latin-1"), _Seq(body.translate(_UPPER, _WS + b"0123456789")))


def get_records(filelist: Iterable[str], input_format: str = "genbank") -> List[Record]:
    """All records of one or more files, in file order, upper-cased (the content of the reference's forward.fasta)."""
    out: List[Record] = []
    for path in filelist:
        out.extend(read_genbank(path) if input_format == "genbank" else read_fasta(path))
    return out
