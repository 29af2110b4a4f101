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


def read_fasta(path: str) -> Iterator[Record]:
    name, chunks = None, []
    with _open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    yield Record(name, "".join(chunks).upper())
                fields = line[1:].split()
                name, chunks = (fields[0] if fields else ""), []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        yield Record(name, "".join(chunks).upper())


def read_genbank(path: str) -> Iterator[Record]:
    """Sequence-only GenBank reader: id = VERSION (as Biopython's record.id), else ACCESSION, else LOCUS name."""
    locus = accession = version = None
    in_origin, chunks = False, []
    with _open(path) as f:
        for line in f:
            if in_origin:
                if line.startswith("//"):
                    yield Record(version or accession or locus or "", "".join(chunks).upper())
                    locus = accession = version = None
                    in_origin, chunks = False, []
                else:
                    chunks.append("".join(line.split()[1:]))
            elif line.startswith("LOCUS"):
                parts = line.split()
                locus = parts[1] if len(parts) > 1 else None
            elif line.startswith("ACCESSION"):
                parts = line.split()
                accession = parts[1] if len(parts) > 1 else None
            elif line.startswith("VERSION"):
                parts = line.split()
                version = parts[1] if len(parts) > 1 else None
            elif line.startswith("ORIGIN"):
                in_origin = True
    if in_origin and chunks:
        yield Record(version or accession or locus or "", "".join(chunks).upper())


def get_records(filelist: Iterable[str], input_format: str = "genbank") -> List[Record]:
    """All records of one or more files, in file order, upper-cased (the content of the reference's forward.fasta)."""
    out: List[Record] = []
    for path in filelist:
        out.extend(read_genbank(path) if input_format == "genbank" else read_fasta(path))
    return out
