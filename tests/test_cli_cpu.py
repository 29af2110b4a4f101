"""Ingest + CLI runner on CPU (C-ABI calls routed to the oracle by the fixture): FASTA and GenBank give the same
records as the reference fixture, and the raw-guide table has the reference's header and row set."""
import gzip
import os

import numpy as np
import pandas as pd

from guidemaker_b200 import cli, fastaio
from tests.conftest import GOLDEN


def _write_genbank(path, rec_id, seq):
    with open(path, "w") as f:
        f.write("LOCUS       TEST%s %d bp    DNA     circular BCT 01-JAN-2000\n" % (rec_id.split(".")[0][-4:], len(seq)))
        f.write("ACCESSION   %s\nVERSION     %s\nFEATURES             Location/Qualifiers\nORIGIN\n" % (rec_id.split(".")[0], rec_id))
        s = seq.lower()
        for i in range(0, len(s), 60):
            f.write("%9d %s\n" % (i + 1, " ".join(s[i + j: i + j + 10] for j in range(0, 60, 10) if s[i + j: i + j + 10])))
        f.write("//\n")


def test_fasta_and_genbank_ingest(tmp_path, carsonella):
    rec_id, seq = carsonella
    recs = fastaio.get_records([os.path.join(GOLDEN, "carsonella.fa.gz")], "fasta")        # gzipped FASTA
    assert [(r.id, len(r)) for r in recs] == [(rec_id, 159662)] and recs[0].seq == seq
    gb = tmp_path / "c.gbk"
    _write_genbank(gb, rec_id, seq)
    recs = fastaio.get_records([str(gb)], "genbank")
    assert recs[0].id == rec_id and recs[0].seq == seq                                        # lower case in the file -> upper-cased
    fa = tmp_path / "two.fa"
    fa.write_text(">r1 first record\nacgtn\nACGT\n>r2\nGG\n")
    recs = fastaio.get_records([str(fa)], "fasta")
    assert [(r.id, r.seq) for r in recs] == [("r1", "ACGTNACGT"), ("r2", "GG")]
    assert fastaio.is_gzip(os.path.join(GOLDEN, "carsonella.fa.gz")) and not fastaio.is_gzip(str(fa))


def test_cli_raw_output_and_offtargets(oracle_engine, tmp_path, carsonella_ref):
    out = tmp_path / "out"
    cli.main(["--fasta", os.path.join(GOLDEN, "carsonella.fa.gz"), "--pamseq", "NGG", "--outdir", str(out), "--pam_orientation", "3prime",
              "--guidelength", "20", "--lsr", "10", "--dist", "2", "--knum", "3", "--controls", "0", "--log", str(tmp_path / "log.txt"),
              "--restriction_enzyme_list", "NRAGCA"])
    raw = pd.read_csv(out / "rawguides.csv.gz")
    assert list(raw.columns) == ["Chromosome", "Start", "Stop", "gRNA", "Strand"]              # cli.py:191
    r = carsonella_ref
    keep = ~r["ngg3p20/isseedduplicated"]
    assert len(raw) == int(keep.sum())
    assert sorted(raw["gRNA"]) == sorted(x.decode() for x in r["ngg3p20/target"][keep])
    assert raw["Start"].is_monotonic_increasing and set(raw["Strand"]) == {"+", "-"}
    off = pd.read_csv(out / "offtargets.csv.gz")
    ref_dist = {k.decode(): ";".join(str(int(x)) for x in d) for k, d in zip(r["ngg3p20/nb_keys"], r["ngg3p20/nb_dist"])}
    assert set(off["Guide sequence"]) <= set(ref_dist)
    assert all(ref_dist[g] == d for g, d in zip(off["Guide sequence"], off["Similar guide distances"]))
    assert (off["Similar guides"].str.split(";").str[0] == off["Guide sequence"]).all()        # the self hit comes first
    assert (off["Guide end"] - off["Guide start"] + 1 == 20).all()
    cli.main(["--fasta", os.path.join(GOLDEN, "carsonella.fa.gz"), "--pamseq", "NGG", "--outdir", str(tmp_path / "raw"), "--raw_output_only",
              "--controls", "0", "--log", str(tmp_path / "log.txt")])
    assert os.listdir(tmp_path / "raw") == ["rawguides.csv.gz"]


def test_cli_targets_table_from_genbank(oracle_engine, tmp_path):
    """--genbank: the guide table targets.csv.gz (cli.py:195-227) with the reference's 23 columns and 1-based starts"""
    out = tmp_path / "out"
    cli.main(["--genbank", os.path.join(GOLDEN, "carsonella.gbk.gz"), "--pamseq", "NGG", "--outdir", str(out), "--pam_orientation", "5prime",
              "--guidelength", "20", "--lsr", "10", "--dist", "2", "--knum", "10", "--controls", "0", "--log", str(tmp_path / "log.txt"),
              "--restriction_enzyme_list", "NRAGCA"])
    t = pd.read_csv(out / "targets.csv.gz")
    assert t.shape == (899, 23)                      # reference: (900, 23) with its approximate search (tests/test_core.py:222)
    assert list(t.columns[:9]) == ['Guide name', 'Guide sequence', 'GC', 'dtype', 'Accession', 'Guide start', 'Guide end', 'Guide strand', 'PAM']
    assert {"locus_tag", "product", "protein_id"} <= set(t.columns)
    assert (t["Guide end"] - t["Guide start"] + 1 == 20).all() and (t["Accession"] == "AP009180.1").all()
    assert t["Feature start"].is_monotonic_increasing
    out2 = tmp_path / "out2"
    cli.main(["--genbank", os.path.join(GOLDEN, "carsonella.gbk.gz"), "--pamseq", "NGG", "--outdir", str(out2), "--pam_orientation", "5prime",
              "--knum", "10", "--controls", "0", "--log", str(tmp_path / "log.txt"), "--restriction_enzyme_list", "NRAGCA",
              "--attribute_key", "locus_tag", "--filter_by_attribute", "CRP_001"])
    assert pd.read_csv(out2 / "targets.csv.gz").shape == (4, 23)                      # tests/test_core.py:246


def test_bulk_readers_edge_cases(tmp_path):
    """the bulk byte-level readers give what a line-by-line reader (and Biopython's upper-cased records) would"""
    fa = tmp_path / "x.fa"
    fa.write_bytes(b"; junk before the first header\n>r1 desc\r\nacgt nn\r\n\r\nACGT\r\n>r2\n>r3 empty above\nGG>A\n")
    recs = fastaio.get_records([str(fa)], "fasta")
    assert [(r.id, str(r.seq)) for r in recs] == [("r1", "ACGTNNACGT"), ("r2", ""), ("r3", "GG>A")]
    assert len(recs[0]) == 10 and recs[0].seq == "ACGTNNACGT" and recs[0].seq[2:5] == "GTN"
    gz = tmp_path / "x.fa.gz"
    gz.write_bytes(gzip.compress(fa.read_bytes()))
    assert [(r.id, str(r.seq)) for r in fastaio.get_records([str(gz)], "fasta")] == [(r.id, str(r.seq)) for r in recs]
    gb = tmp_path / "two.gbk"
    _write_genbank(gb, "AB000001.1", "ACGTACGTAC" * 13 + "GGG")
    first = gb.read_text()
    _write_genbank(gb, "AB000002.2", "TTTTGGGGCCCCAAAA")
    gb.write_text(first + gb.read_text())
    recs = fastaio.get_records([str(gb)], "genbank")
    assert [(r.id, len(r)) for r in recs] == [("AB000001.1", 133), ("AB000002.2", 16)]
    assert str(recs[0].seq) == "ACGTACGTAC" * 13 + "GGG" and str(recs[1].seq) == "TTTTGGGGCCCCAAAA"
    empty = tmp_path / "none.fa"
    empty.write_text("no header here\n")
    assert fastaio.get_records([str(empty)], "fasta") == []
