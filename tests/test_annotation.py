"""Feature annotation: GenBank / GFF feature tables, the nearest-feature join that replaces
``bedtools closest``, the seven feature filters and the vectorised guide table.

Pinned by the reference's own known answers on Carsonella (tests/test_core.py:169-246: 7 qualifier keys, 182 CDS,
qualifiers (182, 7), nearby (7074, 12), locus filter) and by hand-computed interval fixtures.  bedtools, pybedtools
and Biopython are absent from this image, so bedtools' output itself cannot be compared."""
import os

import numpy as np
import pandas as pd
import pytest

import guidemaker_b200 as guidemaker
from guidemaker_b200.annotation import Annotation, closest_features, _parse_location
from tests.conftest import GOLDEN, Rec

GBK = os.path.join(GOLDEN, "carsonella.gbk.gz")


def test_parse_location():
    assert _parse_location("1..1317")[:3] == (0, 1317, 1)
    assert _parse_location("complement(1314..2816)")[:3] == (1313, 2816, -1)
    assert _parse_location("join(10..20,30..45)")[:3] == (9, 45, 1)
    assert _parse_location("complement(join(<10..20,30..>45))")[:3] == (9, 45, -1)
    assert _parse_location("join(complement(30..45),complement(10..20))")[:3] == (9, 45, -1)
    assert _parse_location("join(10..20,complement(30..45))")[2] is None
    assert _parse_location("77")[:3] == (76, 77, 1)
    assert _parse_location("1..1317")[3] == "[0:1317](+)"


def test_genbank_features_reference_known_answers(config_yaml):
    """tests/test_core.py:169-181"""
    anno = Annotation(annotation_list=[GBK], annotation_type="genbank", target_bed_df=pd.DataFrame())
    anno.get_annotation_features()
    assert 7 == len(anno.feature_dict)
    assert 182 == len(anno.genbank_bed_df)
    assert list(anno.genbank_bed_df.columns) == ["chrom", "chromStart", "chromEnd", "name", "strand"]
    assert anno.genbank_bed_df["chrom"].unique().tolist() == ["AP009180.1"]
    first = anno.genbank_bed_df.iloc[0]
    assert (first["chromStart"], first["chromEnd"], first["strand"]) == (0, 1317, "+")
    anno._get_qualifiers(configpath=config_yaml)
    assert anno.qualifiers.shape == (182, 7)
    assert "translation" not in anno.qualifiers.columns
    assert anno.locuslen() == ("locus_tag", 182)


def test_closest_features_hand_computed():
    """features (sorted by start): A [10,20)  B [15,40) nested/overlapping  C [60,70)  D [60,65)  E [100,110)"""
    fs, fe = np.array([10, 15, 60, 60, 100]), np.array([20, 40, 70, 65, 110])

    def run(gs, ge, upstream_frame):
        i, d = closest_features(np.array([gs]), np.array([ge]), fs, fe, upstream_frame)
        return int(i[0]), int(d[0])
    # upstream frame (-id): overlap -> distance 0, first overlapping feature in sorted order
    assert run(12, 18, True) == (0, 0)
    assert run(25, 30, True) == (1, 0)
    assert run(19, 61, True) == (0, 0)                         # overlaps A, B, C, D: A is first
    # ... else the nearest feature ending at or before the guide start; book-ended is distance 1, not an overlap
    assert run(40, 45, True) == (1, -1)
    assert run(45, 50, True) == (1, -6)
    assert run(72, 75, True) == (2, -3)                        # by END: C ends 70 (3 away), D ends 65 (8 away)
    assert run(0, 5, True) == (-1, -1)                         # nothing upstream at the contig start
    # downstream frame (-fd): the first feature starting at or behind the guide end; overlaps are not candidates
    assert run(12, 18, False) == (2, 43)                       # inside A and B: the next feature downstream is C (60 - 18 + 1)
    assert run(45, 50, False) == (2, 11)                       # C before D (same start, sorted order)
    assert run(50, 60, False) == (2, 1)                        # book-ended
    assert run(61, 64, False) == (4, 37)                       # inside C and D -> E
    assert run(120, 130, False) == (-1, -1)                    # nothing downstream
    # no features at all
    i, d = closest_features(np.array([1]), np.array([5]), np.array([], int), np.array([], int), False)
    assert (int(i[0]), int(d[0])) == (-1, -1)


def test_closest_features_against_brute_force():
    rng = np.random.default_rng(2)
    fs = np.sort(rng.integers(0, 5000, size=120))
    fe = fs + rng.integers(1, 300, size=120)
    gs = rng.integers(0, 5300, size=600)
    ge = gs + 20
    for upstream_frame in (False, True):
        fi, d = closest_features(gs, ge, fs, fe, upstream_frame)
        for g in range(600):
            best = None                                         # (|distance|, file order) -> smallest wins
            for j in range(len(fs)):
                if fs[j] < ge[g] and fe[j] > gs[g]:
                    dist, side = 0, "o"
                elif fe[j] <= gs[g]:
                    dist, side = gs[g] - fe[j] + 1, "l"
                else:
                    dist, side = fs[j] - ge[g] + 1, "r"
                if (upstream_frame and side == "r") or (not upstream_frame and side != "r"):
                    continue
                key = (dist, j)
                if best is None or key < best[0]:
                    best = (key, j, dist if side == "r" else -dist)
            if best is None:
                assert (fi[g], d[g]) == (-1, -1)
            else:
                assert (fi[g], d[g]) == (best[1], best[2]), (g, upstream_frame)


def _pipeline(carsonella, config_yaml, orientation, knum, enzymes):
    rec = Rec(*carsonella)
    pamobj = guidemaker.PamTarget("NGG", orientation, "hamming")
    targets = pamobj.find_targets(seq_record_iter=[rec], target_len=20)
    tl = guidemaker.TargetProcessor(targets=targets, lsr=10, editdist=2, knum=knum)
    tl.check_restriction_enzymes(enzymes)
    tl.find_unique_near_pam()
    tl.create_index(configpath=config_yaml)
    tl.get_neighbors(configpath=config_yaml)
    tf_df = tl.export_bed()
    anno = Annotation(annotation_list=[GBK], annotation_type="genbank", target_bed_df=tf_df)
    anno.get_annotation_features()
    anno._get_nearby_features()
    return tl, anno


def case_nearby_and_table(carsonella, config_yaml):
    """tests/test_core.py:183-246 on the exact engine"""
    tl, anno = _pipeline(carsonella, config_yaml, "5prime", 10, ['NRAGCA'])
    assert anno.nearby.shape == (7074, 12)                              # tests/test_core.py:200
    assert list(anno.nearby.columns) == ["Accession", "Guide start", "Guide end", "Guide sequence", "Guide strand", "Feature Accession",
                                         "Feature start", "Feature end", "Feature id", "Feature strand", "Feature distance", "direction"]
    nb = anno.nearby
    down, up = nb[nb["direction"] == "downstream"], nb[nb["direction"] == "upstream"]
    assert len(down) == len(up) == 3537
    assert (up["Feature distance"] <= 0).all()                          # -id: overlapping or upstream only (or none: -1)
    has = down["Feature id"].to_numpy() != "."
    assert (down["Feature distance"].to_numpy()[has] > 0).all()         # -fd: strictly downstream
    # distances re-derived from the coordinates
    for fr in (down, up):
        ok = fr["Feature id"].to_numpy() != "."
        g0, g1, f0, f1 = (fr[c].to_numpy()[ok] for c in ("Guide start", "Guide end", "Feature start", "Feature end"))
        gap = np.where((f0 < g1) & (f1 > g0), 0, np.where(f1 <= g0, g0 - f1 + 1, f0 - g1 + 1))
        assert np.array_equal(gap, np.abs(fr["Feature distance"].to_numpy()[ok]))
    anno._filter_features()
    anno._get_qualifiers(configpath=config_yaml)
    anno._format_guide_table(tl)
    p = anno.pretty_df
    # the reference (HNSW, approximate) reports (900, 23); the exact search can only drop guides the HNSW search kept
    # because it missed their nearest neighbour, never add any
    assert p.shape[1] == 23 and 890 <= p.shape[0] <= 900                 # 899: see guidemaker_b200/annotation.py
    assert list(p.columns[:17]) == ['Guide name', 'Guide sequence', 'GC', 'dtype', 'Accession', 'Guide start', 'Guide end', 'Guide strand',
                                    'PAM', 'Feature id', 'Feature start', 'Feature end', 'Feature strand', 'Feature distance',
                                    'Similar guides', 'Similar guide distances', 'target_seq30']
    assert (p['target_seq30'].str.len() == 30).all()
    # the vectorised columns equal the reference's per-row recipes (core.py:897-921)
    import hashlib
    for _, row in p.iloc[:: max(len(p) // 60, 1)].iterrows():
        seq = row['Guide sequence']
        assert row['Guide name'] == hashlib.md5(seq.encode()).hexdigest()
        assert row['GC'] == sum(c in "GC" for c in seq) / len(seq)
        nbr = tl.neighbors[seq]["neighbors"]
        assert row['Similar guide distances'] == ";".join(str(i) for i in nbr["dist"])
        assert row['Similar guides'] == ";".join(nbr["seqs"])
        assert seq in tl.neighbors
        t = tl.targets[(tl.targets['target'] == seq) & (tl.targets['start'] == row['Guide start'] - 1)]
        assert len(t) >= 1 and str(t['exact_pam'].iloc[0]) == str(row['PAM'])
    assert (p['Similar guide distances'].str.split(";").str[1].astype(int) >= 2).all()
    f = anno._filterlocus(attribute='locus_tag', filter_by_locus=['CRP_001'])
    assert f.shape == (4, 23)                                           # tests/test_core.py:246
    assert (f['locus_tag'] == 'CRP_001').all()
    assert anno._filterlocus(attribute='locus_tag').shape == p.shape


def test_nearby_and_table_cpu(oracle_engine, carsonella, config_yaml):
    case_nearby_and_table(carsonella, config_yaml)


@pytest.mark.gpu
def test_nearby_and_table_gpu(cuda_engine, carsonella, config_yaml):
    case_nearby_and_table(carsonella, config_yaml)


def test_gff_features(tmp_path):
    gff = tmp_path / "a.gff"
    lines = ["##gff-version 3",
             "chr1\tsrc\tCDS\t11\t20\t.\t+\t0\tID=cds1;locus_tag=L1;product=alpha beta",
             "chr1\tsrc\tgene\t11\t20\t.\t+\t.\tID=gene1",
             "chr1\tsrc\tCDS\t61\t70\t.\t-\t0\tID=cds2;locus_tag=L2; ;bad"]
    gff.write_text("\n".join(lines) + "\n")
    anno = Annotation([str(gff)], "gff", pd.DataFrame())
    assert anno.check_annotation_type() == "gff"
    anno.get_annotation_features()
    bed = anno.genbank_bed_df
    assert bed["chromStart"].tolist() == ["11", "61"] and bed["strand"].tolist() == ["+", "-"]
    import hashlib
    assert bed["name"].iloc[0] == hashlib.md5((lines[1] + "\n").encode()).hexdigest()     # str(pybedtools.Interval)
    assert set(anno.feature_dict) == {"ID", "locus_tag", "product"}
    assert anno.feature_dict["product"][bed["name"].iloc[0]] == "alpha beta"
    gtf = tmp_path / "a.gtf"
    gtf.write_text("#gtf-version 2.2\nchr1\tsrc\tCDS\t11\t20\t.\t+\t0\tgene_id \"g1\"; transcript_id \"t1\";\n")
    anno = Annotation([str(gtf)], "gff", pd.DataFrame())
    anno.get_annotation_features()
    assert {k: list(v.values()) for k, v in anno.feature_dict.items()} == {"gene_id": ["g1"], "transcript_id": ["t1"]}
