"""API-level parity cases shared by the CPU suite (host logic over an injected oracle engine) and
the GPU suite (the real CUDA engine).  They read like the reference's own tests/test_core.py."""
import numpy as np
import pandas as pd
import pytest

import guidemaker_b200 as guidemaker
from guidemaker_b200._encode import encode_guides
from tests.conftest import Rec

REF_COLUMNS = ["target", "exact_pam", "start", "stop", "strand", "pam_orientation", "target_seq30", "seqid",
               "seedseq", "hasrestrictionsite", "isseedduplicated", "dtype"]


def S(col):
    return np.array([str(x) for x in col], dtype="S")


def assert_frame_matches(df, r, prefix):
    assert list(df.columns) == [c.decode() for c in r[prefix + "columns"]] == REF_COLUMNS
    assert np.array_equal(S(df["target"]), r[prefix + "target"])
    assert np.array_equal(S(df["exact_pam"]), r[prefix + "exact_pam"])
    assert np.array_equal(df["start"].to_numpy(), r[prefix + "start"]) and df["start"].dtype == np.uint32
    assert np.array_equal(df["stop"].to_numpy(), r[prefix + "stop"]) and df["stop"].dtype == np.uint32
    assert np.array_equal(df["strand"].to_numpy(), r[prefix + "strand"]) and df["strand"].dtype == bool
    assert np.array_equal(df["pam_orientation"].to_numpy(), r[prefix + "pam_orientation"])
    assert np.array_equal(S(df["target_seq30"]), r[prefix + "target_seq30"])
    assert np.array_equal(S(df["seqid"]), r[prefix + "seqid"])
    assert str(df["exact_pam"].dtype) == "category" and str(df["seqid"].dtype) == "category" and str(df["dtype"].dtype) == "category"


# ---- reference tests/test_core.py, same call sequences -----------------------------------------------------

def case_pam_attributes():
    assert getattr(guidemaker.core.PamTarget("NGG", "5prime", "hamming"), "pam") == "NGG"          # test_core.py:28-30
    assert getattr(guidemaker.core.PamTarget("GATN", "3prime", "hamming"), "pam_orientation") == "3prime"
    with pytest.raises(AssertionError):
        guidemaker.core.PamTarget("NGZ", "3prime", "hamming")
    with pytest.raises(AssertionError):
        guidemaker.core.PamTarget("NGG", "4prime", "hamming")


def case_find_targets_inline(inline_ref):
    r = inline_ref
    pamobj = guidemaker.core.PamTarget("NGG", "5prime", "hamming")                                   # test_core.py:41-46
    target = pamobj.find_targets(seq_record_iter=[Rec("testseq1", r["t5p/seq"][0].decode())], target_len=6)
    assert target['target'][0] == "ATGCAC"
    assert target['target'][1] == "TAACAA"
    assert_frame_matches(target, r, "t5p/")
    pamobj = guidemaker.core.PamTarget("NGG", "3prime", "hamming")                                   # test_core.py:52-57
    target = pamobj.find_targets(seq_record_iter=[Rec("testseq1", r["t3p/seq"][0].decode())], target_len=6)
    assert target['target'][0] == "ATGATC"
    assert target['target'][1] == "ATTAGA"
    assert_frame_matches(target, r, "t3p/")


CARSONELLA_CASES = {"ngg3p20": ("NGG", "3prime", 20, 10), "ngg5p20": ("NGG", "5prime", 20, 10),
                    "tttv5p23": ("TTTV", "5prime", 23, 10), "nngrrt3p21": ("NNGRRT", "3prime", 21, 12)}


def case_carsonella(name, carsonella, carsonella_ref, config_yaml):
    pam, orient, L, lsr = CARSONELLA_CASES[name]
    r = carsonella_ref
    pamobj = guidemaker.core.PamTarget(pam, orient, "hamming")
    target = pamobj.find_targets(seq_record_iter=[Rec(*carsonella)], target_len=L)
    if name == "ngg5p20":
        assert target['target'][0] == "AAATGGTACGTTATGTGTTA"                                       # test_core.py:59-65
    tl = guidemaker.core.TargetProcessor(targets=target, lsr=lsr, editdist=2, knum=3)
    tl.check_restriction_enzymes(['NRAGCA'])
    tl.find_unique_near_pam()
    assert_frame_matches(tl.targets, r, name + "/")
    assert np.array_equal(S(tl.targets["seedseq"]), r[name + "/seedseq"])
    assert np.array_equal(tl.targets["isseedduplicated"].to_numpy(), r[name + "/isseedduplicated"])
    assert np.array_equal(tl.targets["hasrestrictionsite"].to_numpy(), r[name + "/hasrestrictionsite"])
    assert target["isseedduplicated"].all()          # the caller's frame is not mutated (deepcopy, core.py:414)
    if name + "/nb_keys" in r.files:
        tl.create_index(configpath=config_yaml)
        tl.get_neighbors(configpath=config_yaml)
        keys = list(tl.neighbors.keys())
        assert [k.encode() for k in keys] == r[name + "/nb_keys"].tolist()
        assert np.array_equal(np.array([tl.neighbors[k]["neighbors"]["dist"] for k in keys[:200]]), r[name + "/nb_dist"][:200])
        assert np.array_equal(tl.neighbors.distance_matrix(), r[name + "/nb_dist"])
        bed = tl.export_bed()
        assert bed.shape == (int((~r[name + "/isseedduplicated"]).sum()), 5)
        assert list(bed.columns) == ["chrom", "chromstart", "chromend", "name", "strand"]
        assert bed["chromstart"].is_monotonic_increasing and set(bed["strand"]) == {"+", "-"}


tardict = {'target': ['AAATGGTACGTTATGTGTTA', 'AAATGGTACGTTATGTGTTA', 'AACAGTAAAATGGTTTAATG'],      # test_core.py:67-82
           'exact_pam': ["AGG", "TGG", "CGG"],
           'start': [35, 41, 158572],
           'stop': [55, 61, 158592],
           'strand': [True, True, False],
           'pam_orientation': [False, False, False],
           'target_seq30': ['TTAGGAAATGGTACGTTATGTGTTATAAGA', 'AATGGTACGTTATGTGTTATAAGAATTTCT', 'AACGGAACAGTAAAATGGTTTAATGATACA'],
           'seqid': ['AP009180.1', 'AP009180.2', 'AP009180.1'],
           'seedseq': [np.nan, np.nan, np.nan],
           'isseedduplicated': [np.nan, np.nan, np.nan],
           'hasrestrictionsite': [np.nan, np.nan, np.nan],
           'dtype': ['hamming', 'hamming', 'hamming']}


def handmade_targets():
    targets = pd.DataFrame(tardict)
    return targets.astype({"target": 'str', "exact_pam": 'category', "start": 'uint32', "stop": 'uint32',
                           "strand": 'bool', "pam_orientation": 'bool', "seqid": 'category'})


def case_handmade_frame(config_yaml):
    targets = handmade_targets()
    tl = guidemaker.core.TargetProcessor(targets=targets, lsr=10, editdist=2, knum=2)
    tl.check_restriction_enzymes(['NGGTAB'])                                                        # test_core.py:86-92
    assert tl.targets['hasrestrictionsite'][0] == True  # noqa: E712
    tl.find_unique_near_pam()                                                                       # test_core.py:95-102
    assert tl.targets[tl.targets['isseedduplicated'] == False].shape == (2, 12)  # noqa: E712
    tl = guidemaker.core.TargetProcessor(targets=handmade_targets(), lsr=10, editdist=2, knum=2)
    tl.check_restriction_enzymes(['NRAGCA'])
    tl.find_unique_near_pam()
    tl.create_index(configpath=config_yaml)                                                         # test_core.py:105-112
    tl.get_neighbors(configpath=config_yaml)                                                        # test_core.py:116-126
    assert tl.neighbors["AAATGGTACGTTATGTGTTA"]["neighbors"]["dist"][1] == 12
    assert tl.neighbors["AAATGGTACGTTATGTGTTA"]["neighbors"]["seqs"] == ["AAATGGTACGTTATGTGTTA", "AACAGTAAAATGGTTTAATG"]
    assert tl.neighbors["AAATGGTACGTTATGTGTTA"]["target"] == "AAATGGTACGTTATGTGTTA"
    assert "AAAAAAAAAAAAAAAAAAAA" not in tl.neighbors and len(tl.neighbors) == 2
    tl = guidemaker.core.TargetProcessor(targets=handmade_targets(), lsr=10, editdist=2, knum=10)   # test_core.py:129-139
    tl.check_restriction_enzymes(['NRAGCA'])
    tl.find_unique_near_pam()
    tl.create_index(configpath=config_yaml)
    tl.get_neighbors(configpath=config_yaml)
    df = tl.export_bed()
    assert df.shape == (2, 5)
    assert len(tl) == 3 and "3 potential PAM targets" in str(tl)
    # nmslib protocol facade: one-hot queries in, doubled distances out (core.py:501-503, :512)
    res = tl.nmslib_index.knnQueryBatch(tl._one_hot_encode(["AAATGGTACGTTATGTGTTA"]), k=2, num_threads=2)
    assert res[0][0].tolist() == [0, 1] and res[0][1].tolist() == [0, 24]
    # knum=1 -> editdist[1] does not exist (core.py:512) -> IndexError, as the reference
    tl1 = guidemaker.core.TargetProcessor(targets=handmade_targets(), lsr=10, editdist=2, knum=1)
    tl1.check_restriction_enzymes([]); tl1.find_unique_near_pam(); tl1.create_index(configpath=config_yaml)
    with pytest.raises(IndexError):
        tl1.get_neighbors(configpath=config_yaml)


def case_levin_dist(inline_ref, config_yaml):                                                        # test_core.py:319-347
    distseq = [Rec("distseq", inline_ref["lev/seq"][0].decode())]
    pt_levin = guidemaker.core.PamTarget("NGG", "3prime", "levin")
    pt_hamming = guidemaker.core.PamTarget("NGG", "3prime", "hamming")
    pd_levin = pt_levin.find_targets(seq_record_iter=distseq, target_len=20)
    pd_hamming = pt_hamming.find_targets(seq_record_iter=distseq, target_len=20)
    tp_levin = guidemaker.core.TargetProcessor(targets=pd_levin, lsr=0, editdist=1, knum=3)
    tp_hamming = guidemaker.core.TargetProcessor(targets=pd_hamming, lsr=0, editdist=1, knum=3)
    tp_levin.find_unique_near_pam()
    tp_hamming.find_unique_near_pam()
    tp_hamming.check_restriction_enzymes()
    tp_levin.check_restriction_enzymes()
    tp_levin.create_index(configpath=config_yaml)
    tp_hamming.create_index(configpath=config_yaml)
    tp_levin.get_neighbors(configpath=config_yaml)
    tp_hamming.get_neighbors(configpath=config_yaml)
    assert (tp_levin.neighbors['CTAGTCACTAGCTGACAGCA']['neighbors']['dist'] == [0, 1, 2])
    assert (tp_hamming.neighbors['CTAGTCACTAGCTGACAGCA']['neighbors']['dist'] == [0, 1, 16])
    for tp, key in ((tp_levin, "levin"), (tp_hamming, "hamming")):
        keys = list(tp.neighbors.keys())
        assert [k.encode() for k in keys] == inline_ref["lev/%s_keys" % key].tolist()
        assert np.array_equal(tp.neighbors.distance_matrix(), inline_ref["lev/%s_dist" % key])


SYNTH = ["ngg3p", "ngg5p", "tttv", "nnagaaw", "yg10", "nggnorest"]


def case_synthetic(name, synthetic_ref, config_yaml):
    """multi-record genome with N runs, lower case, a record shorter than the guide, hits at record
    ends, exact repeats; hamming and leven; with and without check_restriction_enzymes (query-mask NaN
    branch, core.py:495)."""
    r = synthetic_ref
    pam, orient, dtype = (x.decode() for x in r[name + "/meta"])
    L, lsr, dist, knum = (int(x) for x in r[name + "/params"])
    recs = [Rec(i.decode(), s.decode()) for i, s in zip(r["rec_ids"], r["rec_seqs"])]
    df = guidemaker.core.PamTarget(pam, orient, dtype).find_targets(recs, L)
    tp = guidemaker.core.TargetProcessor(targets=df, lsr=lsr, editdist=dist, knum=knum)
    if name != "nggnorest":
        tp.check_restriction_enzymes(["GGTCTC", "NGGTAB"])
    tp.find_unique_near_pam()
    assert_frame_matches(tp.targets, r, name + "/")
    assert np.array_equal(S(tp.targets["seedseq"]), r[name + "/seedseq"])
    assert np.array_equal(tp.targets["isseedduplicated"].to_numpy(), r[name + "/isseedduplicated"])
    if name != "nggnorest":
        assert np.array_equal(tp.targets["hasrestrictionsite"].to_numpy(), r[name + "/hasrestrictionsite"])
    tp.create_index(configpath=config_yaml)
    tp.get_neighbors(configpath=config_yaml)
    keys = list(tp.neighbors.keys())
    assert [k.encode() for k in keys] == r[name + "/nb_keys"].tolist()
    ref_dist = r[name + "/nb_dist"]
    got = tp.neighbors.distance_matrix()
    assert got.shape == ref_dist.shape and np.array_equal(got, ref_dist)
    # the neighbour sequences are the true ones: distance of (key, seq) equals the reported distance
    k0 = keys[len(keys) // 2]
    ent = tp.neighbors[k0]["neighbors"]
    from oracle import oracle as O
    f = O.py_hamming if dtype == "hamming" else O.py_leven
    assert [f(k0, s) for s in ent["seqs"]] == ent["dist"]


def case_controls(controls_ref, carsonella, synthetic_ref, config_yaml, tmp_path):
    """get_control_seqs with the legacy global RNG seeded as in make_golden.py: same sequences, same
    distances, same frame as the reference (core.py:545-633)."""
    import yaml
    r = controls_ref
    # hamming, Carsonella, n=100  (test_core.py:144-155 pins shape (100, 3))
    pamobj = guidemaker.core.PamTarget("NGG", "5prime", "hamming")
    targets = pamobj.find_targets(seq_record_iter=[Rec(*carsonella)], target_len=20)
    tl = guidemaker.core.TargetProcessor(targets=targets, lsr=10, editdist=2, knum=10)
    tl.check_restriction_enzymes(['NRAGCA'])
    tl.find_unique_near_pam()
    tl.create_index(configpath=config_yaml)
    np.random.seed(12345)
    data = tl.get_control_seqs([Rec(*carsonella)], length=20, n=100, num_threads=2, configpath=config_yaml)
    assert data[2].shape == (100, 3)
    assert list(data[2].columns) == [c.decode() for c in r["ham/columns"]]
    assert np.array_equal(S(data[2]["Sequences"]), r["ham/seqs"])
    assert np.array_equal(data[2]["Hamming distance"].to_numpy(np.float64), r["ham/dist"])
    assert np.array_equal(S(data[2]["name"]), r["ham/names"])
    assert [float(data[0]), float(data[1])] == r["ham/min_med"].tolist()
    assert tl.ncontrolsearched == int(r["ham/ncontrolsearched"][0])
    assert abs(tl.gc_percent - float(r["ham/gc_percent"][0])) < 1e-9 and abs(tl.genomesize - float(r["ham/genomesize"][0])) < 1e-12
    # leven, small record, MINIMUM_HMDIST lowered to a reachable value as in make_golden.py
    rec4 = Rec("rec4", synthetic_ref["rec_seqs"][3].decode())
    targets = guidemaker.core.PamTarget("NGG", "5prime", "leven").find_targets([rec4], 20)
    tl = guidemaker.core.TargetProcessor(targets=targets, lsr=10, editdist=2, knum=10)
    tl.check_restriction_enzymes(['NRAGCA']); tl.find_unique_near_pam(); tl.create_index(configpath=config_yaml)
    cfg = yaml.safe_load(open(config_yaml)); cfg["CONTROL"]["MINIMUM_HMDIST"] = 5
    p = tmp_path / "lev.yaml"; p.write_text(yaml.safe_dump(cfg))
    np.random.seed(12345)
    cmin, cmed, cdf = tl.get_control_seqs([rec4], configpath=str(p), length=20, n=5)
    assert np.array_equal(S(cdf["Sequences"]), r["lev/seqs"])
    assert np.array_equal(cdf["Hamming distance"].to_numpy(np.float64), r["lev/dist"])
    assert [float(cmin), float(cmed)] == r["lev/min_med"].tolist() and tl.ncontrolsearched == int(r["lev/ncontrolsearched"][0])
    # unreachable threshold -> the reference walks off CONTROL_SEARCH_MULTIPLE and raises IndexError
    cfg["CONTROL"] = {"MINIMUM_HMDIST": 21, "CONTROL_SEARCH_MULTIPLE": [2, 3]}
    p.write_text(yaml.safe_dump(cfg))
    with pytest.raises(IndexError):
        tl.get_control_seqs([rec4], configpath=str(p), length=20, n=5)
    with pytest.raises(ValueError):
        tl.get_control_seqs([rec4], configpath=str(p), length=19, n=5)


def case_errors(config_yaml):
    pamobj = guidemaker.core.PamTarget("NGG", "3prime", "hamming")
    with pytest.raises(ValueError):                       # zero hits -> pd.concat([]) (core.py:286-287)
        pamobj.find_targets([Rec("empty", "ATATATATATATATATATATATATATAT")], 20)
    with pytest.raises(ValueError):
        pamobj.find_targets([], 20)
    t = handmade_targets()
    t.loc[1, "target"] = "AAATGGTACGTTATGTGTT"          # ragged guide lengths are refused, not silently mangled
    tl = guidemaker.core.TargetProcessor(targets=t, lsr=10)
    with pytest.raises(ValueError):
        tl.find_unique_near_pam()
    t = handmade_targets()
    t.loc[1, "target"] = "AAATGGTACGTTATGTGTTN"
    with pytest.raises(ValueError):
        guidemaker.core.TargetProcessor(targets=t, lsr=10).find_unique_near_pam()
    assert guidemaker.core.extend_ambiguous_dna('NGG') == ['GGG', 'AGG', 'TGG', 'CGG']                # test_core.py:254-257
    assert np.array_equal(encode_guides(["ACGT"]), np.array([0b11100100], np.uint64))


def case_unsorted_contigs(config_yaml):
    """Contig ids that are NOT in sorted order (chr2, chr10, chr1): the reference's per-record concat leaves `seqid` as
    plain strings, so export_bed (core.py:525-543) sorts the contigs lexicographically -- chr1, chr10, chr2 -- and rows
    within a contig by start.  The frame itself stays in record order."""
    rng = np.random.default_rng(3)
    recs = [Rec(name, "".join(rng.choice(list("ACGT"), size=4000))) for name in ("chr2", "chr10", "chr1")]
    df = guidemaker.PamTarget("NGG", "3prime", "hamming").find_targets(recs, 20)
    assert list(dict.fromkeys(df["seqid"].astype(str))) == ["chr2", "chr10", "chr1"]          # frame: record order
    # every row's coordinates are record-relative and the guide is what the record holds there
    for i in (0, len(df) // 2, len(df) - 1):
        row = df.iloc[i]
        seq = {r.id: r.seq for r in recs}[str(row["seqid"])]
        piece = seq[int(row["start"]): int(row["stop"])]
        assert row["target"] == (piece if row["strand"] else guidemaker.core._reverse_complement(piece))
    tp = guidemaker.TargetProcessor(df, lsr=10, editdist=2, knum=2)
    tp.find_unique_near_pam()
    bed = tp.export_bed()
    chroms = bed["chrom"].astype(str).tolist()
    assert list(dict.fromkeys(chroms)) == ["chr1", "chr10", "chr2"]                           # lexicographic, as the reference
    assert chroms == sorted(chroms)
    for c in ("chr1", "chr10", "chr2"):
        st = bed.loc[bed["chrom"].astype(str) == c, "chromstart"].to_numpy()
        assert (np.diff(st.astype(np.int64)) >= 0).all()
    # the literal reference recipe on the same rows (plain-string chrom column) gives the same order
    ref = tp.targets.loc[tp.targets["isseedduplicated"] == False, ["seqid", "start", "stop", "target", "strand"]].copy()  # noqa: E712
    ref["seqid"] = ref["seqid"].astype(str)
    ref = ref.sort_values(by=["seqid", "start"])
    assert ref["target"].tolist() == bed["name"].tolist()
