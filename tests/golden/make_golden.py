#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (``python tests/golden/make_golden.py``); it reads
/root/reference, which does not exist on the GPU box.  The fixtures it writes are committed.

What runs for real: ``/root/reference/guidemaker/core.py`` -- loaded unmodified from where it
lies -- i.e. ``PamTarget.find_targets`` (regex scan, hit geometry, row order, dtypes),
``TargetProcessor.find_unique_near_pam`` (seed + keep-first duplicate flag),
``check_restriction_enzymes``, ``create_index`` / ``get_neighbors`` (query mask, dist[1] threshold,
/2 convention, dict shape) and ``get_control_seqs`` (loop, sort, halving).

What is a stand-in (the third-party modules are absent from this image and from
/opt/wheelhouse): ``nmslib`` (replaced by an EXACT brute-force index with the same Python
protocol, spaces ``bit_hamming`` and ``leven`` as published: popcount(xor) / unit-cost edit
distance; ids = insertion positions, ascending (distance, id)), ``Bio`` (a str-based ``Seq`` with
IUPAC ``reverse_complement``, ``SeqRecord``, a FASTA ``SeqIO.parse`` and ``gc_fraction``),
``pybedtools``/``altair``/``onnxruntime`` (never called on this path).  Hence: the PAM-scan and
seed-dedupe fixtures are the reference's output bit for bit; the neighbour fixtures pin the
reference's host logic on top of exact distances (the reference's real HNSW is approximate;
SURVEY.md Appendix A.3 Q14).  Only distances are stored for neighbours: the reference's
``seqs`` are mis-mapped and process-random.
"""
import gzip
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# ------------------------------------------------------------------ stand-ins for absent third-party modules
_COMP = str.maketrans("ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn")


class Seq(str):
    def reverse_complement(self):
        return Seq(str(self).translate(_COMP)[::-1])

    def upper(self):
        return Seq(str(self).upper())


class SeqRecord:
    def __init__(self, seq, id="<unknown id>"):
        self.seq, self.id = seq, id

    def __len__(self):
        return len(self.seq)

    def upper(self):
        return SeqRecord(self.seq.upper(), self.id)


def fasta_parse(handle, fmt="fasta"):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, str):
        handle = gzip.open(handle, "rt") if handle.endswith(".gz") else open(handle)
        close = True
    name, chunks = None, []
    for line in handle:
        if line.startswith(">"):
            if name is not None:
                yield SeqRecord(Seq("".join(chunks)), id=name)
            name, chunks = line[1:].split()[0], []
        else:
            chunks.append(line.strip())
    if name is not None:
        yield SeqRecord(Seq("".join(chunks)), id=name)
    if close:
        handle.close()


def gc_fraction(seq):
    s = str(seq).upper()
    gc = sum(s.count(c) for c in "GCS")
    tot = sum(s.count(c) for c in "ACGTSW")
    return gc / tot if tot else 0.0


class _ExactIndex:
    """nmslib protocol (core.py:451-457, :501-503, :603), exact brute force."""

    def __init__(self, space):
        self.space, self.data = space, []

    def addDataPointBatch(self, data):
        self.data = list(data)

    def createIndex(self, params, print_progress=False):
        if self.space == "bit_hamming":
            self.bits = np.array([np.array(s.split(" "), dtype=np.uint8) for s in self.data], dtype=np.float32)
            self.w = self.bits.sum(1)

    def setQueryTimeParams(self, params):
        pass

    @staticmethod
    def _leven(a, b):
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j - 1] + (ca != cb), prev[j] + 1, cur[j - 1] + 1))
            prev = cur
        return prev[-1]

    def knnQueryBatch(self, queries, k=10, num_threads=0):
        out = []
        if self.space == "bit_hamming":
            q = np.array([np.array(s.split(" "), dtype=np.uint8) for s in queries], dtype=np.float32)
            for lo in range(0, len(q), 4096):
                qq = q[lo:lo + 4096]
                d = (qq.sum(1)[:, None] + self.w[None, :] - 2.0 * (qq @ self.bits.T)).astype(np.int32)   # popcount(xor)
                for row in d:
                    ids = np.lexsort((np.arange(len(row)), row))[:k]
                    out.append((ids.astype(np.int32), row[ids].astype(np.int32)))
        else:
            for s in queries:
                row = np.array([self._leven(s, t) for t in self.data], dtype=np.int32)
                ids = np.lexsort((np.arange(len(row)), row))[:k]
                out.append((ids.astype(np.int32), row[ids].astype(np.int32)))
        return out


def install_stubs():
    nm = types.ModuleType("nmslib")
    nm.DistType = types.SimpleNamespace(INT="INT")
    nm.DataType = types.SimpleNamespace(OBJECT_AS_STRING="OBJECT_AS_STRING")
    nm.init = lambda space, dtype=None, data_type=None, method=None: _ExactIndex(space)
    sys.modules["nmslib"] = nm
    bio = types.ModuleType("Bio")
    bseq = types.ModuleType("Bio.Seq"); bseq.Seq = Seq
    bio.Seq = bseq
    bio.SeqIO = types.ModuleType("Bio.SeqIO"); bio.SeqIO.parse = fasta_parse
    bio.SeqUtils = types.ModuleType("Bio.SeqUtils"); bio.SeqUtils.gc_fraction = gc_fraction
    bio.SeqRecord = types.ModuleType("Bio.SeqRecord"); bio.SeqRecord.SeqRecord = SeqRecord
    for name in ("Seq", "SeqIO", "SeqUtils", "SeqRecord"):
        sys.modules["Bio." + name] = getattr(bio, name)
    sys.modules["Bio"] = bio
    pbt = types.ModuleType("pybedtools"); pbt.BedTool = object
    sys.modules["pybedtools"] = pbt
    sys.modules["altair"] = types.ModuleType("altair")
    gm = types.ModuleType("guidemaker"); gm.__path__ = [os.path.join(REF, "guidemaker")]
    gm.doench_predict = types.ModuleType("guidemaker.doench_predict")
    gm.cfd_score_calculator = types.ModuleType("guidemaker.cfd_score_calculator")
    sys.modules["guidemaker"] = gm
    sys.modules["guidemaker.doench_predict"] = gm.doench_predict
    sys.modules["guidemaker.cfd_score_calculator"] = gm.cfd_score_calculator


def load_reference_core():
    install_stubs()
    spec = importlib.util.spec_from_file_location("guidemaker.core", os.path.join(REF, "guidemaker", "core.py"))
    core = importlib.util.module_from_spec(spec)
    sys.modules["guidemaker.core"] = core
    spec.loader.exec_module(core)
    return core


# ------------------------------------------------------------------ fixture helpers
CONFIG = os.path.join(REF, "guidemaker", "data", "config_default.yaml")


def S(col):
    return np.array([str(x) for x in col], dtype="S")


def frame_arrays(df, prefix=""):
    return {
        prefix + "target": S(df["target"]), prefix + "exact_pam": S(df["exact_pam"]),
        prefix + "start": df["start"].to_numpy(np.uint32), prefix + "stop": df["stop"].to_numpy(np.uint32),
        prefix + "strand": df["strand"].to_numpy(bool), prefix + "pam_orientation": df["pam_orientation"].to_numpy(bool),
        prefix + "target_seq30": S(df["target_seq30"]), prefix + "seqid": S(df["seqid"]),
        prefix + "columns": S(df.columns), prefix + "dtypes": S([str(t) for t in df.dtypes]),
    }


def synthetic_records(rng):
    """Multi-record genome with N runs, lower case, a too-short record, and PAMs at the very ends."""
    recs = []
    for i, n in enumerate([5000, 31, 12, 2500, 800]):
        s = "".join(rng.choice(list("ACGT"), size=n, p=[0.2, 0.3, 0.3, 0.2]))
        if i == 0:
            s = "CCAGG" + s[5:1000] + "N" * 7 + s[1007:2000] + s[2000:2040].lower() + s[2040:3000] + "NGGRY" + s[3005:-3] + "TGG"
            s = s[:4000] + s[100:400] + s[4300:]                       # an exact repeat -> duplicate guides
        if i == 3:
            s = "GG" + s[2:-2] + "CC"
        recs.append(("rec%d" % (i + 1), s))
    return recs


def main():
    core = load_reference_core()
    rng = np.random.default_rng(20261018)

    # ---- 1. Carsonella (config 0 of BASELINE.json): copy the sequence as an input fixture
    fasta = os.path.join(REF, "tests", "test_data", "Carsonella_ruddii.fasta")
    rec = list(fasta_parse(fasta))
    assert len(rec) == 1
    with gzip.GzipFile(os.path.join(OUT, "carsonella.fa.gz"), "wb", mtime=0) as f:
        f.write((">%s\n%s\n" % (rec[0].id, str(rec[0].seq))).encode())

    out = {}
    cases = [("ngg3p20", "NGG", "3prime", 20, 10), ("ngg5p20", "NGG", "5prime", 20, 10),
             ("tttv5p23", "TTTV", "5prime", 23, 10), ("nngrrt3p21", "NNGRRT", "3prime", 21, 12)]
    for name, pam, orient, L, lsr in cases:
        pt = core.PamTarget(pam, orient, "hamming")
        df = pt.find_targets(seq_record_iter=fasta_parse(fasta), target_len=L)
        tp = core.TargetProcessor(targets=df, lsr=lsr, editdist=2, knum=3)
        tp.check_restriction_enzymes(["NRAGCA"])
        tp.find_unique_near_pam()
        d = frame_arrays(tp.targets, name + "/")
        d[name + "/seedseq"] = S(tp.targets["seedseq"])
        d[name + "/isseedduplicated"] = tp.targets["isseedduplicated"].to_numpy(bool)
        d[name + "/hasrestrictionsite"] = tp.targets["hasrestrictionsite"].to_numpy(bool)
        if name in ("ngg3p20", "ngg5p20"):
            tp.create_index(configpath=CONFIG)
            tp.get_neighbors(configpath=CONFIG)
            keys = list(tp.neighbors.keys())
            d[name + "/nb_keys"] = S(keys)
            d[name + "/nb_dist"] = np.array([tp.neighbors[k]["neighbors"]["dist"] for k in keys], dtype=np.int32)
            # (export_bed, core.py:525-543, cannot run under this image's pandas 3: it assigns strings into
            #  a bool column, which pandas >= 2.2 refuses; the reference pins pandas 2.1.1.)
        out.update(d)
        print(name, len(df), int(tp.targets["isseedduplicated"].sum()))
    np.savez_compressed(os.path.join(OUT, "carsonella_ref.npz"), **out)

    # ---- 2. the reference's own inline test sequences (tests/test_core.py:43,54,329-331)
    out = {}
    s5 = "AATGATCTGGATGCACATGCACTGCTCCAAGCTGCATGAAAAGTACAAAGCACGTTATTAGATGGTAACAATGATCTGGATGCACATGCACTGCTCCAAGCTGCATGAAAAGTACAAAGCACGTTATTAGATGGTGGGAAC"
    for name, seq, orient in (("t5p", s5, "5prime"), ("t3p", s5 + "]", "3prime")):
        df = core.PamTarget("NGG", orient, "hamming").find_targets([SeqRecord(Seq(seq), id="testseq1")], 6)
        out.update(frame_arrays(df, name + "/")); out[name + "/seq"] = S([seq])
    dseq = "CGTAGCTAGTCACTAGCTGACAGCAAGGTTTTTCGTAGCTAGACACTAGCTGACAGCAAGGTTTTTTCGTAGCTAGTCACTAGCTGACTAGCAAGG"
    out["lev/seq"] = S([dseq])
    for dtype in ("levin", "hamming"):
        df = core.PamTarget("NGG", "3prime", dtype).find_targets([SeqRecord(Seq(dseq), id="distseq")], 20)
        tp = core.TargetProcessor(targets=df, lsr=0, editdist=1, knum=3)
        tp.find_unique_near_pam(); tp.check_restriction_enzymes()
        tp.create_index(configpath=CONFIG); tp.get_neighbors(configpath=CONFIG)
        keys = list(tp.neighbors.keys())
        out["lev/%s_keys" % dtype] = S(keys)
        out["lev/%s_dist" % dtype] = np.array([tp.neighbors[k]["neighbors"]["dist"] for k in keys], dtype=np.int32)
        out.update(frame_arrays(tp.targets, "lev/%s_" % dtype))
    assert out["lev/levin_dist"][list(out["lev/levin_keys"]).index(b"CTAGTCACTAGCTGACAGCA")].tolist() == [0, 1, 2]
    assert out["lev/hamming_dist"][list(out["lev/hamming_keys"]).index(b"CTAGTCACTAGCTGACAGCA")].tolist() == [0, 1, 16]
    np.savez_compressed(os.path.join(OUT, "inline_ref.npz"), **out)

    # ---- 3. multi-record synthetic genome with edge cases, several PAMs/orientations, hamming + leven neighbours
    recs = synthetic_records(rng)
    out = {"rec_ids": S([r[0] for r in recs]), "rec_seqs": S([r[1] for r in recs])}
    for name, pam, orient, L, lsr, dtype, dist, knum in [
            ("ngg3p", "NGG", "3prime", 20, 10, "hamming", 2, 5), ("ngg5p", "NGG", "5prime", 20, 0, "hamming", 3, 4),
            ("tttv", "TTTV", "5prime", 23, 23, "leven", 2, 3), ("nnagaaw", "NNAGAAW", "3prime", 27, 5, "hamming", 1, 2),
            ("yg10", "YG", "5prime", 10, 4, "leven", 1, 6), ("nggnorest", "NGG", "3prime", 17, 8, "hamming", 2, 20)]:
        pt = core.PamTarget(pam, orient, dtype)
        df = pt.find_targets([SeqRecord(Seq(s), id=i) for i, s in recs], L)
        tp = core.TargetProcessor(targets=df, lsr=lsr, editdist=dist, knum=knum)
        if name != "nggnorest":       # exercises the NaN branch of the query mask (core.py:495)
            tp.check_restriction_enzymes(["GGTCTC", "NGGTAB"])
        tp.find_unique_near_pam()
        d = frame_arrays(tp.targets, name + "/")
        d[name + "/seedseq"] = S(tp.targets["seedseq"])
        d[name + "/isseedduplicated"] = tp.targets["isseedduplicated"].to_numpy(bool)
        if name != "nggnorest":
            d[name + "/hasrestrictionsite"] = tp.targets["hasrestrictionsite"].to_numpy(bool)
        tp.create_index(configpath=CONFIG)
        tp.get_neighbors(configpath=CONFIG)
        keys = list(tp.neighbors.keys())
        d[name + "/nb_keys"] = S(keys)
        d[name + "/nb_dist"] = np.array([tp.neighbors[k]["neighbors"]["dist"] for k in keys], dtype=np.int32)
        d[name + "/params"] = np.array([L, lsr, dist, knum])
        d[name + "/meta"] = S([pam, orient, dtype])
        out.update(d)
        print(name, len(df), len(keys))
    np.savez_compressed(os.path.join(OUT, "synthetic_ref.npz"), **out)

    # ---- 4. control sequences: reference loop with the legacy global RNG seeded (core.py:545-633)
    out = {}
    for name, dtype in (("ham", "hamming"), ("lev", "leven")):
        pt = core.PamTarget("NGG", "5prime", dtype)
        src = fasta if name == "ham" else None
        if name == "ham":
            df = pt.find_targets(seq_record_iter=fasta_parse(fasta), target_len=20)
            it = fasta_parse(fasta)
            n = 100
        else:
            small = [SeqRecord(Seq(recs[3][1]), id="rec4")]
            df = pt.find_targets(seq_record_iter=small, target_len=20)
            it = small
            n = 5
        tp = core.TargetProcessor(targets=df, lsr=10, editdist=2, knum=10)
        tp.check_restriction_enzymes(["NRAGCA"]); tp.find_unique_near_pam(); tp.create_index(configpath=CONFIG)
        np.random.seed(12345)
        import yaml, tempfile
        cfgp = CONFIG
        if name == "lev":     # a reachable threshold for edit distance on a tiny target set
            cfg = yaml.safe_load(open(CONFIG)); cfg["CONTROL"]["MINIMUM_HMDIST"] = 5
            tf = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False); yaml.safe_dump(cfg, tf); tf.close(); cfgp = tf.name
        cmin, cmed, cdf = tp.get_control_seqs(it, configpath=cfgp, length=20, n=n)
        out[name + "/min_med"] = np.array([cmin, cmed], dtype=np.float64)
        out[name + "/seqs"] = S(cdf["Sequences"]); out[name + "/dist"] = cdf["Hamming distance"].to_numpy(np.float64)
        out[name + "/names"] = S(cdf["name"]); out[name + "/columns"] = S(cdf.columns)
        out[name + "/ncontrolsearched"] = np.array([tp.ncontrolsearched])
        out[name + "/gc_percent"] = np.array([tp.gc_percent]); out[name + "/genomesize"] = np.array([tp.genomesize])
        out[name + "/targets"] = S(tp.targets["target"])
        print("controls", name, cmin, cmed, tp.ncontrolsearched)
        del src
    np.savez_compressed(os.path.join(OUT, "controls_ref.npz"), **out)


if __name__ == "__main__":
    main()
