"""Feature annotation of the guide table: ``guidemaker.core.Annotation`` (core.py:636-983) without Biopython,
pybedtools or the ``bedtools`` binary (none of them is in this image).

Same class, method names, attributes and column names as the reference:

    get_annotation_features   core.py:690-772   GenBank feature table / GFF / GTF  ->  genbank_bed_df + feature_dict
    _get_qualifiers           core.py:775-813
    _get_nearby_features      core.py:815-848   ``bedtools sort`` + ``closest -d -fd -D a -t first`` / ``-d -id -D a -t first``
    _filter_features          core.py:851-886
    _format_guide_table       core.py:888-948   (vectorised: no per-row ``.apply``)
    _filterlocus / locuslen   core.py:950-983

The nearest-feature join is a sorted-interval search (``numpy.searchsorted``) restating what the reference's two
``bedtools closest`` calls deliver IN THE REFERENCE'S PIPELINE:

* both BED frames have FIVE columns (chrom, start, end, name, strand), so bedtools reads the strand letters as the
  BED *score* column and treats every interval as unstranded: ``-D a`` orients everything as if the guide were on
  the + strand (upstream = lower coordinates) whatever the guide's real strand is;
* ``-d -fd -D a -t first`` ("downstream" frame): the first feature DOWNSTREAM of the guide -- the nearest feature that
  starts at or behind the guide's end; overlapping and upstream features are not candidates; distance = gap + 1 > 0;
* ``-d -id -D a -t first`` ("upstream" frame): downstream features are ignored -- an overlapping feature (distance 0,
  the first one in sorted order) or else the nearest feature that ends at or before the guide's start (distance
  -(gap + 1) < 0);
* a guide without an eligible feature gets ``.``/-1 columns (one output row per guide and frame either way).

bedtools itself cannot be run here.  The semantics above were pinned against the reference's own known answers on
Carsonella (``tests/test_core.py:183-246``): of the 16 combinations of {strand-aware | strand-blind} x {downstream frame
keeps | drops overlaps} x {upstream frame keeps | drops overlaps} x {downstream frame = closest overall | downstream
only}, exactly this one reproduces ``nearby.shape == (7074, 12)``, the locus filter ``(4, 23)`` and the guide table
``(900, 23)`` -- the exact search yields 899 rows, one fewer, which is the one direction an exact search can differ from
the reference's approximate HNSW search (a guide the HNSW search kept because it missed a neighbour at distance 1).
Hand-computed interval fixtures and a brute-force cross-check are in tests/test_annotation.py.

Feature ids: the reference hashes Biopython's ``str(SeqFeature)`` (core.py:721) or pybedtools' ``str(Interval)``
(core.py:739).  The GFF form (the tab-joined line) is reproduced exactly; the GenBank form follows Biopython's
``SeqFeature.__str__`` layout but cannot be checked against Biopython in this image.
"""
from __future__ import annotations

import gzip
import hashlib
import logging
import re
from copy import deepcopy
from typing import Dict, List

import numpy as np
import pandas as pd
import yaml

from .fastaio import is_gzip

logger = logging.getLogger(__name__)


def _open_text(path: str):
    return gzip.open(path, "rt") if is_gzip(path) else open(path, "r")


# ---- GenBank feature table ---------------------------------------------------------------------------------------------
_LOC_NUM = re.compile(r"[<>]?(\d+)")


def _parse_location(loc: str):
    """GenBank location string -> (start0, end, strand, text) as Biopython's FeatureLocation reports them:
    0-based start = smallest coordinate - 1, end = largest coordinate, strand -1 iff complemented (mixed -> None)."""
    s = loc.replace(" ", "")
    n_comp = s.count("complement(")
    parts = re.findall(r"(complement\()?[<>]?\d+(?:\.\.[<>]?\d+|\^[<>]?\d+)?", s)
    nums = [int(x) for x in _LOC_NUM.findall(s)]
    if not nums:
        raise ValueError("unparsable location %r" % loc)
    start0, end = min(nums) - 1, max(nums)
    if n_comp == 0:
        strand = 1
    elif s.startswith("complement(") or n_comp == len(parts):
        strand = -1
    else:
        strand = None                                        # parts on both strands
    sign = {1: "+", -1: "-", None: "?"}[strand]
    pieces = re.findall(r"([<>]?)(\d+)(?:\.\.([<>]?)(\d+))?", s)
    if len(pieces) == 1:
        b, a, c, d = pieces[0]
        text = "[%s%d:%s%s](%s)" % (b, int(a) - 1, c, d if d else a, sign)
    else:
        inner = ", ".join("[%s%d:%s%s](%s)" % (b, int(a) - 1, c, d if d else a, sign) for b, a, c, d in pieces)
        text = "join{%s}" % inner
    return start0, end, strand, text


def _feature_string(ftype: str, loc_text: str, qualifiers: Dict[str, List[str]]) -> str:
    """layout of Biopython's SeqFeature.__str__ (type, location, qualifiers sorted by key)"""
    out = "type: %s\nlocation: %s\nqualifiers:\n" % (ftype, loc_text)
    for key in sorted(qualifiers):
        out += "    Key: %s, Value: %s\n" % (key, qualifiers[key])
    return out


def read_genbank_features(path: str):
    """-> iterator of (record id, feature type, start0, end, strand, qualifiers, location text) over all records.

    Record id = VERSION (as Biopython's ``record.id``), else ACCESSION, else the LOCUS name.  Qualifier values are lists
    of strings as in Biopython (quotes stripped, continuation lines joined with a space -- without one for
    ``translation``); a valueless qualifier (``/pseudo``) has the value ``['']``."""
    locus = accession = version = None
    in_features = False
    cur = None                              # [type, location string, [qualifier lines]]

    def flush(cur):
        ftype, loc, qlines = cur
        quals: Dict[str, List[str]] = {}
        key, val = None, None
        items = []
        for ln in qlines:
            if ln.startswith("/") and (val is None or val.count('"') % 2 == 0):
                if key is not None:
                    items.append((key, val))
                if "=" in ln:
                    key, val = ln[1:].split("=", 1)
                else:
                    key, val = ln[1:], ""
            elif key is not None:
                val += ("" if key == "translation" else " ") + ln
        if key is not None:
            items.append((key, val))
        for k, v in items:
            v = v.strip()
            if len(v) >= 2 and v[0] == '"' and v[-1] == '"':
                v = v[1:-1].replace('""', '"')
            quals.setdefault(k, []).append(v)
        start0, end, strand, text = _parse_location(loc)
        return (version or accession or locus or "", ftype, start0, end, strand, quals, text)

    with _open_text(path) as f:
        for line in f:
            if line.startswith("LOCUS"):
                parts = line.split()
                locus, accession, version = (parts[1] if len(parts) > 1 else None), None, None
                in_features, cur = False, None
            elif line.startswith("ACCESSION") and not in_features:
                parts = line.split()
                accession = parts[1] if len(parts) > 1 else None
            elif line.startswith("VERSION") and not in_features:
                parts = line.split()
                version = parts[1] if len(parts) > 1 else None
            elif line.startswith("FEATURES"):
                in_features = True
            elif in_features and (line.startswith("ORIGIN") or line.startswith("CONTIG") or line.startswith("//") or
                                  (line[:1] not in (" ", "\n") and not line.startswith("     "))):
                if cur is not None:
                    yield flush(cur)
                    cur = None
                in_features = False
            elif in_features:
                key = line[5:21].strip() if len(line) > 5 else ""
                body = line[21:].strip()
                if line[:5] == "     " and key and not line[5:6].isspace():
                    if cur is not None:
                        yield flush(cur)
                    cur = [key, body, []]
                elif cur is not None and body:
                    if not cur[2] and not body.startswith("/"):
                        cur[1] += body                      # location continued on the next line
                    else:
                        cur[2].append(body)
        if cur is not None and in_features:
            yield flush(cur)


class Annotation:
    """Annotation class for data and methods on targets and gene annotations (core.py:636-983)."""

    def __init__(self, annotation_list: List[str], annotation_type: str, target_bed_df: object) -> None:
        self.annotation_list: List[str] = annotation_list
        self.annotation_type = annotation_type
        self.target_bed_df: object = target_bed_df
        self.genbank_bed_df: object = None
        self.feature_dict: Dict = None
        self.nearby: object = None
        self.filtered_df: object = None
        self.qualifiers: object = None

    def check_annotation_type(self):
        """'gff' or 'gtf' from the first line of the first file (core.py:665-688)"""
        with _open_text(self.annotation_list[0]) as f:
            line1 = f.readline()
        if re.search("gff-version", line1) is not None:
            return "gff"
        if re.search("gtf-version", line1) is not None:
            return "gtf"
        logger.error("Could not verify the GFF/GTF file type. Please make sure your GFF/GTF file starts with '#gtf-version' or '##gff-version'")
        raise ValueError

    def get_annotation_features(self, feature_types: List[str] = None) -> None:
        """Parse annotation records into a BED-like frame and a dict of qualifiers (core.py:690-772)."""
        if feature_types is None:
            feature_types = ["CDS"]
        feature_dict: Dict = {}
        pddict = dict(chrom=[], chromStart=[], chromEnd=[], name=[], strand=[])
        if self.annotation_type == "genbank":
            for gbfile in self.annotation_list:
                try:
                    feats = read_genbank_features(gbfile)
                    for rec_id, ftype, start0, end, strand, quals, loc_text in feats:
                        if ftype not in feature_types:
                            continue
                        if strand in (1, -1):
                            pddict["strand"].append("-" if strand == -1 else "+")
                        featid = hashlib.md5(_feature_string(ftype, loc_text, quals).encode()).hexdigest()
                        pddict["chrom"].append(rec_id)
                        pddict["chromStart"].append(int(start0))
                        pddict["chromEnd"].append(int(end))
                        pddict["name"].append(featid)
                        for qualifier_key, qualifier_val in quals.items():
                            feature_dict.setdefault(qualifier_key, {})[featid] = qualifier_val
                except IOError as e:
                    logger.error("The genbank file %s could not be opened" % gbfile)
                    raise e
        elif self.annotation_type == "gff":
            anno_format = self.check_annotation_type()
            for gff in self.annotation_list:
                with _open_text(gff) as f:
                    for line in f:
                        if line.startswith("#") or not line.strip():
                            continue
                        rec = line.rstrip("\n").split("\t")
                        if len(rec) < 9 or rec[2] not in feature_types:
                            continue
                        pddict["chrom"].append(rec[0])
                        pddict["chromStart"].append(rec[3])
                        pddict["chromEnd"].append(rec[4])
                        pddict["strand"].append(rec[6])
                        featid = hashlib.md5(("\t".join(rec) + "\n").encode()).hexdigest()     # str(pybedtools.Interval)
                        pddict["name"].append(featid)
                        for feat in rec[8].split(';'):
                            try:
                                if feat.isspace() or not feat:
                                    continue
                                if anno_format == 'gtf':
                                    fl = re.search('^[^"]*', feat)
                                    fv = re.search('"([^"]*)"', feat)
                                    feat_key = fl.group(0).strip()
                                    feat_val = fv.group(0).strip('"')
                                else:
                                    fl = feat.split('=')
                                    feat_key = fl[0]
                                    feat_val = fl[1]
                                feature_dict.setdefault(feat_key, {})[featid] = feat_val
                            except Exception:  # noqa: BLE001 -- the reference skips malformed attributes with a warning
                                logger.warning("There appears to be an error in the formatting of an attribute in the "
                                               "record below. Please check your input GFF or GTF file. The record is: {rec} "
                                               "and the attribute is: {att}. Skipping this feature.".format(rec=rec[8].split(';'), att=feat))
                                continue
        self.genbank_bed_df = pd.DataFrame.from_dict(pddict)
        self.feature_dict = feature_dict

    def _get_qualifiers(self, configpath, excluded: List[str] = None) -> object:
        """Frame of features x qualifier values for the qualifiers over the MINIMUM_PROPORTION threshold (core.py:775-813)."""
        with open(configpath) as cf:
            config = yaml.safe_load(cf)
        min_prop = config['MINIMUM_PROPORTION']
        if excluded is None:
            excluded = ["translation"]
        final_quals = []
        qual_df = pd.DataFrame(data={"Feature id": []})
        for featkey, quals in self.feature_dict.items():
            if len(quals) / len(self.feature_dict[featkey]) > min_prop:      # (sic) always 1 > min_prop, as in the reference
                final_quals.append(featkey)
        for qualifier in final_quals:
            if qualifier not in excluded:
                featlist, quallist = [], []
                for feat, qual in self.feature_dict[qualifier].items():
                    featlist.append(feat)
                    quallist.append(";".join([str(i) for i in qual]) if isinstance(qual, list) else qual)
                tempdf = pd.DataFrame({'Feature id': featlist, qualifier: quallist})
                qual_df = qual_df.merge(tempdf, how="outer", on="Feature id")
        self.qualifiers = qual_df

    # ---- nearest-feature join ---------------------------------------------------------------------------------------------
    def _get_nearby_features(self) -> None:
        """First feature downstream of every guide ("downstream" frame) and the overlapping-or-nearest-upstream feature
        ("upstream" frame), core.py:815-848; see the module docstring for the bedtools semantics restated here."""
        feat = self.genbank_bed_df
        tb = self.target_bed_df
        f_chrom = feat["chrom"].astype(str).to_numpy()
        f_start = pd.to_numeric(feat["chromStart"]).to_numpy(dtype=np.int64)
        f_end = pd.to_numeric(feat["chromEnd"]).to_numpy(dtype=np.int64)
        f_name = feat["name"].astype(str).to_numpy()
        f_strand = feat["strand"].astype(str).to_numpy()
        g_chrom = tb.iloc[:, 0].astype(str).to_numpy()
        g_start = pd.to_numeric(tb.iloc[:, 1]).to_numpy(dtype=np.int64)
        g_end = pd.to_numeric(tb.iloc[:, 2]).to_numpy(dtype=np.int64)
        g_name = tb.iloc[:, 3].astype(str).to_numpy()
        g_strand = tb.iloc[:, 4].astype(str).to_numpy()
        # bedtools sort: chromosome (lexicographic), then start; stable for equal keys
        g_order = np.lexsort((g_start, g_chrom))
        g_chrom, g_start, g_end, g_name, g_strand = g_chrom[g_order], g_start[g_order], g_end[g_order], g_name[g_order], g_strand[g_order]
        f_order = np.lexsort((f_start, f_chrom))
        f_chrom, f_start, f_end, f_name, f_strand = f_chrom[f_order], f_start[f_order], f_end[f_order], f_name[f_order], f_strand[f_order]

        n = len(g_start)
        res = {False: (np.full(n, -1, np.int64), np.full(n, -1, np.int64)), True: (np.full(n, -1, np.int64), np.full(n, -1, np.int64))}
        for chrom in np.unique(g_chrom):
            gsel = np.flatnonzero(g_chrom == chrom)
            fsel = np.flatnonzero(f_chrom == chrom)
            if len(fsel) == 0:
                continue
            for upstream_frame in (False, True):
                fi, d = closest_features(g_start[gsel], g_end[gsel], f_start[fsel], f_end[fsel], upstream_frame)
                res[upstream_frame][0][gsel] = np.where(fi >= 0, fsel[np.maximum(fi, 0)], -1)
                res[upstream_frame][1][gsel] = d

        def frame(fi, d, direction):
            has = fi >= 0
            j = np.maximum(fi, 0)
            return pd.DataFrame({
                "Accession": g_chrom, "Guide start": g_start, "Guide end": g_end, "Guide sequence": g_name, "Guide strand": g_strand,
                "Feature Accession": np.where(has, f_chrom[j] if len(f_chrom) else ".", "."),
                "Feature start": np.where(has, f_start[j] if len(f_start) else -1, -1),
                "Feature end": np.where(has, f_end[j] if len(f_end) else -1, -1),
                "Feature id": np.where(has, f_name[j] if len(f_name) else ".", "."),
                "Feature strand": np.where(has, f_strand[j] if len(f_strand) else ".", "."),
                "Feature distance": np.where(has, d, -1), "direction": direction})
        downstream = frame(*res[False], "downstream")
        upstream = frame(*res[True], "upstream")
        self.nearby = pd.concat([downstream, upstream], axis=0)

    def _filter_features(self, before_feat: int = 100, after_feat: int = 200) -> None:
        """Keep guides close enough to a feature start to interact (core.py:851-886); same seven selections, same order."""
        nb = self.nearby
        gs, fs = nb["Guide strand"].to_numpy(), nb["Feature strand"].to_numpy()
        dist = nb["Feature distance"].to_numpy()
        g0, g1 = nb["Guide start"].to_numpy(), nb["Guide end"].to_numpy()
        f0, f1 = nb["Feature start"].to_numpy(), nb["Feature end"].to_numpy()
        plus_g, minus_g, plus_f, minus_f = gs == "+", gs == "-", fs == "+", fs == "-"
        masks = [
            (gs == fs) & (0 < dist) & (dist < before_feat),
            plus_g & plus_f & (dist == 0) & (g1 - f0 < after_feat),
            minus_g & minus_f & (dist == 0) & (f1 - g0 < after_feat),
            minus_g & plus_f & (0 < f0 - g1) & (f0 - g1 < before_feat),
            plus_g & minus_f & (0 < g0 - f1) & (g0 - f1 < before_feat),
            minus_g & plus_f & (0 < g1 - f0) & (g1 - f0 < after_feat),
            plus_g & minus_f & (0 < f1 - g0) & (f1 - g0 < after_feat),
        ]
        self.filtered_df = pd.concat([nb[m] for m in masks], axis=0)

    def _format_guide_table(self, targetprocessor_object) -> pd.DataFrame:
        """The output guide table (core.py:888-948), assembled from arrays instead of per-row ``.apply`` calls."""
        nb = targetprocessor_object.neighbors
        pretty_df = deepcopy(self.filtered_df)
        guides = pretty_df["Guide sequence"].astype(str).to_numpy()
        pos = neighbor_positions(nb, guides)
        keep = pos >= 0
        pretty_df = pretty_df[keep]
        guides, pos = guides[keep], pos[keep]
        mat = np.frombuffer("".join(guides).encode("ascii", "replace"), np.uint8).reshape(len(guides), -1) if len(guides) and \
            len(set(map(len, guides))) == 1 else None
        if mat is not None:
            pretty_df['GC'] = ((mat == ord("G")) | (mat == ord("C"))).sum(axis=1) / mat.shape[1]
        else:
            pretty_df['GC'] = [sum(c in "GC" for c in s) / len(s) for s in guides]
        uniq_g, inv = np.unique(guides, return_inverse=True) if len(guides) else (np.array([], dtype=str), np.array([], dtype=np.int64))
        names = np.array([hashlib.md5(s.encode()).hexdigest() for s in uniq_g], dtype=object)
        pretty_df['Guide name'] = names[inv] if len(guides) else []
        pretty_df['Target strand'] = np.where(pretty_df['Guide strand'].to_numpy() == pretty_df['Feature strand'].to_numpy(), 'coding', 'non-coding')
        dist_s, seq_s = similar_guide_strings(nb, pos)
        pretty_df['Similar guide distances'] = dist_s
        pretty_df['Similar guides'] = seq_s
        pretty_df = pd.merge(pretty_df, targetprocessor_object.targets, how="left",
                             left_on=['Guide sequence', 'Guide start', 'Guide end', 'Accession'],
                             right_on=['target', 'start', 'stop', 'seqid'])
        pretty_df = pretty_df.rename(columns={"exact_pam": "PAM"})
        pretty_df = pretty_df[['Guide name', 'Guide sequence', 'GC', 'dtype', 'Accession', 'Guide start', 'Guide end',
                               'Guide strand', 'PAM', 'Feature id',
                               'Feature start', 'Feature end', 'Feature strand',
                               'Feature distance', 'Similar guides', 'Similar guide distances', 'target_seq30']]
        pretty_df = pretty_df.merge(self.qualifiers, how="left", on="Feature id")
        pretty_df = pretty_df.sort_values(by=['Accession', 'Feature start'])
        pretty_df['Guide start'] = pretty_df['Guide start'] + 1          # 1-based, to match other tools (core.py:945-946)
        pretty_df['Feature start'] = pretty_df['Feature start'] + 1
        pretty_df = pretty_df.loc[pretty_df['target_seq30'].str.len() == 30]
        self.pretty_df = pretty_df

    def _filterlocus(self, attribute: str, filter_by_locus: list = []) -> pd.DataFrame:
        df = deepcopy(self.pretty_df)
        if len(filter_by_locus) > 0:
            df = df[df[attribute].isin(filter_by_locus)]
        return df

    def locuslen(self) -> int:
        da_keys = self.feature_dict.keys()
        firsttag = (list(da_keys)[0])
        if firsttag:
            return firsttag, len(self.feature_dict[firsttag].keys())
        logger.warning("A locus key could not be found.")
        return "notag", 0


# ---- bedtools closest, restated on sorted arrays -------------------------------------------------------------------------
def closest_features(g_start, g_end, f_start, f_end, upstream_frame: bool):
    """For every guide interval [g_start, g_end) on one contig: index of the reported feature [f_start, f_end) (features
    sorted by start; -1 if none is eligible) and the signed distance, for one of the reference's two frames.

    downstream frame (``-fd``): the feature with the smallest start >= guide end (equal starts: first in sorted order);
        distance = start - guide end + 1 (book-ended intervals are 1 apart).
    upstream frame (``-id``): the first feature in sorted order that overlaps the guide (distance 0), else the feature
        with the largest end <= guide start (equal ends: first in sorted order); distance = -(guide start - end + 1)."""
    g_start, g_end = np.asarray(g_start, np.int64), np.asarray(g_end, np.int64)
    f_start, f_end = np.asarray(f_start, np.int64), np.asarray(f_end, np.int64)
    n, m = len(g_start), len(f_start)
    if m == 0 or n == 0:
        return np.full(n, -1, np.int64), np.full(n, -1, np.int64)
    if not upstream_frame:
        ri = np.searchsorted(f_start, g_end, side="left")
        has = ri < m
        j = np.minimum(ri, m - 1)
        return np.where(has, j, -1), np.where(has, f_start[j] - g_end + 1, -1)
    # overlap: the first feature (in start order) whose end lies behind the guide start, provided it starts before
    # the guide ends.  pmax[i] > g_start says SOME feature j <= i ends behind g_start; the first such i is that feature.
    pmax = np.maximum.accumulate(f_end)
    first_open = np.searchsorted(pmax, g_start, side="right")
    n_start_lt = np.searchsorted(f_start, g_end, side="left")           # features with start < g_end
    overlap = first_open < n_start_lt
    e_order = np.lexsort((-np.arange(m), f_end))                         # by end; equal ends: later file position first
    e_sorted = f_end[e_order]
    li = np.searchsorted(e_sorted, g_start, side="right") - 1            # largest end <= g_start, first in file order
    has_l = li >= 0
    left = e_order[np.maximum(li, 0)]
    idx = np.where(overlap, np.minimum(first_open, m - 1), np.where(has_l, left, -1))
    dist = np.where(overlap, 0, np.where(has_l, -(g_start - f_end[left] + 1), -1))
    return idx, dist


# ---- neighbour columns of the guide table, from the neighbour arrays -------------------------------------------------------
def neighbor_positions(nb, guides: np.ndarray) -> np.ndarray:
    """row of every guide string in the neighbour map (-1 if absent); dict-like maps are served by membership tests"""
    if hasattr(nb, "codes") and hasattr(nb, "_lookup_index"):
        from ._encode import encode_guides
        out = np.full(len(guides), -1, np.int64)
        ok = np.array([len(s) == nb.L and set(s) <= set("ACGT") for s in guides], dtype=bool) if len(guides) else np.zeros(0, bool)
        if ok.any():
            codes = encode_guides(list(guides[ok]), nb.L)
            srt, pos = nb._lookup_index()
            j = np.searchsorted(srt, codes)
            jj = np.minimum(j, max(len(srt) - 1, 0))
            hit = (j < len(srt)) & (srt[jj] == codes) if len(srt) else np.zeros(len(codes), bool)
            out[np.flatnonzero(ok)[hit]] = pos[jj[hit]]
        return out
    keys = {k: i for i, k in enumerate(nb.keys())}
    return np.array([keys.get(s, -1) for s in guides], dtype=np.int64)


def similar_guide_strings(nb, pos: np.ndarray):
    """(';'-joined distances, ';'-joined neighbour sequences) for the map rows `pos` (core.py:914-921)"""
    if hasattr(nb, "index_matrix"):
        from ._encode import decode_guides
        idx, dist = nb.index_matrix()[pos], nb.distance_matrix()[pos]
        if len(pos) == 0:
            return [], []
        valid = idx >= 0
        if valid.all():                                                   # the usual case: k neighbours everywhere
            d = dist.astype(np.int64).astype(str)
            ds = d[:, 0]
            for j in range(1, d.shape[1]):
                ds = np.char.add(np.char.add(ds, ";"), d[:, j])
            u, inv = np.unique(idx, return_inverse=True)
            seqs = np.char.decode(decode_guides(nb.uniq[u], nb.L), "ascii")[inv.reshape(idx.shape)]
            ss = seqs[:, 0]
            for j in range(1, seqs.shape[1]):
                ss = np.char.add(np.char.add(ss, ";"), seqs[:, j])
            return ds.astype(object), ss.astype(object)
        uniq = np.char.decode(decode_guides(nb.uniq, nb.L), "ascii")
        return ([";".join(str(int(x)) for x in r[v]) for r, v in zip(dist, valid)],
                [";".join(uniq[r[v]]) for r, v in zip(idx, valid)])
    keys = list(nb.keys())
    return ([";".join(str(i) for i in nb[keys[p]]["neighbors"]["dist"]) for p in pos],
            [";".join(nb[keys[p]]["neighbors"]["seqs"]) for p in pos])
