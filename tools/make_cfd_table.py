"""Writes guidemaker_b200/data/cfd_mm_scores.json: the CFD mismatch weights of Doench et al. 2016 (Nat. Biotechnol. 34:184,
Supplementary Table 19) as a dense [rna base A,C,G,U][dna base A,C,G,T][position 1..20] table, taken from the copy the
reference ships (guidemaker/data/cfd_data.json, key "mm"; keys 'r<X>:d<Y>,<pos>').  Combinations that are not
mismatches (rA:dT, rC:dG, rG:dC, rU:dA) never occur in a product and are stored as 1.0.

    python tools/make_cfd_table.py [/root/reference]"""
import json
import os
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = json.load(open(os.path.join(ref, "guidemaker", "data", "cfd_data.json")))["mm"]
table = [[[1.0] * 20 for _ in "ACGT"] for _ in "ACGU"]
for key, val in src.items():
    r, rest = key[1], key[4:]
    d, pos = rest.split(",")
    table["ACGU".index(r)]["ACGT".index(d)][int(pos) - 1] = float(val)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "guidemaker_b200", "data", "cfd_mm_scores.json")
json.dump({"source": "Doench et al. 2016 CFD mismatch weights, via the reference's guidemaker/data/cfd_data.json ('mm')",
           "axes": ["rna base of the guide: A,C,G,U", "dna base = complement of the off-target base: A,C,G,T", "position 1..20 (PAM-distal -> PAM-proximal)"],
           "mm": table}, open(out, "w"), indent=0)
print(out, len(src), "weights")
