"""Thin runner of the off-target hot path with the reference CLI's flags (guidemaker/cli.py:22-76).

It runs the reference's workflow (cli.py:161-189, :230-245) up to the point where the annotation join starts:
PAM scan -> restriction flag -> seed uniqueness -> exact kNN -> BED frame, and the random controls.  It writes

  rawguides.csv.gz   the reference's --raw_output_only table (header Chromosome,Start,Stop,gRNA,Strand; cli.py:191)
  offtargets.csv.gz  one row per kept guide locus: position, PAM, and the k nearest guides with their distances
  controls.csv.gz    the reference's control table (written with its index column, cli.py:239)

The feature-annotation join (bedtools), Doench / CFD scoring and plots are outside this repo's scope;
`targets.csv.gz` needs them and is therefore not produced here."""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import numpy as np
import pandas as pd
import yaml

from . import core
from .fastaio import get_records

DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_default.yaml")


def myparser():
    p = argparse.ArgumentParser(description="guidemaker_b200: GuideMaker's off-target hot path on a B200")
    p.add_argument('--genbank', '-i', nargs='+', type=str, required=False)
    p.add_argument('--fasta', '-f', nargs='+', type=str, required=False)
    p.add_argument('--gff', '-g', nargs='+', type=str, required=False, help="accepted for compatibility; annotation is out of scope")
    p.add_argument('--pamseq', '-p', type=str, required=True)
    p.add_argument('--outdir', '-o', type=str, required=True)
    p.add_argument('--raw_output_only', action='store_true')
    p.add_argument('--pam_orientation', '-r', choices=['5prime', '3prime'], default='3prime')
    p.add_argument('--guidelength', '-l', type=int, default=20, choices=range(10, 28, 1), metavar="[10-27]")
    p.add_argument('--lsr', type=int, default=10, choices=range(0, 28, 1), metavar="[0-27]")
    p.add_argument('--dtype', type=str, choices=['hamming', 'leven'], default='hamming')
    p.add_argument('--dist', type=int, choices=range(0, 6, 1), metavar="[0-5]", default=2)
    p.add_argument('--knum', type=int, default=5, choices=range(2, 21, 1), metavar="[2-20]")
    p.add_argument('--controls', type=int, default=1000, choices=range(0, 100001, 1), metavar="[0-100000]")
    p.add_argument('--threads', type=int, default=2, help="accepted and ignored (the search runs on the GPU)")
    p.add_argument('--log', default="guidemaker.log")
    p.add_argument('--restriction_enzyme_list', nargs="*", default=[])
    p.add_argument('--config', default=DEFAULT_CONFIG)
    return p


def parserval(args):
    """cli.py:80-89"""
    assert args.lsr <= args.guidelength, "The length of sequence near the PAM .i.e seed sequence that must be less than the guide length"
    assert 1 < len(args.pamseq) < 9, "The length of the PAM sequence must be between 2-8"
    assert args.genbank is not None or args.fasta is not None, "Please provide either Genbank files or Fasta files."


def main(arglist: list = None):
    args = myparser().parse_args(arglist)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s",
                        handlers=[logging.StreamHandler(), logging.FileHandler(args.log)])
    logger = logging.getLogger("guidemaker_b200")
    parserval(args)
    try:
        with open(args.config) as cf:
            config = yaml.safe_load(cf)
        logger.info("Configuration data loaded from %s: %s", args.config, config)
        t0 = time.perf_counter()
        records = get_records(args.genbank, "genbank") if args.genbank else get_records(args.fasta, "fasta")
        logger.info("Identifying PAM sites in the genome")
        pamobj = core.PamTarget(args.pamseq, args.pam_orientation, args.dtype)
        pamtargets = pamobj.find_targets(seq_record_iter=records, target_len=args.guidelength)
        tl = core.TargetProcessor(targets=pamtargets, lsr=args.lsr, editdist=args.dist, knum=args.knum)
        lengthoftl = len(tl.targets)
        logger.info("Checking guides for restriction enzymes")
        tl.check_restriction_enzymes(restriction_enzyme_list=args.restriction_enzyme_list)
        logger.info("Identifing guides that are unique near the PAM site")
        tl.find_unique_near_pam()
        logger.info("Number of guides with non unique seed sequence: %d", tl.targets.isseedduplicated.sum())
        tl.create_index(num_threads=args.threads, configpath=args.config)
        logger.info("Indexing all potential guide sites: %s.", len(tl.nmslib_index))
        logger.info("Identifying guides that have a hamming distance <= %s to all other potential guides", str(args.dist))
        tl.get_neighbors(num_threads=args.threads, configpath=args.config)
        tf_df = tl.export_bed()
        os.makedirs(args.outdir, exist_ok=True)
        tf_df.to_csv(os.path.join(args.outdir, "rawguides.csv.gz"), index=False, header=["Chromosome", "Start", "Stop", "gRNA", "Strand"])
        if not args.raw_output_only:
            offtarget_table(tl).to_csv(os.path.join(args.outdir, "offtargets.csv.gz"), index=False)
            if args.controls > 0:
                logger.info("Creating random control guides")
                cmin, cmed, randomdf = tl.get_control_seqs(records, configpath=args.config, length=args.guidelength,
                                                           n=args.controls, num_threads=args.threads)
                randomdf.to_csv(os.path.join(args.outdir, "controls.csv.gz"))
                logger.info("Number of random control searched: %d", tl.ncontrolsearched)
                logger.info("Created %i control guides with a minimum distance of %d and a median distance of %d", args.controls, cmin, cmed)
                logger.info("Percentage of GC content in the input genome: %.2f", tl.gc_percent)
                logger.info("Total length of the genome: %.1f MB", tl.genomesize)
        logger.info("guidemaker_b200 completed in %.2f s, results are at %s", time.perf_counter() - t0, args.outdir)
        logger.info("PAM sequence: %s", args.pamseq)
        logger.info("PAM orientation: %s", args.pam_orientation)
        logger.info("Genome strand(s) searched: %s", "both")
        logger.info("Total PAM sites considered: %d", lengthoftl)
        logger.info("Guide RNA candidates found: %d", len(tl.neighbors))
    except Exception:
        logger.exception("guidemaker_b200 terminated with errors. See the log file for details.")
        raise SystemExit(1)


def offtarget_table(tl: "core.TargetProcessor") -> pd.DataFrame:
    """One row per locus whose guide survived both filters (first-seen seed, nearest other guide >= dist), with the
    columns of the reference's final table that come from the hot path (core.py:917-942): guide, position, strand,
    PAM, `Similar guides` and `Similar guide distances` (';'-joined).  Vectorised over the neighbour arrays."""
    nb = tl.neighbors
    t = tl.targets.loc[tl.targets['isseedduplicated'] == False]  # noqa: E712
    keys = pd.Index(np.char.decode(nb.key_array(), "ascii"))
    pos = keys.get_indexer(t['target'])
    t = t.loc[pos >= 0]
    pos = pos[pos >= 0]
    idx, dist = nb.index_matrix()[pos], nb.distance_matrix()[pos]
    from ._encode import decode_guides
    uniq = np.char.decode(decode_guides(nb.uniq, nb.L), "ascii")
    valid = idx >= 0
    seqs = np.where(valid, uniq[np.where(valid, idx, 0)], "")
    sim = [";".join(r[v]) for r, v in zip(seqs, valid)]
    dd = [";".join(str(int(x)) for x in r[v]) for r, v in zip(dist, valid)]
    return pd.DataFrame({"Guide sequence": t['target'].to_numpy(), "Accession": t['seqid'].astype(str).to_numpy(),
                         "Guide start": t['start'].to_numpy() + 1, "Guide end": t['stop'].to_numpy(),
                         "Guide strand": np.where(t['strand'].to_numpy(dtype=bool), '+', '-'), "PAM": t['exact_pam'].astype(str).to_numpy(),
                         "Similar guides": sim, "Similar guide distances": dd, "target_seq30": t['target_seq30'].to_numpy()})


if __name__ == "__main__":
    main(sys.argv[1:])
