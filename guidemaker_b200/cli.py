"""Runner of the off-target hot path and the annotation join with the reference CLI's flags (guidemaker/cli.py:22-76).

It follows the reference's workflow (cli.py:161-245): PAM scan -> restriction flag -> seed uniqueness -> exact kNN -> BED
frame -> nearest-feature join -> filters -> guide table (-> CFD scores) -> random controls, and writes

  rawguides.csv.gz   the reference's --raw_output_only table (header Chromosome,Start,Stop,gRNA,Strand; cli.py:191)
  targets.csv.gz     the guide table (cli.py:226-227) when annotation is available (--genbank or --gff)
  offtargets.csv.gz  one row per kept guide locus with its k nearest guides (always written; not in the reference)
  controls.csv.gz    the reference's control table (written with its index column, cli.py:239)

Doench efficiency scoring (an ONNX model) and the plots are outside this repo's scope."""
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
    p.add_argument('--gff', '-g', nargs='+', type=str, required=False, help="GFF/GTF annotation (with --fasta)")
    p.add_argument('--pamseq', '-p', type=str, required=True)
    p.add_argument('--outdir', '-o', type=str, required=True)
    p.add_argument('--raw_output_only', action='store_true')
    p.add_argument('--pam_orientation', '-r', choices=['5prime', '3prime'], default='3prime')
    p.add_argument('--guidelength', '-l', type=int, default=20, choices=range(10, 28, 1), metavar="[10-27]")
    p.add_argument('--lsr', type=int, default=10, choices=range(0, 28, 1), metavar="[0-27]")
    p.add_argument('--dtype', type=str, choices=['hamming', 'leven'], default='hamming')
    p.add_argument('--dist', type=int, choices=range(0, 6, 1), metavar="[0-5]", default=2)
    p.add_argument('--before', type=int, default=100, choices=range(1, 501, 1), metavar="[1-500]")
    p.add_argument('--into', type=int, default=200, choices=range(1, 501, 1), metavar="[1-500]")
    p.add_argument('--attribute_key', type=str, default=None)
    p.add_argument('--filter_by_attribute', nargs="*", default=[])
    p.add_argument('--cfd_score', action='store_true')
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
        n_guides = len(tl.neighbors)
        if not args.raw_output_only:
            offtarget_table(tl).to_csv(os.path.join(args.outdir, "offtargets.csv.gz"), index=False)
            if args.genbank or args.gff:
                logger.info("Create GuideMaker Annotation object")
                anno = core.Annotation(annotation_list=args.genbank or args.gff, annotation_type="genbank" if args.genbank else "gff",
                                       target_bed_df=tf_df)
                logger.info("Identify genomic features")
                anno.get_annotation_features()
                logger.info("Total number of %s in the input genome: %d" % anno.locuslen())
                logger.info("Find genomic features closest the guides")
                anno._get_nearby_features()
                logger.info("Select guides that start between +%s and -%s of a feature start" % (args.before, args.into))
                anno._filter_features(before_feat=args.before, after_feat=args.into)
                logger.info("Select description columns")
                anno._get_qualifiers(configpath=args.config)
                logger.info("Format the output")
                anno._format_guide_table(tl)
                prettydf = anno._filterlocus(args.attribute_key, args.filter_by_attribute)
                if args.cfd_score:
                    logger.info("Calculating CFD score for assessing off-target activity of gRNAs")
                    prettydf = core.cfd_score(df=prettydf)
                logger.info("Number of Guides within a gene coordinates i.e. zero Feature distance: %d", prettydf['Feature distance'].isin([0]).sum())
                prettydf.to_csv(os.path.join(args.outdir, "targets.csv.gz"), index=False)
                n_guides = len(prettydf)
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
        logger.info("Guide RNA candidates found: %d", n_guides)
    except Exception:
        logger.exception("guidemaker_b200 terminated with errors. See the log file for details.")
        raise SystemExit(1)


def offtarget_table(tl: "core.TargetProcessor") -> pd.DataFrame:
    """One row per locus whose guide survived both filters (first-seen seed, nearest other guide >= dist), with the
    columns of the reference's final table that come from the hot path (core.py:917-942): guide, position, strand,
    PAM, `Similar guides` and `Similar guide distances` (';'-joined).  Assembled from the neighbour arrays, no per-row Python."""
    from .annotation import neighbor_positions, similar_guide_strings
    nb = tl.neighbors
    t = tl.targets.loc[tl.targets['isseedduplicated'] == False]  # noqa: E712
    pos = neighbor_positions(nb, t['target'].astype(str).to_numpy())
    t = t.loc[pos >= 0]
    pos = pos[pos >= 0]
    dd, sim = similar_guide_strings(nb, pos)
    return pd.DataFrame({"Guide sequence": t['target'].to_numpy(), "Accession": t['seqid'].astype(str).to_numpy(),
                         "Guide start": t['start'].to_numpy() + 1, "Guide end": t['stop'].to_numpy(),
                         "Guide strand": np.where(t['strand'].to_numpy(dtype=bool), '+', '-'), "PAM": t['exact_pam'].astype(str).to_numpy(),
                         "Similar guides": sim, "Similar guide distances": dd, "target_seq30": t['target_seq30'].to_numpy()})


if __name__ == "__main__":
    main(sys.argv[1:])
