#!/usr/bin/env python
"""``yacht sketch sample`` -- mirror of the reference's src/yacht/sketch_sample.py (same arguments): one sketch with
abundances of all reads of one file (single-end, :31-39) or of two files together (paired-end: the reference concatenates
them into a temporary file first, :42-57)."""
import argparse

from . import sketch
from .utils import _log


def add_arguments(parser):
    parser.add_argument("--infile", nargs="+", help="Input FASTA/Q file(s). For paired-end reads, provide two files.", required=True)
    parser.add_argument("--kmer", type=int, help="K-mer size.", default=31)
    parser.add_argument("--scaled", type=int, help="Scaled factor.", default=1000)
    parser.add_argument("--outfile", help="Output file name.", required=True)


def _sketch_together(infiles, kmer, scaled, outfile, filename):
    groups = [[seq for path in infiles for _, seq in sketch.read_records(path)]]
    (mins, ab), = sketch.sketch_record_groups(groups, kmer, scaled)
    sketch.write_sketches(outfile, [{"name": "", "filename": filename, "mins": mins, "abundances": ab}], kmer, scaled)


def sketch_single_end(infile, kmer, scaled, outfile):
    _log("INFO", f"Starting sketching a single-end FASTA/Q file: {infile}")
    _sketch_together([infile], kmer, scaled, outfile, infile)
    _log("SUCCESS", "Successfully sketched!!")


def sketch_paired_end(infile1, infile2, kmer, scaled, outfile):
    _log("INFO", f"Starting sketching paired-end FASTA/Q files: {infile1} {infile2}")
    _sketch_together([infile1, infile2], kmer, scaled, outfile, infile1)
    _log("SUCCESS", "Successfully sketched!!")


def main(args):
    if len(args.infile) == 1:
        sketch_single_end(args.infile[0], args.kmer, args.scaled, args.outfile)
    elif len(args.infile) == 2:
        sketch_paired_end(args.infile[0], args.infile[1], args.kmer, args.scaled, args.outfile)
    else:
        raise ValueError("Please provide either one file for single-end reads or two files for paired-end reads.")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Sketch single-end or paired-end reads on the GPU.", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_arguments(parser)
    main(parser.parse_args())
