#!/usr/bin/env python
"""``yacht sketch ref`` -- mirror of the reference's src/yacht/sketch_ref_genomes.py (same arguments, same output kinds),
with the hashing on the GPU instead of a ``sourmash sketch`` subprocess.

  --infile FILE   -> one sketch per sequence record (the reference passes ``--singleton``, :25)
  --infile FOLDER -> one sketch per sequence file found below it, named after the file without its extension
                     (the reference writes dataset.csv and calls ``sourmash sketch fromfile``, :31-67)
Sketches carry abundances (``abund``) and are written to ``--outfile`` (``.zip`` -> sourmash zip database).
"""
import argparse
import os
from pathlib import Path

from . import sketch
from .utils import _log

FILE_EXTENSIONS = ["*.fasta", "*.fna", "*.fas", "*.fa", "*.fasta.gz", "*.fna.gz", "*.fas.gz", "*.fa.gz"]   # reference :32-41


def add_arguments(parser):
    parser.add_argument("--infile", help="Input file or folder path.", required=True)
    parser.add_argument("--kmer", type=int, help="K-mer size.", default=31)
    parser.add_argument("--scaled", type=int, help="Scaled factor.", default=1000)
    parser.add_argument("--outfile", help="Output file name.", required=True)


def sketch_single_file(infile, kmer, scaled, outfile):
    _log("INFO", f"Starting sketching a single file: {infile}")
    sketches = sketch.sketch_files([infile], kmer, scaled, singleton=True)
    sketch.write_sketches(outfile, sketches, kmer, scaled)
    _log("SUCCESS", "Successfully sketched!!")


def dataset_rows(folder_path):
    """(name, absolute path) per sequence file, in the order the reference writes dataset.csv (:47-58)."""
    rows = []
    for extension in FILE_EXTENSIONS:
        for path in Path(folder_path).glob(f"**/{extension}"):
            rows.append((path.name.replace(extension.replace("*", ""), ""), str(path.absolute())))
    return rows


def sketch_multiple_files(folder_path, kmer, scaled, outfile):
    dataset_file = os.path.join(folder_path, "dataset.csv")
    _log("INFO", f"Preparing dataset file for multiple sequence files in {folder_path}")
    rows = dataset_rows(folder_path)
    with open(dataset_file, "w") as f:
        f.write("name,genome_filename,protein_filename\n")
        for name, path in rows:
            f.write(f"{name},{path},\n")
    _log("INFO", f"Starting sketching multiple sequence files in: {folder_path}")
    sketches = sketch.sketch_files([p for _, p in rows], kmer, scaled, singleton=False, names=[n for n, _ in rows])
    sketch.write_sketches(outfile, sketches, kmer, scaled)
    _log("SUCCESS", f"Successfully sketched files in: {folder_path}")


def main(args):
    try:
        if os.path.isfile(args.infile):
            sketch_single_file(args.infile, args.kmer, args.scaled, args.outfile)
        elif os.path.isdir(args.infile):
            sketch_multiple_files(args.infile, args.kmer, args.scaled, args.outfile)
        else:
            raise FileNotFoundError(f"Input path {args.infile} does not exist.")
    except FileNotFoundError as e:
        _log("ERROR", str(e))


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Sketch genomes on the GPU.", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_arguments(parser)
    main(parser.parse_args())
