"""`python -m yacht_b200 train|run|sketch ...` -- the sub-commands of the reference CLI that sit on the
hot path (reference src/yacht/__init__.py:54-136) plus `sketch ref|sample` (:106-121); the
download and convert sub-commands are out of scope."""
import argparse
import sys

from . import make_training_data_from_sketches, run_YACHT, sketch_ref_genomes, sketch_sample
from .utils import __version__


def main(argv=None):
    parser = argparse.ArgumentParser(prog="yacht", description="YACHT hot path on B200")
    parser.add_argument("--version", action="version", version=f"yacht_b200 {__version__}")
    sub = parser.add_subparsers(dest="command")
    p_train = sub.add_parser("train", description="Pre-process the reference genomes")
    make_training_data_from_sketches.add_arguments(p_train)
    p_train.set_defaults(func=make_training_data_from_sketches.main)
    p_run = sub.add_parser("run", description="Run the YACHT algorithm")
    run_YACHT.add_arguments(p_run)
    p_run.set_defaults(func=run_YACHT.main)
    p_sketch = sub.add_parser("sketch", description="Sketch reference genomes or metagenomics samples")
    sk_sub = p_sketch.add_subparsers(dest="sketch_subcommand")
    p_ref = sk_sub.add_parser("ref", description="Sketch fasta files and make them as references")
    sketch_ref_genomes.add_arguments(p_ref)
    p_ref.set_defaults(func=sketch_ref_genomes.main)
    p_sample = sk_sub.add_parser("sample", description="Sketch metagenomics samples")
    sketch_sample.add_arguments(p_sample)
    p_sample.set_defaults(func=sketch_sample.main)
    args = parser.parse_args(argv)
    if "func" not in args:
        parser.print_help(file=sys.stderr)
        return 1
    args.func(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
