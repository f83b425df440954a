"""Write a synthetic database as sourmash signature files on all host cores (test infrastructure)."""
from oracle.train_oracle import write_sig_dir_parallel


def write_sig_files(db, root, threads=None):
    """root/signatures/g<id>.sig for every sketch + root/training_sig_files.tsv; returns the paths (genome id order)."""
    return write_sig_dir_parallel(db, root, threads)
