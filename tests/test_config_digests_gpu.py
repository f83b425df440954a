"""The CUDA path (through the C ABI) and the drop-in executable at BASELINE.json's own sizes, against digests of what
the UNMODIFIED reference core wrote for the same seeded databases (tests/golden/config_digests.json, produced by
tests/golden/make_config_digests.py with oracle/_ref): banner statistics, every pair line, the retained set."""
import glob
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import train_oracle as to
from yacht_b200 import _lib, pairfmt, synth
from test_config_digests import as_pairs, check_against_digest, load_digests
from harness.sigfiles import write_sig_files

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "yacht_b200", "run_yacht_train_core")


def _abi_result(ctx, db, thr):
    ctx.load_sketches(db.hashes, db.offsets)
    st = ctx.build_index()
    pairs = ctx.pairwise_flag(thr)
    sel = _lib.greedy_select(db.offsets, pairs)
    return st, pairs, sel


@pytest.mark.parametrize("name", ["config2", "config3", "config4s"])
def test_c_abi_matches_reference_digest(gpu_ctx, name, config_db):
    entry = load_digests()[name]
    db = config_db(name)
    st, pairs, sel = _abi_result(gpu_ctx, db, entry["threshold"])
    assert st["index_path"] == 1
    check_against_digest(entry, db.sizes, (st["n_distinct"], st["n_singleton"], st["n_index"]), pairs, sel)
    if name == "config4s":
        # 40 M flagged pairs: the text digest is skipped above, so the counts are held against the oracle port (itself
        # held against the same reference digest in tests/test_config_digests.py), and both grouping kernels must agree
        ref = to.oracle_train(db.hashes, db.offsets, entry["threshold"])
        assert np.array_equal(pairs["count"], ref.pairs["count"]) and np.array_equal(pairs["i"], ref.pairs["i"])
        assert st["big_buckets"] > 0            # the skewed shape really leaves the shared-memory grouping kernel
        gpu_ctx.set_option("group_kernel", 1)
        try:
            st1, pairs1, _ = _abi_result(gpu_ctx, db, entry["threshold"])
        finally:
            gpu_ctx.set_option("group_kernel", 0)
        assert pairs1.tobytes() == pairs.tobytes() and st1["n_increments"] == st["n_increments"]


@pytest.mark.parametrize("name", ["config2", "config3"])
def test_executable_matches_reference_digest(name, config_db):
    entry = load_digests()[name]
    db = config_db(name)
    root = tempfile.mkdtemp(prefix=f"yacht_exe_{name}_")
    try:
        paths = write_sig_files(db, root)
        wd = os.path.join(root, "wd")
        os.makedirs(wd)
        sel = os.path.join(wd, "selected_result.tsv")
        cp = subprocess.run([EXE, "-t", str(os.cpu_count() or 8), "-c", repr(float(entry["threshold"])), "-p", "2",
                             os.path.join(root, "training_sig_files.tsv"), wd, sel], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert cp.returncode == 0, cp.stderr[-2000:]
        got = to.parse_core_outputs(wd, paths, sel, cp.stdout)
        b = entry["banners"]
        assert (got.n_distinct, got.n_singleton, got.n_index) == (b["n_distinct"], b["n_singleton"], b["n_index"])
        assert len(got.lines) == entry["F"]
        assert pairfmt.digest_lines(got.lines) == entry["pairs_sha256"]          # the files' own text, floats included
        assert pairfmt.digest_ids(got.selected) == entry["selected_sha256"]
    finally:
        shutil.rmtree(root, ignore_errors=True)
