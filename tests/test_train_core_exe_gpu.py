"""The drop-in executable yacht_b200/run_yacht_train_core against the outputs of the UNMODIFIED
reference core (tests/golden/train_golden.json): same pair-file lines, same selected_result.tsv
(byte-for-byte, order included), same banner statistics, same file partition."""
import glob
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import train_oracle as to
from test_oracle_pinned import golden_case_db, _load_train_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "yacht_b200", "run_yacht_train_core")


def run_exe(workdir, thr, t, p, extra_env=None):
    sel = os.path.join(workdir, "selected_result.tsv")
    cmd = [EXE, "-t", str(t), "-c", repr(float(thr)), "-p", str(p), os.path.join(workdir, "training_sig_files.tsv"), workdir, sel]
    env = dict(os.environ)
    env.update(extra_env or {})
    cp = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    return cp, sel


@pytest.mark.parametrize("name", sorted(_load_train_golden().keys()))
def test_exe_matches_reference_golden(name, tmp_path):
    case = _load_train_golden()[name]
    db = golden_case_db(name, case)
    wd = str(tmp_path)
    paths = to.write_sig_dir(db.hashes, db.offsets, wd)
    t, p = 3, 2
    cp, sel = run_exe(wd, case["thr"], t, p)
    assert cp.returncode == 0, cp.stderr
    got = to.parse_core_outputs(wd, paths, sel, cp.stdout)
    assert (got.n_distinct, got.n_singleton, got.n_index) == (case["n_distinct"], case["n_singleton"], case["n_index"])
    assert got.lines == case["lines"]
    assert [int(x) for x in got.selected] == case["selected"]
    # P x T pair files exist (reference main.cpp:265-271), possibly empty
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(wd, "*_*.txt")))
    assert files == sorted(f"{pp}_{tt:03d}.txt" for pp in range(p) for tt in range(t))
    # selected file = the paths exactly as they appear in the file list
    with open(sel) as f:
        assert [l.rstrip("\n") for l in f] == [paths[g] for g in case["selected"]]


@pytest.mark.skipif(not to.reference_available(), reason="oracle/_ref not present")
def test_exe_file_partition_matches_reference_binary(tmp_path):
    from yacht_b200 import synth
    db = synth.make_reference_db(257, 77, mean_size=200, sd_size=60)
    thr = 0.3
    ref_dir, my_dir = str(tmp_path / "ref"), str(tmp_path / "mine")
    os.makedirs(ref_dir); os.makedirs(my_dir)
    paths_r = to.write_sig_dir(db.hashes, db.offsets, ref_dir)
    paths_m = to.write_sig_dir(db.hashes, db.offsets, my_dir)
    to.run_core_binary(to.REF_BIN, os.path.join(ref_dir, "training_sig_files.tsv"), ref_dir, thr, threads=4, passes=3)
    cp, _ = run_exe(my_dir, thr, 4, 3)
    assert cp.returncode == 0, cp.stderr
    for pp in range(3):
        for tt in range(4):
            fn = f"{pp}_{tt:03d}.txt"
            with open(os.path.join(ref_dir, fn)) as a, open(os.path.join(my_dir, fn)) as b:
                assert a.read() == b.read(), fn   # byte-identical per file
    with open(os.path.join(ref_dir, "selected_result.tsv")) as a, open(os.path.join(my_dir, "selected_result.tsv")) as b:
        assert [os.path.basename(x) for x in a.read().split()] == [os.path.basename(x) for x in b.read().split()]


def test_exe_missing_and_malformed_files(tmp_path):
    from yacht_b200 import synth
    db = synth.from_sketches([np.arange(1, 30, dtype=np.uint64), np.arange(1, 30, dtype=np.uint64), np.arange(5, 40, dtype=np.uint64)])
    wd = str(tmp_path)
    paths = to.write_sig_dir(db.hashes, db.offsets, wd)
    os.remove(paths[1])   # unreadable file => "Could not open the file!" and an EMPTY sketch (main.cpp:68-71)
    cp, sel = run_exe(wd, 0.5, 1, 1)
    assert cp.returncode == 0 and "Could not open the file!" in cp.stderr
    assert "Number of empty sketches: 1" in cp.stdout and "Empty sketch ids: 1" in cp.stdout
    got = to.parse_core_outputs(wd, paths, sel, cp.stdout)
    exp = to.oracle_train(*(lambda d: (d.hashes, d.offsets))(synth.from_sketches([db.sketch(0), np.zeros(0, np.uint64), db.sketch(2)])), 0.5)
    assert got.lines == exp.lines and list(got.selected) == list(exp.selected)
    with open(paths[2], "w") as f:
        f.write("{not json")
    cp, _ = run_exe(wd, 0.5, 1, 1)
    assert cp.returncode != 0 and "cannot parse signature" in cp.stderr


def test_exe_real_signature_layout(tmp_path):
    # a file laid out like the reference's fixture: two sub-signatures, extra keys, nested arrays,
    # escaped quotes in the name -- only [0]["signatures"][0]["mins"] counts (main.cpp:78)
    wd = str(tmp_path)
    os.makedirs(os.path.join(wd, "signatures"))
    docs = [
        [{"class": "sourmash_signature", "email": "", "hash_function": "0.murmur64", "filename": "a \"quoted\" [name].fa",
          "name": "g0 {x} \\", "license": "CC0",
          "signatures": [{"num": 0, "ksize": 31, "seed": 42, "max_hash": 18446744073709552, "mins": [5, 9, 18446744073709551615],
                          "md5sum": "x", "abundances": [1, 2, 3], "molecule": "dna"},
                         {"num": 0, "ksize": 51, "mins": [1, 2, 3, 4]}], "version": 0.4},
         {"signatures": [{"mins": [77]}]}],
        [{"name": "g1", "signatures": [{"ksize": 21, "mins": [9, 5, 100], "extra": {"mins": [1]}}]}],
    ]
    paths = []
    for k, d in enumerate(docs):
        p = os.path.join(wd, "signatures", f"s{k}.sig")
        with open(p, "w") as f:
            json.dump(d, f, indent=1 if k else None)
        paths.append(p)
    with open(os.path.join(wd, "training_sig_files.tsv"), "w") as f:
        f.write("\n".join(paths) + "\n")
    cp, sel = run_exe(wd, 0.5, 1, 1)
    assert cp.returncode == 0, cp.stderr
    got = to.parse_core_outputs(wd, paths, sel, cp.stdout)
    assert got.lines == ["0,1,0.5,0.666667,0.666667", "1,0,0.5,0.666667,0.666667"]
    assert (got.n_distinct, got.n_singleton, got.n_index) == (4, 2, 2)


def _n_gpus():
    from yacht_b200 import _lib
    return _lib.load_library().ygpu_device_count()


@pytest.mark.parametrize("ngpu", [2])      # larger rank counts of the same library entry points: tests/test_sharded_step_gpu.py
def test_exe_multi_gpu_output_equals_single_gpu(ngpu, tmp_path):
    """YACHT_NUM_GPUS=k: threads as ranks, hash-range residency, sharded step (the replicated step when the database does not
    qualify) -- the files must be byte-identical to the single-GPU run's."""
    if _n_gpus() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    from yacht_b200 import synth
    outs = {}
    for label, seed_db in (("clustered", synth.make_reference_db(3000, 5, mean_size=1500, sd_size=300)),
                           ("skewed", synth.make_skewed_db(1500, seed=9, mean_size=600, sd_size=100, zipf_cap=900, core_hashes=40))):
        for k in (1, ngpu):
            wd = str(tmp_path / f"{label}_{k}")
            os.makedirs(wd)
            to.write_sig_dir(seed_db.hashes, seed_db.offsets, wd)
            cp, sel = run_exe(wd, 0.95 ** 31, 4, 2, {"YACHT_NUM_GPUS": str(k)})
            assert cp.returncode == 0, cp.stderr
            files = {}
            for fn in sorted(glob.glob(os.path.join(wd, "*_*.txt"))):
                with open(fn) as f:
                    files[os.path.basename(fn)] = f.read()
            with open(sel) as f:
                files["selected"] = [os.path.basename(x) for x in f.read().split()]
            outs[(label, k)] = files
        assert outs[(label, 1)] == outs[(label, ngpu)], label
        assert sum(len(v) for kf, v in outs[(label, 1)].items() if kf != "selected") > 0
