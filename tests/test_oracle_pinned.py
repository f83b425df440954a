"""The CPU oracles are pinned to the reference (CPU only, no GPU needed).

* train: oracle/train_oracle.cpp against the golden outputs of the UNMODIFIED reference core
  (tests/golden/train_golden.json, produced by tests/golden/make_train_golden.py from
  oracle/_ref/run_yacht_train_core_ref), and -- when that binary is present -- live against it.
* run: oracle/run_oracle.py against every evaluation stored in the reference's checked-in result
  workbooks (tests/golden/run_golden.json.gz), the known-answer tests of the reference's own
  tests/test_unit.py:11-20 and tests/test_unittests.py:86-111, and the end-to-end known answer of
  tests/test_workflow.py:58-66 on the reference's 20-genome fixture.
"""
import gzip
import json
import math
import os
import tempfile

import numpy as np
import pytest

from oracle import run_oracle as ro
from oracle import train_oracle as to
from yacht_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load_train_golden():
    with open(os.path.join(GOLD, "train_golden.json")) as f:
        return json.load(f)


def _fixture20():
    z = np.load(os.path.join(GOLD, "fixture20.npz"))
    return synth.SketchDB(hashes=z["hashes"], offsets=z["offsets"], cluster=np.full(len(z["offsets"]) - 1, -1)), z


def _edge():
    parts = [np.arange(1, 11), np.arange(1, 11), np.zeros(0), np.array([1, 2, 3, 4, 5] + list(range(100, 107))),
             np.array([7, 7, 7, 200])]
    return synth.from_sketches([np.asarray(p, dtype=np.uint64) for p in parts])


def _ties():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mtg", os.path.join(GOLD, "make_train_golden.py"))
    # the generator of the tie case lives in the golden script; re-create it without running main()
    rng = np.random.default_rng(123)
    parts = []
    for grp in range(12):
        base = rng.integers(0, synth.MAX_HASH, size=40, dtype=np.uint64)
        for m in range(int(rng.integers(2, 5))):
            x = base.copy()
            x[m] = rng.integers(0, synth.MAX_HASH, dtype=np.uint64)
            parts.append(np.unique(x))
    for k in range(30):
        parts.append(np.unique(rng.integers(0, synth.MAX_HASH, size=int(rng.integers(20, 60)), dtype=np.uint64)))
    order = rng.permutation(len(parts))
    return synth.from_sketches([parts[i] for i in order])


def golden_case_db(name, case):
    if name.startswith("fixture20"):
        return _fixture20()[0]
    if name.startswith("edge"):
        return _edge()
    if name == "ties":
        return _ties()
    g = case["gen"]
    db = synth.make_reference_db(g["n"], g["seed"], mean_size=g["mean_size"], sd_size=g["sd_size"])
    assert int(db.offsets[-1]) == case["T"] and int(np.bitwise_xor.reduce(db.hashes)) == case["checksum"], \
        "synthetic generator no longer reproduces the golden inputs"
    return db


@pytest.mark.parametrize("name", sorted(_load_train_golden().keys()))
def test_train_port_matches_reference_golden(name):
    case = _load_train_golden()[name]
    db = golden_case_db(name, case)
    r = to.oracle_train(db.hashes, db.offsets, case["thr"])
    assert (r.n_distinct, r.n_singleton, r.n_index) == (case["n_distinct"], case["n_singleton"], case["n_index"])
    assert r.lines == case["lines"]
    assert [int(x) for x in r.selected] == case["selected"]


def test_fixture20_survey_facts():
    # 63 888 distinct, 63 879 singletons, index 9, no pair, all 20 selected
    case = _load_train_golden()["fixture20"]
    assert (case["n_distinct"], case["n_singleton"], case["n_index"]) == (63888, 63879, 9)
    assert case["lines"] == [] and len(case["selected"]) == 20


@pytest.mark.skipif(not to.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,seed,thr,t,p", [(200, 31, 0.95 ** 31, 1, 1), (350, 32, 0.3, 4, 3), (64, 33, 0.0, 7, 2)])
def test_train_port_matches_reference_live(n, seed, thr, t, p):
    db = synth.make_reference_db(n, seed, mean_size=300, sd_size=120)
    with tempfile.TemporaryDirectory() as d:
        ref = to.reference_train(db.hashes, db.offsets, thr, d, threads=t, passes=p)
    r = to.oracle_train(db.hashes, db.offsets, thr)
    assert r.lines == ref.lines
    assert list(r.selected) == list(ref.selected)
    assert (r.n_distinct, r.n_singleton, r.n_index) == (ref.n_distinct, ref.n_singleton, ref.n_index)


# ---- run path ---------------------------------------------------------------------------------
def test_alt_mut_rate_reference_kats():
    # reference tests/test_unit.py:11-20 ("calculated with Steve's code"), np.isclose defaults
    assert ro.get_alt_mut_rate(100, 10000, 21, significance=0.99) == -1
    assert np.isclose(ro.get_alt_mut_rate(10, 0, 21), 0.28015945851802826)
    assert np.isclose(ro.get_alt_mut_rate(10, 0, 31), 0.19963312102481723)
    assert np.isclose(ro.get_alt_mut_rate(10, 5, 21), 0.0698992155957967)
    assert np.isclose(ro.get_alt_mut_rate(10, 5, 31), 0.047902071848511696)
    assert np.isclose(ro.get_alt_mut_rate(10, 9, 21), 0.02169068099465221)
    assert np.isclose(ro.get_alt_mut_rate(100, 10, 11), 0.2397729973308742)
    assert np.isclose(ro.get_alt_mut_rate(1000, 0, 1), 0.9999899497147453)
    # reference tests/test_unittests.py:86-111
    assert math.isclose(ro.get_alt_mut_rate(10, 5, 31, 0.99), 0.047902071844405425, rel_tol=1e-6, abs_tol=1e-6)
    assert ro.get_alt_mut_rate(0, 5, 31, 0.99) == -1
    assert ro.get_alt_mut_rate(10, 20, 31, 0.99) == -1


def test_single_hyp_test_types():
    # reference tests/test_unittests.py:158-174
    res = ro.single_hyp_test((100, 90), 31)
    assert isinstance(res[0], (bool, int)) and isinstance(res[1], float)
    assert all(isinstance(res[k], int) for k in (2, 3, 4)) and all(isinstance(res[k], float) for k in (5, 6, 7))


def load_run_golden():
    with gzip.open(os.path.join(GOLD, "run_golden.json.gz"), "rb") as f:
        return json.loads(f.read().decode())


def test_run_port_matches_workbook_golden():
    gold = load_run_golden()
    worst = 0.0
    n = 0
    for book in gold["books"]:
        k, ani, sig = book["ksize"], book["ani_thresh"], book["significance"]
        rows = book["rows"]
        # every 3rd evaluation of the big books keeps the CPU suite short; the GPU parity test
        # (tests/test_run_parity_gpu.py) checks all of them
        step = 3 if len(rows) > 1000 else 1
        for ne, cov, m, nc, thr, conf, alt, p, ins, p_ok in rows[::step]:
            got = ro.single_hyp_test((ne, m), k, sig, ani, cov)
            assert got[3] == nc and got[5] == thr and bool(got[0]) == bool(ins), (book["source"], ne, cov, m, got)
            assert ro.float_close(got[6], conf, 1e-12), (ne, cov, m, got[6], conf)
            assert ro.float_close(got[7], alt, 1e-12), (ne, cov, m, got[7], alt)
            if p_ok:
                assert ro.float_close(got[1], p, 1e-11), (ne, cov, m, got[1], p)
            n += 1
    assert n > 9000


def test_run_known_answer_fixture20():
    # reference tests/test_workflow.py:58-66 + SURVEY.md appendix B: only CP032507.1 overlaps the
    # sample; n_excl=3741, num_matches=2; cov=1 -> thr 706, cov=0.001 -> thr 0 and in_sample True
    db, z = _fixture20()
    counts = ro.exclusive_counts(db.hashes, db.offsets, z["sample_hashes"])
    nt = np.flatnonzero(counts["nontrivial"])
    assert len(nt) == 1
    g = int(nt[0])
    assert str(z["names"][g]) == "CP032507.1 Ectothiorhodospiraceae bacterium BW-2 chromosome, complete genome"
    assert (int(counts["n_exclusive"][g]), int(counts["n_match"][g])) == (3741, 2)
    r1 = ro.single_hyp_test((3741, 2), 31, 0.99, 0.95, 1.0)
    assert r1[5] == 706 and r1[0] is False and r1[3] == 3741
    assert ro.float_close(r1[6], 0.9893565463905609, 1e-12) and ro.float_close(r1[7], 0.054795863080029594, 1e-12)
    r2 = ro.single_hyp_test((3741, 2), 31, 0.99, 0.95, 0.001)
    assert r2[5] == 0 and r2[0] is True and r2[3] == 3 and r2[4] == 2
    assert ro.float_close(r2[1], 0.9915219633070561, 1e-12)
    # n_c == 0 sentinels
    r0 = ro.single_hyp_test((5, 0), 31, 0.99, 0.95, 0.1)
    assert (r0[3], r0[5], r0[6], r0[7], r0[1], r0[0]) == (0, 0.0, 0.0, -1.0, 1.0, False)


# ---- randomised live pinning of the train restatement against the UNMODIFIED reference core ---------------------------
@pytest.mark.skipif(not to.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_train_port_matches_reference_on_random_small_databases():
    """Hand-rolled fuzz (seeded): tiny databases built from a SMALL hash universe so that ties, twins, subsets,
    in-sketch duplicates, empty sketches and thresholds sitting exactly on a containment value all occur; the
    restatement must reproduce the reference binary's pair lines, retained genomes (order included) and the three
    banner statistics, for any thread / pass split."""
    rng = np.random.default_rng(20260101)
    checked = 0
    for case in range(40):
        n = int(rng.integers(1, 14))
        universe = rng.integers(1, 2 ** 63, size=int(rng.integers(3, 40)), dtype=np.uint64)
        parts = []
        for g in range(n):
            k = int(rng.integers(0, min(len(universe), 12) + 1))
            s = rng.choice(universe, size=k, replace=False) if k else np.zeros(0, dtype=np.uint64)
            if k and rng.random() < 0.2:                       # the same hash twice inside one sketch
                s = np.concatenate([s, s[: int(rng.integers(1, k + 1))]])
            if g and rng.random() < 0.25:                      # an exact twin / subset of an earlier sketch
                src = parts[int(rng.integers(0, g))]
                s = src.copy() if rng.random() < 0.5 else src[: len(src) // 2]
            parts.append(np.asarray(s, dtype=np.uint64))
        db = synth.from_sketches(parts)
        sizes = [len(p) for p in parts if len(p)]
        thr = float(rng.choice([0.0, 1.0, 0.5, 1 / 3, 0.95 ** 31] + ([1.0 * int(rng.integers(1, max(sizes) + 1)) / max(sizes)] if sizes else [])))
        t, p = int(rng.integers(1, 6)), int(rng.integers(1, 4))
        with tempfile.TemporaryDirectory() as d:
            ref = to.reference_train(db.hashes, db.offsets, thr, d, threads=t, passes=p)
        r = to.oracle_train(db.hashes, db.offsets, thr)
        ctx = (case, n, thr, t, p, [p_.tolist() for p_ in parts])
        assert r.lines == ref.lines, ctx
        assert list(r.selected) == list(ref.selected), ctx
        assert (r.n_distinct, r.n_singleton, r.n_index) == (ref.n_distinct, ref.n_singleton, ref.n_index), ctx
        checked += 1
    assert checked == 40
