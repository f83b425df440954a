"""The reference's own tests, re-run against this repo's mirrors of the same functions / CLI.

Modelled on reference tests/test_workflow.py:10-66, tests/test_y_integration_tests.py:38-85,
tests/test_unit.py:11-20 and tests/test_unittests.py:86-111,158-174 -- same inputs (the 20-genome
fixture and sample, rebuilt as sourmash zip files from tests/golden/fixture20.npz), same assertions.
"""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from yacht_b200 import sigio, xlsx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PRESENT = "CP032507.1 Ectothiorhodospiraceae bacterium BW-2 chromosome, complete genome"


@pytest.fixture(scope="module")
def fixture_files(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("wf"))
    z = np.load(os.path.join(GOLD, "fixture20.npz"))
    off = z["offsets"]
    sk = []
    for g in range(20):
        a, b = int(off[g]), int(off[g + 1])
        sk.append(dict(name=str(z["names"][g]), mins=z["hashes"][a:b], abundances=z["abundances"][a:b]))
    ref_zip = os.path.join(d, "20_genomes_sketches.zip")
    sigio.write_sig_zip(ref_zip, sk, 31)
    sample_zip = os.path.join(d, "sample.sig.zip")
    sigio.write_sig_zip(sample_zip, [dict(name=str(z["sample_name"]), mins=z["sample_hashes"], abundances=z["sample_abundances"])], 31)
    return d, ref_zip, sample_zip


def _yacht(*argv):
    return subprocess.run([sys.executable, "-m", "yacht_b200", *argv], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_full_workflow(fixture_files):
    d, ref_zip, sample_zip = fixture_files
    # reference test_workflow.py:24-26: yacht train --force --ref_file ... --ksize 31 --prefix ... --ani_thresh 0.95 --outdir ...
    res = _yacht("train", "--force", "--ref_file", ref_zip, "--ksize", "31", "--prefix", "20_genomes_trained",
                 "--ani_thresh", "0.95", "--outdir", d, "--num_threads", "4")
    assert res.returncode == 0, res.stderr[-3000:]
    expected = [os.path.join(d, "20_genomes_trained_config.json"), os.path.join(d, "20_genomes_trained_processed_manifest.tsv")]
    for f in expected:
        assert os.path.exists(f) and os.stat(f).st_size > 200     # the reference asks > 291 bytes with its longer paths
    cfg = json.load(open(expected[0]))
    assert cfg["ksize"] == 31 and cfg["ani_thresh"] == 0.95 and cfg["scale"] == 1000      # test_y_integration_tests.py:63-66
    manifest = pd.read_csv(expected[1], sep="\t")
    assert list(manifest.columns) == ["organism_name", "md5sum", "num_unique_kmers_in_genome_sketch",
                                      "num_total_kmers_in_genome_sketch", "genome_scale_factor"]
    assert len(manifest) == 20                                    # no pair reaches the threshold: all genomes retained
    inter = os.path.join(d, "20_genomes_trained_intermediate_files")
    assert os.path.isdir(os.path.join(inter, "comparison_files")) and os.path.exists(os.path.join(inter, "selected_result.tsv"))

    # reference test_workflow.py:48: yacht run --json ... --sample_file ... --significance 0.99 --min_coverage_list 0.001 --outdir ... --show_all
    res = _yacht("run", "--json", expected[0], "--sample_file", sample_zip, "--significance", "0.99", "--min_coverage_list", "0.001",
                 "--outdir", d, "--show_all", "--num_threads", "4")
    assert res.returncode == 0, res.stderr[-3000:]
    abundance_file = os.path.join(d, "results", "result.xlsx")
    assert os.path.exists(abundance_file) and os.path.exists(os.path.join(d, "results", "result_all.txt"))
    sheets = xlsx.read_xlsx(abundance_file)
    assert list(sheets.keys()) == ["min_coverage0.001"]
    df = sheets["min_coverage0.001"]
    row = df[df["organism_name"] == PRESENT]
    assert str(row["in_sample_est"].values[0]) == "True"                                   # test_workflow.py:61
    assert row["num_matches"].values[0] == 2                                               # :63
    assert row["acceptance_threshold_with_coverage"].values[0] == 0                        # :65
    assert list(df.columns) == ["organism_name", "num_unique_kmers_in_genome_sketch", "num_total_kmers_in_genome_sketch", "scale_factor",
                                "num_exclusive_kmers_in_sample_sketch", "num_total_kmers_in_sample_sketch", "min_coverage",
                                "in_sample_est", "p_vals", "num_exclusive_kmers_to_genome", "num_exclusive_kmers_to_genome_coverage",
                                "num_matches", "acceptance_threshold_with_coverage", "actual_confidence_with_coverage",
                                "alt_confidence_mut_rate_with_coverage"]
    assert row["num_exclusive_kmers_to_genome"].values[0] == 3741
    txt = pd.read_csv(os.path.join(d, "results", "result_all.txt"), sep="\t")
    assert len(txt) == 1 and txt["organism_name"].iloc[0] == PRESENT

    # the first run left the packed sketch cache next to the reference's intermediate files ...
    assert os.path.exists(os.path.join(inter, "ygpu_cache", "meta.json"))
    assert sorted(os.listdir(os.path.join(inter, "signatures"))) == sorted(m + ".sig" for m in manifest["md5sum"])   # ... and touched nothing else

    # --keep_raw adds the raw_result sheet with *_wo_coverage columns; default coverage list
    res = _yacht("run", "--json", expected[0], "--sample_file", sample_zip, "--outdir", d, "--keep_raw", "--show_all")
    assert res.returncode == 0, res.stderr[-3000:]
    assert "Packed sketch cache found" in res.stdout + res.stderr          # second run: no signature file parsed
    sheets = xlsx.read_xlsx(abundance_file)
    # with the DEFAULT list the first element is the int 1 (argparse's type= is not applied to defaults), so the
    # reference names that sheet "min_coverage1" (run_YACHT.py:59,252) -- kept
    assert list(sheets.keys()) == ["raw_result", "min_coverage1", "min_coverage0.5", "min_coverage0.1", "min_coverage0.05", "min_coverage0.01"]
    raw = sheets["raw_result"]
    assert "acceptance_threshold_wo_coverage" in raw.columns and raw["acceptance_threshold_wo_coverage"].values[0] == 706


def test_train_refuses_existing_dir_and_bad_input(fixture_files):
    d, ref_zip, _ = fixture_files
    res = _yacht("train", "--ref_file", ref_zip, "--ksize", "31", "--prefix", "20_genomes_trained", "--outdir", d)
    assert res.returncode != 0 and "already exists" in res.stderr                          # make_training_data_from_sketches.py:94-97
    res = _yacht("train", "--ref_file", os.path.join(d, "nope.zip"), "--ksize", "31", "--prefix", "x", "--outdir", d)
    assert res.returncode != 0 and "does not exist" in res.stderr                          # test_workflow.py:69-74 (rc != 0)
    res = _yacht("train", "--ref_file", os.path.join(d, "nope.txt"), "--ksize", "31", "--prefix", "x", "--outdir", d)
    assert res.returncode != 0 and "is not a zip file" in res.stderr


def test_get_alt_mut_rate_1():
    # reference tests/test_unit.py:11-20, verbatim
    from yacht_b200.hypothesis_recovery_src import get_alt_mut_rate
    assert get_alt_mut_rate(100, 10000, 21, significance=0.99) == -1
    assert np.isclose(get_alt_mut_rate(10, 0, 21), 0.28015945851802826)
    assert np.isclose(get_alt_mut_rate(10, 0, 31), 0.19963312102481723)
    assert np.isclose(get_alt_mut_rate(10, 5, 21), 0.0698992155957967)
    assert np.isclose(get_alt_mut_rate(10, 5, 31), 0.047902071848511696)
    assert np.isclose(get_alt_mut_rate(10, 9, 21), 0.02169068099465221)
    assert np.isclose(get_alt_mut_rate(100, 10, 11), 0.2397729973308742)
    assert np.isclose(get_alt_mut_rate(1000, 0, 1), 0.9999899497147453)


def test_get_alt_mut_rate_unittests():
    # reference tests/test_unittests.py:86-111, verbatim
    from yacht_b200.hypothesis_recovery_src import get_alt_mut_rate
    assert math.isclose(get_alt_mut_rate(10, 5, 31, 0.99), 0.047902071844405425, rel_tol=1e-6, abs_tol=1e-6)
    assert get_alt_mut_rate(0, 5, 31, 0.99) == -1
    assert get_alt_mut_rate(10, 20, 31, 0.99) == -1


def test_single_hyp_test():
    # reference tests/test_unittests.py:158-174, verbatim
    from yacht_b200.hypothesis_recovery_src import single_hyp_test
    result = single_hyp_test((100, 90), 31)
    in_sample_est, p_val, num_exclusive_kmers, num_exclusive_kmers_coverage, num_matches, \
        acceptance_threshold_with_coverage, actual_confidence_with_coverage, alt_confidence_mut_rate_with_coverage = result
    assert isinstance(in_sample_est, int)
    assert isinstance(p_val, float)
    assert isinstance(num_exclusive_kmers, int)
    assert isinstance(num_exclusive_kmers_coverage, int)
    assert isinstance(num_matches, int)
    assert isinstance(acceptance_threshold_with_coverage, float)
    assert isinstance(actual_confidence_with_coverage, float)
    assert isinstance(alt_confidence_mut_rate_with_coverage, float)
    from oracle import run_oracle as ro
    exp = ro.single_hyp_test((100, 90), 31)
    assert result[0] == exp[0] and result[2:6] == exp[2:6]
    for a, b in zip((result[1], result[6], result[7]), (exp[1], exp[6], exp[7])):
        assert ro.float_close(a, b)
