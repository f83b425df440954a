"""bench.py's reference arm on a tiny configuration (CPU only): the JSON line carries the keys the driver reads, the
reference's own output is checked against the oracle port, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

import pytest

from oracle import train_oracle as to

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *argv):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


@pytest.mark.skipif(not to.reference_available(), reason="oracle/_ref not present")
def test_reference_arm_json_line():
    cp = _run(None, "--impl", "reference", "--genomes", "300", "--seed", "11", "--steps", "2", "--warmup", "1", "--no-reference-cache")
    assert cp.returncode == 0, cp.stderr[-2000:]
    lines = [l for l in cp.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["steps_run"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "300 genomes" in d["config"]["workload"] and d["config"]["genomes"] == 300
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "ALL 300 of 300" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    cp = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--genomes", "300")
    assert cp.returncode == 0 and cp.stdout.strip() == ""
