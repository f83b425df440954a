"""BASELINE.json's own configurations, pinned by the UNMODIFIED reference core: tests/golden/config_digests.json holds
what oracle/_ref/run_yacht_train_core_ref printed and wrote at 10 000 genomes (config 2), 85 205 genomes (config 3) and
a scaled-down skewed config 4 (made by tests/golden/make_config_digests.py).  Here the CPU oracle port is held against
those digests at the sizes it finishes in seconds; tests/test_config_digests_gpu.py holds the CUDA path and the
drop-in executable against all of them."""
import json
import os

import numpy as np
import pytest

from oracle import train_oracle as to
from yacht_b200 import pairfmt, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load_digests():
    with open(os.path.join(HERE, "golden", "config_digests.json")) as f:
        return json.load(f)


def as_pairs(p):
    out = np.zeros(len(p), dtype=[("i", "<i4"), ("j", "<i4"), ("count", "<i4")])
    out["i"], out["j"], out["count"] = p["i"], p["j"], p["count"]
    return out


def check_against_digest(entry, sizes, stats, pairs, selected):
    """stats: (n_distinct, n_singleton, n_index); pairs: structured (i, j, count) sorted by (i, j)."""
    b = entry["banners"]
    assert tuple(int(x) for x in stats) == (b["n_distinct"], b["n_singleton"], b["n_index"])
    assert len(pairs) == entry["F"]
    if "pairs_ij_sha256" in entry:
        assert pairfmt.digest_ij(pairs) == entry["pairs_ij_sha256"]
    if entry["F"] <= 2_000_000:          # the text lines (floats included); beyond that only (i, j) -- counts are checked against the port
        assert pairfmt.digest_lines(pairfmt.pair_lines(pairs, sizes)) == entry["pairs_sha256"]
    assert len(selected) == entry["n_selected"]
    assert pairfmt.digest_ids(selected) == entry["selected_sha256"]


@pytest.mark.parametrize("name", ["config2", "config4s"])
def test_oracle_port_matches_reference_digest(name):
    entry = load_digests()[name]
    db = synth.make_reference_db(**entry["generator"])
    assert (db.n, int(db.offsets[-1])) == (entry["genomes"], entry["hashes"])
    r = to.oracle_train(db.hashes, db.offsets, entry["threshold"])
    check_against_digest(entry, db.sizes, (r.n_distinct, r.n_singleton, r.n_index), as_pairs(r.pairs), r.selected)
