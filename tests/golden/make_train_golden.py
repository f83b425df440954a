"""Produce the train-path golden vectors by RUNNING THE UNMODIFIED REFERENCE CORE.

Run in the build container (needs /root/reference and oracle/_ref, built by `make -C oracle`):
    python tests/golden/make_train_golden.py

Outputs (committed):
  tests/golden/fixture20.npz     the reference's own test fixture tests/testdata/20_genomes_sketches.zip
                                 (+ tests/testdata/sample.sig.zip) reduced to arrays: sketch hashes,
                                 offsets, names, md5sums, abundances; sample hashes + abundances.  The
                                 signatures carry "license": "CC0".
  tests/golden/train_golden.json for each case: what run_yacht_train_core (compiled from
                                 /root/reference/src/cpp/main.cpp) printed and wrote -- the three
                                 index statistics, the sorted pair-file lines and the selected
                                 genome ids in output order.  Cases: the 20-genome fixture, the
                                 hand-made edge set of SURVEY.md appendix B, an equal-size-twins tie
                                 case, and seeded synthetic sets (inputs are regenerated from the
                                 seed by yacht_b200.synth, only outputs are stored).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("YACHT_REFERENCE", "/root/reference")

from oracle import train_oracle as to  # noqa: E402
from yacht_b200 import sigio, synth  # noqa: E402


def edge_set():
    parts = [np.arange(1, 11), np.arange(1, 11), np.zeros(0), np.array([1, 2, 3, 4, 5] + list(range(100, 107))),
             np.array([7, 7, 7, 200])]
    return synth.from_sketches([np.asarray(p, dtype=np.uint64) for p in parts])


def ties_set():
    # groups of mutually similar genomes of EQUAL size: which twin survives depends on the visit
    # order std::sort produces (main.cpp:379-381, README: "randomly selected")
    rng = np.random.default_rng(123)
    parts = []
    for grp in range(12):
        base = rng.integers(0, synth.MAX_HASH, size=40, dtype=np.uint64)
        for m in range(int(rng.integers(2, 5))):
            x = base.copy()
            x[m] = rng.integers(0, synth.MAX_HASH, dtype=np.uint64)   # same size, 39/40 shared
            parts.append(np.unique(x))
    for k in range(30):
        parts.append(np.unique(rng.integers(0, synth.MAX_HASH, size=int(rng.integers(20, 60)), dtype=np.uint64)))
    order = rng.permutation(len(parts))
    return synth.from_sketches([parts[i] for i in order])


SYNTH_CASES = [
    dict(name="synth_n300_s1", n=300, seed=1, mean_size=600, sd_size=200, thr=0.95 ** 31),
    dict(name="synth_n1000_s7", n=1000, seed=7, mean_size=600, sd_size=200, thr=0.95 ** 31),
    dict(name="synth_n500_s21_thr0.5", n=500, seed=21, mean_size=300, sd_size=100, thr=0.5),
]


def run_ref(db, thr, threads=2, passes=2):
    with tempfile.TemporaryDirectory() as d:
        r = to.reference_train(db.hashes, db.offsets, thr, d, threads=threads, passes=passes)
    return dict(thr=thr, n_distinct=r.n_distinct, n_singleton=r.n_singleton, n_index=r.n_index, lines=r.lines,
                selected=[int(x) for x in r.selected])


def main():
    assert to.reference_available(), "build oracle/_ref first: make -C oracle"
    out = {}
    # ---- the reference's own fixture ------------------------------------------------------------
    sigs = sigio.read_sig_zip(os.path.join(REF, "tests/testdata/20_genomes_sketches.zip"))
    sigs = [s for s in sigs if s.ksize == 31]
    sample = sigio.load_signature_with_ksize(os.path.join(REF, "tests/testdata/sample.sig.zip"), 31)
    db = synth.from_sketches([s.mins for s in sigs])
    ab = np.concatenate([s.abundances if s.abundances is not None else np.zeros(0, np.int64) for s in sigs])
    np.savez_compressed(os.path.join(HERE, "fixture20.npz"), hashes=db.hashes, offsets=db.offsets,
                        names=np.array([s.name for s in sigs]), md5=np.array([s.md5sum for s in sigs]),
                        abundances=ab, has_abund=np.array([s.abundances is not None for s in sigs]),
                        max_hash=np.array([s.max_hash for s in sigs], dtype=np.uint64),
                        sample_hashes=sample.mins, sample_abundances=sample.abundances,
                        sample_name=np.array(str(sample.name or "sample")), sample_md5=np.array(sample.md5sum),
                        sample_max_hash=np.array(sample.max_hash, dtype=np.uint64))
    out["fixture20"] = run_ref(db, 0.95 ** 31)
    out["fixture20_thr0"] = run_ref(db, 0.0)
    out["edge"] = run_ref(edge_set(), 0.4)
    out["edge_thr0"] = run_ref(edge_set(), 0.0, threads=1, passes=1)
    out["edge_thr1"] = run_ref(edge_set(), 1.0, threads=3, passes=1)
    out["ties"] = run_ref(ties_set(), 0.9)
    for c in SYNTH_CASES:
        db = synth.make_reference_db(c["n"], c["seed"], mean_size=c["mean_size"], sd_size=c["sd_size"])
        out[c["name"]] = dict(run_ref(db, c["thr"]), gen=dict(n=c["n"], seed=c["seed"], mean_size=c["mean_size"], sd_size=c["sd_size"]),
                              T=int(db.offsets[-1]), checksum=int(np.bitwise_xor.reduce(db.hashes)))
    with open(os.path.join(HERE, "train_golden.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    for k, v in out.items():
        print(k, "pairs", len(v["lines"]), "selected", len(v["selected"]), (v["n_distinct"], v["n_singleton"], v["n_index"]), file=sys.stderr)


if __name__ == "__main__":
    main()
