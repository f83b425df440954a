"""Pin parity at BASELINE.json's own sizes by RUNNING THE UNMODIFIED REFERENCE CORE on them.

Run in the build container (needs oracle/_ref, built from /root/reference by `make -C oracle`):
    python tests/golden/make_config_digests.py [config2] [config3] [config4s]

For each configuration the synthetic database is regenerated from its seed (yacht_b200.synth), written
as sourmash signature files to a scratch directory, and oracle/_ref/run_yacht_train_core_ref (compiled
from /root/reference/src/cpp/main.cpp, reference Makefile flags) is run on the file list with
`-t <cores> -p <passes>` (results are invariant under -t / -p: SURVEY.md 8a).  Stored in
tests/golden/config_digests.json (committed):
  banners          the three index statistics the reference prints (main.cpp:242-244)
  F                number of pair lines over all <pass>_<tid>.txt files
  pairs_sha256     sha256 of the sorted pair lines joined by "\n" (line = i,j,jaccard,c_ij,c_ji, main.cpp:305)
  pairs_ij_sha256  sha256 of the (i, j) pairs as little-endian int32, sorted by (i, j)
  selected_sha256  sha256 of the selected genome ids (file-list line indices) in output order, one per line
                   (selected_result.tsv holds paths; ids make the digest independent of the scratch directory)
  n_selected, phases_ms (the reference's own timers here: 8 vCPUs), generator arguments
Only outputs are stored; inputs are regenerated from the seed by the tests.
"""
import glob
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import train_oracle as to  # noqa: E402
from yacht_b200 import sigio, synth  # noqa: E402

THR = 0.95 ** 31
OUT = os.path.join(HERE, "config_digests.json")

CONFIGS = {
    # BASELINE.json configs[1]: synthetic 10k-genome ref DB (seed 2)
    "config2": dict(gen=dict(n=10000, seed=2), passes=1),
    # BASELINE.json configs[2]: GTDB-rs214-representatives shape, 85 205 genomes (seed 3); -p 4 keeps the
    # reference's dense int matrix at 7.3 GB per pass
    "config3": dict(gen=dict(n=85205, seed=3), passes=4),
    # BASELINE.json configs[3] scaled down 20x: Zipf(1.3) cluster sizes (cap 8 000 so that one cluster exceeds the
    # 3 072-word shared-memory bucket), conserved core hashes present in 1-10 % of all genomes
    "config4s": dict(gen=dict(n=20000, seed=4, zipf_clusters=True, zipf_cap=8000, core_hashes=40, mean_size=1200.0,
                              sd_size=300.0), passes=1),
}

_db = None
_paths = None


def _write(rng):
    for g in range(rng[0], rng[1]):
        sigio.write_signature(_paths[g], f"genome_{g}", _db.hashes[int(_db.offsets[g]):int(_db.offsets[g + 1])])
    return rng[1] - rng[0]


def digest_lines(lines):
    return hashlib.sha256("\n".join(sorted(lines)).encode()).hexdigest()


def digest_ij(lines):
    """sha256 of the (i, j) int32 little-endian pairs sorted by (i, j): the cheap digest for pair lists of
    tens of millions of lines (config4s), where formatting every line again on the test side is too slow."""
    ij = np.array([l.split(",", 2)[:2] for l in lines], dtype=np.int64).reshape(-1, 2)
    key = (ij[:, 0] << 32) | ij[:, 1]
    ij = ij[np.argsort(key, kind="stable")].astype("<i4")
    return hashlib.sha256(np.ascontiguousarray(ij).tobytes()).hexdigest()


def digest_ids(ids):
    return hashlib.sha256("\n".join(str(int(g)) for g in ids).encode()).hexdigest()


def run_config(name, cfg, threads):
    global _db, _paths
    t0 = time.time()
    _db = synth.make_reference_db(**cfg["gen"])
    n, T = _db.n, int(_db.offsets[-1])
    print(f"[{name}] generated {n} genomes, {T} hashes in {time.time() - t0:.1f}s", flush=True)
    root = tempfile.mkdtemp(prefix=f"yacht_digest_{name}_", dir=os.environ.get("YACHT_SCRATCH", "/tmp"))
    try:
        os.makedirs(os.path.join(root, "signatures"))
        _paths = [os.path.join(root, "signatures", f"g{g:07d}.sig") for g in range(n)]
        t0 = time.time()
        step = max(1, n // (4 * threads))
        with Pool(threads) as pool:
            pool.map(_write, [(a, min(n, a + step)) for a in range(0, n, step)])
        print(f"[{name}] wrote {n} .sig files in {time.time() - t0:.1f}s", flush=True)
        fl = os.path.join(root, "training_sig_files.tsv")
        with open(fl, "w") as f:
            f.write("\n".join(_paths) + "\n")
        paths = list(_paths)
        _db = None                      # the reference core needs the RAM at 85k genomes
        wd = os.path.join(root, "wd")
        os.makedirs(wd)
        out, wall = to.run_core_binary(to.REF_BIN, fl, wd, THR, threads=threads, passes=cfg["passes"])
        res = to.parse_core_outputs(wd, paths, os.path.join(wd, "selected_result.tsv"), out)
        entry = dict(generator=cfg["gen"], threshold=THR, genomes=n, hashes=T,
                     banners=dict(n_distinct=res.n_distinct, n_singleton=res.n_singleton, n_index=res.n_index),
                     F=len(res.lines), pairs_sha256=digest_lines(res.lines), pairs_ij_sha256=digest_ij(res.lines),
                     n_selected=int(len(res.selected)), selected_sha256=digest_ids(res.selected),
                     reference_cmd=f"run_yacht_train_core_ref -t {threads} -p {cfg['passes']} -c {THR!r}",
                     reference_phases_ms=to.parse_phase_times(out), reference_wall_s=wall,
                     first_lines=res.lines[:3])
        print(f"[{name}] reference: wall {wall:.1f}s {entry['reference_phases_ms']} F={entry['F']} kept={entry['n_selected']}", flush=True)
        return entry
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    if not to.reference_available():
        raise SystemExit("oracle/_ref/run_yacht_train_core_ref missing: make -C oracle (needs /root/reference)")
    which = sys.argv[1:] or list(CONFIGS)
    threads = os.cpu_count() or 8
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    for name in which:
        data[name] = run_config(name, CONFIGS[name], threads)
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
            f.write("\n")
