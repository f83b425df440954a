"""Whole-executable wall time of the drop-in `run_yacht_train_core` at BASELINE.json's full size (run manually on
the GPU box): N synthetic sourmash signature files on disk -> pair files + selected_result.tsv, exactly the command
line utils.run_yacht_train_core issues (reference: src/yacht/utils.py:143-145).  The reference core itself needs
~52 GB and >10 min at this size, so only its 10k-genome run (tests/_scale_check.py) stands beside it.
usage: python tests/_train_wall.py [N] [threads] [out.json]"""
import glob, json, os, shutil, subprocess, sys, tempfile, time
from multiprocessing import Pool
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yacht_b200 import sigio, synth
from oracle import train_oracle as to

n = int(sys.argv[1]) if len(sys.argv) > 1 else 85205
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
out_json = sys.argv[3] if len(sys.argv) > 3 else None
thr = 0.95 ** 31
root = tempfile.mkdtemp(prefix="yacht_wall_")
t0 = time.time(); db = synth.make_reference_db(n, 3); gen_s = time.time() - t0
os.makedirs(os.path.join(root, "signatures"))
paths = [os.path.join(root, "signatures", f"g{g:07d}.sig") for g in range(n)]


def _write(rng):
    for g in range(rng[0], rng[1]):
        sigio.write_signature(paths[g], f"genome_{g}", db.hashes[int(db.offsets[g]):int(db.offsets[g + 1])])
    return rng[1] - rng[0]


if __name__ == "__main__":
    t0 = time.time()
    step = max(1, n // (4 * threads))
    with Pool(threads) as pool:                      # fork: the workers see db without copying it
        done = sum(pool.map(_write, [(a, min(n, a + step)) for a in range(0, n, step)]))
    write_s = time.time() - t0
    nbytes = sum(os.path.getsize(p) for p in paths[:: max(1, n // 200)]) / len(paths[:: max(1, n // 200)]) * n
    fl = os.path.join(root, "training_sig_files.tsv")
    with open(fl, "w") as f:
        f.write("\n".join(paths) + "\n")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yacht_b200", "run_yacht_train_core")
    runs = []
    for rep in range(2):
        wd = os.path.join(root, f"wd{rep}"); os.makedirs(wd)
        sel = os.path.join(wd, "selected_result.tsv")
        t0 = time.time()
        cp = subprocess.run([exe, "-t", str(threads), "-c", repr(thr), "-p", "1", fl, wd, sel], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t0
        assert cp.returncode == 0, cp.stderr[-2000:]
        nlines = sum(1 for f in glob.glob(os.path.join(wd, "*_*.txt")) for _ in open(f))
        kept = sum(1 for _ in open(sel))
        runs.append(dict(wall_s=wall, phases_ms=to.parse_phase_times(cp.stdout), gpu_lines=[l for l in cp.stdout.splitlines() if l.startswith("[gpu")],
                         pair_lines=nlines, genomes_kept=kept))
        print(f"run {rep}: wall {wall:.2f}s", runs[-1], flush=True)
    res = dict(genomes=n, hashes=int(db.offsets[-1]), threads=threads, sig_bytes_estimate=int(nbytes), generate_s=gen_s, write_sig_s=write_s, runs=runs,
               pairs_per_s=n * (n - 1) / min(r["wall_s"] for r in runs))
    print(json.dumps(res))
    if out_json:
        with open(out_json, "w") as f:
            json.dump(res, f, indent=1)
    shutil.rmtree(root, ignore_errors=True)
