"""BASELINE.json configs[3] -- "full-GTDB shape (~400k genomes, skewed posting lists)" -- on ONE B200 (run manually on the GPU
box): Zipf(1.3) cluster sizes capped at 20 000, 200 conserved core hashes each present in 1-10 % of all genomes.
The index build takes the MSD partition path with the oversized buckets on the sort-based side route; the pairwise count runs
in work-balanced row chunks so that the flagged-pair buffers of one chunk stay small.  Prints one JSON line.
usage: python tests/_config4.py [genomes=400000] [row_chunks=16] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yacht_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out_json = sys.argv[3] if len(sys.argv) > 3 else None
thr = 0.95 ** 31
t0 = time.time()
db = synth.make_skewed_db(n, 4, zipf_cap=max(2, n // 20), core_hashes=max(1, n // 2000))
gen_s = time.time() - t0
T = int(db.offsets[-1])
csz = np.bincount(db.cluster[db.cluster >= 0]) if (db.cluster >= 0).any() else np.zeros(1, dtype=np.int64)
print(f"generated {n} genomes, {T} hashes in {gen_s:.0f}s; clusters: {len(csz)}, largest {int(csz.max())}, >=1000 members: {int((csz >= 1000).sum())}", flush=True)
ctx = _lib.GpuContext(0)
t0 = time.perf_counter(); ctx.load_sketches(db.hashes, db.offsets); load_s = time.perf_counter() - t0
res = {}
for rep in range(2):
    ctx.reset_timers()
    t0 = time.perf_counter(); st = ctx.build_index(); index_s = time.perf_counter() - t0
    tm = ctx.timings()
    res = dict(index_wall_s=index_s, ms_partition=tm["ms_sort"], ms_grouping=tm["ms_index"], ms_group_kernel=tm["ms_group"], stats=st)
    print("index", rep, res, flush=True)
bounds = ctx.row_partition(chunks)
F = 0
count_ms = 0.0
t0 = time.perf_counter()
sizes = db.sizes
ok = True
for k in range(chunks):
    ctx.reset_timers()
    nk = ctx.pairwise_flag_device(thr, int(bounds[k]), int(bounds[k + 1]))
    count_ms += ctx.timings()["ms_count"]
    F += nk
    if k in (0, chunks // 2) and nk:      # spot check of one chunk: every emitted pair passes the reference's test, no pair twice
        p = ctx.pairs_host(nk)
        i, j, c = p["i"].astype(np.int64), p["j"].astype(np.int64), p["count"].astype(np.int64)
        ok &= bool(np.all(c / sizes[i] >= thr)) and bool(np.all(np.diff(i * n + j) > 0)) and bool(np.all(c <= np.minimum(sizes[i], sizes[j])))
        for q in np.random.default_rng(k).choice(len(p), size=min(50, len(p)), replace=False):
            ok &= int(c[q]) == int(np.intersect1d(db.sketch(int(i[q])), db.sketch(int(j[q])), assume_unique=True).size)
count_wall = time.perf_counter() - t0
line = dict(workload=f"config 4: {n} genomes, {T} hashes, Zipf(1.3) clusters capped at {max(2, n // 20)}, {max(1, n // 2000)} core hashes, seed 4",
            genomes=n, hashes=T, largest_cluster=int(csz.max()), generate_s=gen_s, load_s=load_s, **res, row_chunks=chunks,
            count_kernel_ms=count_ms, count_wall_s=count_wall, flagged_pairs=int(F), increments_W=int(st["n_increments"]),
            smem_increments_per_s=st["n_increments"] / 2 / max(count_ms * 1e-3, 1e-9), pairs_per_s=n * (n - 1) / (res["index_wall_s"] + count_wall),
            spot_checks_ok=bool(ok))
print(json.dumps(line), flush=True)
if out_json:
    with open(out_json, "w") as f:
        json.dump(line, f, indent=1)
