"""`yacht run` hot path at BASELINE.json config 5 scale (run manually on the GPU box):
synthetic 10M-hash sample vs the 85k-genome reference, min_coverage_list 1 0.6 0.2 0.1, significance 0.99.
Prints one JSON line (K5 / K6 device times, algorithmic GB/s of K5, and the CPU restatement timed on a bounded sample)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yacht_b200 import _lib, synth
from oracle import run_oracle as ro

n = int(sys.argv[1]) if len(sys.argv) > 1 else 85205
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
covs = [1.0, 0.6, 0.2, 0.1]
t0 = time.time(); db = synth.make_reference_db(n, 3); gen = time.time() - t0
sample, present, cov = synth.make_sample(db, 5, n_present=2000, total_hashes=n_sample)
ctx = _lib.GpuContext(0)
ctx.load_sketches(db.hashes, db.offsets)
# the sample in page-locked memory (as a host that reads it from disk would stage it): a pageable 80 MB copy alone is ~10 ms
import ctypes
_hp = ctx.lib.ygpu_host_alloc(max(len(sample), 1) * 8)
_pinned = np.ctypeslib.as_array(ctypes.cast(_hp, ctypes.POINTER(ctypes.c_uint64)), shape=(max(len(sample), 1),))[: len(sample)]
_pinned[:] = sample
sample = _pinned
T = int(db.offsets[-1])
res = {}
for rep in range(3):
    ctx.reset_timers()
    t0 = time.perf_counter(); counts = ctx.exclusive_hashes(sample); wall5 = time.perf_counter() - t0
    nt = np.flatnonzero(counts["nontrivial"])
    t0 = time.perf_counter(); rows = ctx.hyp_test(counts["n_exclusive"][nt], counts["n_match"][nt], 31, 0.99, 0.95, covs); wall6 = time.perf_counter() - t0
    tm = ctx.timings()
    res = dict(k5_ms=tm["ms_sample"], k5_kernels_ms=tm["ms_sample_kernels"], k5_wall_ms=wall5 * 1e3, partition_ms_this_call=tm["ms_sort"],
               k6_ms=tm["ms_stats"], k6_wall_ms=wall6 * 1e3, rep=rep)
    print("rep", rep, res, flush=True)
B_run = 8 * T + 4 * T + 8 * len(sample)
# CPU restatement on a bounded sample of genomes (python sets, like the reference)
ns = 1500
sub = db.subset(range(ns))
t0 = time.perf_counter(); exp = ro.exclusive_counts(sub.hashes, sub.offsets, sample); cpu5 = time.perf_counter() - t0
ids = np.flatnonzero(exp["nontrivial"])[:200]
t0 = time.perf_counter()
for g in ids:
    for c in covs:
        ro.single_hyp_test((int(exp["n_exclusive"][g]), int(exp["n_match"][g])), 31, 0.99, 0.95, c)
cpu6 = time.perf_counter() - t0
# parity on the sub-sample
ctx.load_sketches(sub.hashes, sub.offsets)
got = ctx.exclusive_hashes(sample)
ok = all(np.array_equal(got[f], exp[f]) for f in ("n_overlap", "n_exclusive", "n_match"))
print(json.dumps(dict(workload=f"{n} reference genomes ({T} hashes), sample {len(sample)} hashes, coverages {covs}", nontrivial=int(len(nt)),
                      in_sample=int(rows["in_sample_est"][0].sum()), **res, k5_algorithmic_bytes=B_run,
                      k5_GBps=B_run / (res["k5_ms"] * 1e-3) / 1e9, k5_kernels_GBps=B_run / (max(res["k5_kernels_ms"], 1e-6) * 1e-3) / 1e9, k6_evaluations=int(len(nt) * len(covs)),
                      cpu_restatement=dict(sample=f"first {ns} genomes / first {len(ids)} nontrivial x {len(covs)} coverages",
                                           exclusive_s=cpu5, exclusive_genomes_per_s=ns / cpu5, hyp_s=cpu6,
                                           hyp_evals_per_s=len(ids) * len(covs) / max(cpu6, 1e-9)),
                      gpu_genomes_per_s=n / (res["k5_ms"] * 1e-3), gpu_evals_per_s=len(nt) * len(covs) / (res["k6_ms"] * 1e-3),
                      parity_on_subsample=bool(ok))))
