"""Timing of the sketching kernel (row f-4) on one B200: python tests/_sketch_bench.py [n_bases] [out.json]
Random ACGT with 0.1 % N, k = 31, scaled = 1000; the CPU figure is the C oracle on a 16 MB sample (one core)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sketch_oracle as so  # noqa: E402
from yacht_b200 import _lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 29
    out = sys.argv[2] if len(sys.argv) > 2 else None
    rng = np.random.default_rng(1)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    codes = rng.integers(0, 4, size=n, dtype=np.uint8)
    codes[rng.integers(0, n, size=n // 1000)] = 4
    bases = lut[codes]
    del codes
    res = {"workload": f"{n} random bases (0.1 % N), k=31, scaled=1000, one sketch", "n_bases": n}
    with _lib.GpuContext(0) as ctx:
        mh = so.max_hash_for_scaled(1000)
        for rep in range(3):
            ctx.reset_timers()
            t0 = time.perf_counter()
            h, a, off, n_kmers = ctx.sketch_sequences(bases, [0, n], 31, mh)
            wall = time.perf_counter() - t0
            tm = ctx.timings()
            print(f"rep {rep}: kernel {tm['ms_sketch']:.3f} ms, h2d {tm['ms_h2d']:.1f} ms, wall {wall * 1e3:.1f} ms, {len(h)} distinct hashes, {n_kmers} k-mers", flush=True)
        res.update({"kernel_ms": tm["ms_sketch"], "h2d_ms": tm["ms_h2d"], "wall_ms": wall * 1e3, "kmers": n_kmers, "distinct_hashes": int(len(h)),
                    "kernel_kmers_per_s": n_kmers / (tm["ms_sketch"] * 1e-3), "kernel_GBps_bases": n / (tm["ms_sketch"] * 1e-3) / 1e9})
        ctx.set_option("sketch_kernel", 2)          # A/B: the byte-wise kernel (any k) on the same input
        for rep in range(2):
            ctx.reset_timers()
            h2, a2, _, _ = ctx.sketch_sequences(bases, [0, n], 31, mh)
            tm2 = ctx.timings()
        ctx.set_option("sketch_kernel", 0)
        print(f"byte-wise kernel: {tm2['ms_sketch']:.3f} ms, same result: {bool(np.array_equal(h, h2) and np.array_equal(a, a2))}", flush=True)
        res.update({"kernel_ms_bytewise": tm2["ms_sketch"], "bytewise_same_result": bool(np.array_equal(h, h2) and np.array_equal(a, a2))})
        for rep in range(2):                         # k = 51 (two-word windows)
            ctx.reset_timers()
            ctx.sketch_sequences(bases, [0, n], 51, mh)
            tm3 = ctx.timings()
        print(f"k=51: {tm3['ms_sketch']:.3f} ms", flush=True)
        res["kernel_ms_k51"] = tm3["ms_sketch"]
        m = 1 << 24
        t0 = time.perf_counter()
        em, ea = so.sketch_records([bases[:m].tobytes()], 31, 1000)
        cpu = time.perf_counter() - t0
        hs, as_, _, _ = ctx.sketch_sequences(bases[:m], [0, m], 31, mh)
        res.update({"cpu_oracle_sample_bases": m, "cpu_oracle_s": cpu, "cpu_oracle_kmers_per_s": (m - 30) / cpu,
                    "parity_on_sample": bool(np.array_equal(hs, em) and np.array_equal(as_, ea))})
    print(json.dumps(res))
    if out:
        with open(out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
