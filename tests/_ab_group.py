"""Developer aid (GPU box): A/B of the two MSD grouping kernels at full size -- same statistics, byte-identical
flagged pairs, per-phase device times."""
import sys, time
sys.path.insert(0, '/root/repo')
from yacht_b200 import _lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 85205
db = synth.make_reference_db(n, 3)
ctx = _lib.GpuContext(0)
ctx.load_sketches(db.hashes, db.offsets)
res = {}
for gk in (0, 1, 0, 1, 1):
    ctx.set_option("group_kernel", gk)
    ctx.reset_timers()
    st = ctx.build_index()
    p = ctx.pairwise_flag(0.95 ** 31)
    tm = ctx.timings()
    print("group_kernel", gk, {k: round(tm[k], 3) for k in ("ms_sort", "ms_index", "ms_count", "ms_pairsort")}, len(p), flush=True)
    key = (tuple(st[f] for f in ("n_distinct", "n_singleton", "n_index", "n_postings", "n_increments", "n_row_items", "has_duplicates")), p.tobytes())
    res.setdefault(gk, key)
    assert res[gk] == key, "not deterministic"
print("stats equal:", res[0][0] == res[1][0], res[0][0], res[1][0])
print("pairs identical:", res[0][1] == res[1][1])
