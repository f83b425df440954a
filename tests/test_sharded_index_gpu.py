"""Hash-range sharded index build (ygpu_index_partial / _stream_copy / _finish) against the CPU oracle.

The multi-GPU flow (yacht_b200/sharding.py: build_index_sharded) is emulated on ONE GPU: the shares of the
hash space are built one after the other on the same context, their group streams are concatenated the way
the NCCL all-gather lays them out (padded slices), and the work lists are built from the complete stream.
Bar: bit-exact statistics, counts and flagged pairs.
"""
import numpy as np
import pytest

from oracle import train_oracle as to
from yacht_b200 import sharding, synth

pytestmark = pytest.mark.gpu

THR = 0.95 ** 31


def _pairs_tuple(p):
    return [(int(a), int(b), int(c)) for a, b, c in zip(p["i"], p["j"], p["count"])]


def _sharded_build(ctx, db, nparts, rows=None):
    import torch
    dev = torch.device("cuda", 0)
    ctx.load_sketches(db.hashes, db.offsets)
    parts = []
    for p in range(nparts):
        st, n_p = ctx.index_partial(p, nparts)
        g = torch.zeros(max(n_p, 1), dtype=torch.int32, device=dev)
        r = torch.zeros(max(n_p, 1), dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        ctx.index_stream_copy(g.data_ptr(), r.data_ptr())
        parts.append((st, n_p, g, r))
    m = max(max(n for _, n, _, _ in parts), 1)
    gid = torch.zeros(nparts * m, dtype=torch.int32, device=dev)
    rem = torch.zeros(nparts * m, dtype=torch.int16, device=dev)
    for p, (_, n_p, g, r) in enumerate(parts):
        gid[p * m: p * m + n_p] = g[:n_p]
        rem[p * m: p * m + n_p] = r[:n_p]
    torch.cuda.synchronize()
    total = {k: sum(st[k] for st, _, _, _ in parts) for k in sharding.SUMMED_STATS}
    rb, re = rows if rows is not None else (0, db.n)
    ctx.index_finish(gid.data_ptr(), rem.data_ptr(), nparts * m, rb, re, total)
    return total, [n for _, n, _, _ in parts]


@pytest.mark.parametrize("n,seed,mean,nparts", [(300, 1, 600, 2), (1000, 7, 600, 3), (2500, 12, 1200, 1), (2500, 12, 1200, 2), (2500, 12, 1200, 8)])
def test_sharded_build_matches_oracle(gpu_ctx, n, seed, mean, nparts):
    db = synth.make_reference_db(n, seed, mean_size=mean, sd_size=mean / 3)
    ref = to.oracle_train(db.hashes, db.offsets, THR)
    total, sizes = _sharded_build(gpu_ctx, db, nparts)
    assert (total["n_distinct"], total["n_singleton"], total["n_index"]) == (ref.n_distinct, ref.n_singleton, ref.n_index)
    assert total["n_postings"] == ref.n_postings == sum(sizes)
    assert total["n_increments"] == ref.n_increments
    assert total["n_hashes"] == int(db.offsets[-1])
    if nparts > 1 and n >= 1000:
        assert min(sizes) > 0                                   # every share of the hash space holds groups
    for count_kernel in (1, 2):
        gpu_ctx.set_option("count_kernel", count_kernel)
        got = gpu_ctx.pairwise_flag(THR)
        assert _pairs_tuple(got) == _pairs_tuple(ref.pairs), count_kernel
    gpu_ctx.set_option("count_kernel", 0)


def test_sharded_build_row_ranges_and_all_counts(gpu_ctx):
    # threshold 0: the whole count matrix; work lists built for one row range at a time (what each rank does)
    db = synth.make_reference_db(400, 3, mean_size=500, sd_size=100)
    ref = to.oracle_train(db.hashes, db.offsets, 0.0)
    full = np.zeros(len(ref.pairs), dtype=sharding.PAIR_DTYPE)
    for f in ("i", "j", "count"):
        full[f] = ref.pairs[f]
    bounds = sharding.split_rows_by_size(db.offsets, 3)
    got = []
    for k in range(3):
        rb, re = int(bounds[k]), int(bounds[k + 1])
        _sharded_build(gpu_ctx, db, 3, rows=(rb, re))
        mine = gpu_ctx.pairwise_flag(0.0, rb, re)
        assert mine.tobytes() == sharding.owned_pairs(full, rb, re).tobytes()
        got.append(mine)
    merged = np.sort(np.concatenate(got), order=["i", "j"])
    assert merged.tobytes() == full.tobytes()


def test_sharded_build_long_groups_and_duplicates(gpu_ctx):
    # posting lists of ~1000 genomes (indirect work items pointing into the gathered stream) and in-sketch duplicates
    db = synth.make_reference_db(1500, 10, mean_size=60, sd_size=10, min_size=20, core_hashes=5, core_lo=0.5, core_hi=0.8)
    parts = [db.sketch(g) for g in range(db.n)]
    parts[3] = np.concatenate([parts[3], parts[3][:7]])
    db2 = synth.from_sketches(parts)
    ref = to.oracle_train(db2.hashes, db2.offsets, 0.05)
    total, _ = _sharded_build(gpu_ctx, db2, 4)
    assert total["has_duplicates"] >= 1
    assert total["n_postings"] == ref.n_postings and total["n_increments"] == ref.n_increments
    assert _pairs_tuple(gpu_ctx.pairwise_flag(0.05)) == _pairs_tuple(ref.pairs)


def test_sharded_build_rejects_unqualified_input(gpu_ctx):
    from yacht_b200._lib import YgpuError
    # conserved-core hashes overflow a shared-memory bucket: no SHARDED partition path, callers fall back to build_index
    db = synth.make_reference_db(6000, 9, mean_size=60, sd_size=10, min_size=20, core_hashes=6, core_lo=0.7, core_hi=0.95)
    gpu_ctx.load_sketches(db.hashes, db.offsets)
    with pytest.raises(YgpuError):
        gpu_ctx.index_partial(0, 2)
    st = gpu_ctx.build_index()                      # the unsharded build keeps the partition path for everything else
    assert st["index_path"] == 1 and st["big_buckets"] >= 1
