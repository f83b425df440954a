"""GPU parity of the train hot path (K2 index build, K3+K4 count/flag) against the CPU oracle.

Bar: bit-exact on shared-hash counts, flagged ordered pairs and index statistics.
All calls go through the C ABI (ctypes on libyachtgpu.so).
"""
import numpy as np
import pytest

from oracle import train_oracle as to
from yacht_b200 import synth

pytestmark = pytest.mark.gpu

THR = 0.95 ** 31


class _Pairs:
    """(i, j, count) triples of a pair list, compared field by field with numpy (a threshold that flags most pairs
    yields tens of millions of them); on a mismatch the assertion shows the first differing triples."""

    def __init__(self, p):
        self.t = np.stack([np.asarray(p[f], dtype=np.int64) for f in ("i", "j", "count")], axis=1) if len(p) else np.zeros((0, 3), np.int64)

    def __eq__(self, other):
        return self.t.shape == other.t.shape and bool(np.array_equal(self.t, other.t))

    def __repr__(self):
        return f"<{len(self.t)} pairs, first {self.t[:5].tolist()}>"


def _pairs_tuple(p):
    return _Pairs(p)


def _check(ctx, db, thr, expect_path=None, ref=None):
    """Both index builds (1 = MSD partition when the input qualifies, 0 = general sort path) against the oracle."""
    if ref is None:
        ref = to.oracle_train(db.hashes, db.offsets, thr)
    out = None
    for path in (1, 0):
        ctx.set_option("index_path", path)
        ctx.load_sketches(db.hashes, db.offsets)
        st = ctx.build_index()
        assert (st["n_distinct"], st["n_singleton"], st["n_index"]) == (ref.n_distinct, ref.n_singleton, ref.n_index), (path, st)
        assert st["n_postings"] == ref.n_postings, (path, st)
        assert st["n_increments"] == ref.n_increments, (path, st)
        for count_kernel in (1, 2):          # dense row accumulator per CTA / one warp per row (hash table)
            ctx.set_option("count_kernel", count_kernel)
            got = ctx.pairwise_flag(thr)
            assert _pairs_tuple(got) == _pairs_tuple(ref.pairs), (path, count_kernel)
        ctx.set_option("count_kernel", 0)
        # the containment test as per-genome integer thresholds (default) and as the fp64 expression per pair
        ctx.set_option("count_thresholds", 0)
        try:
            assert _pairs_tuple(ctx.pairwise_flag(thr)) == _pairs_tuple(ref.pairs), (path, "fp64 per pair")
        finally:
            ctx.set_option("count_thresholds", 1)
        if path == 1:
            out = (st, got)
            if expect_path is not None:
                assert st["index_path"] == expect_path, st
        else:
            assert st["index_path"] == 0
    ctx.set_option("index_path", 1)
    return out


def test_edge_set(gpu_ctx):
    # SURVEY.md appendix B hand-made set: identical twins, an empty sketch, duplicates inside a sketch
    parts = [np.arange(1, 11), np.arange(1, 11), np.zeros(0), np.array([1, 2, 3, 4, 5] + list(range(100, 107))),
             np.array([7, 7, 7, 200])]
    db = synth.from_sketches([np.asarray(p, dtype=np.uint64) for p in parts])
    st, got = _check(gpu_ctx, db, 0.4)
    assert st["has_duplicates"] == 1
    assert len(got) == 8
    _check(gpu_ctx, db, 0.0)
    _check(gpu_ctx, db, 1.0)


def test_empty_and_trivial(gpu_ctx):
    db = synth.from_sketches([np.zeros(0, dtype=np.uint64)] * 3)
    _check(gpu_ctx, db, 0.5)
    db = synth.from_sketches([np.array([5, 9], dtype=np.uint64)])
    _check(gpu_ctx, db, 0.5)
    db = synth.from_sketches([])
    gpu_ctx.load_sketches(db.hashes, db.offsets)
    gpu_ctx.build_index()
    assert len(gpu_ctx.pairwise_flag(0.1)) == 0


def test_full_range_hashes(gpu_ctx):
    big = np.array([0, 1, 2**63, 2**64 - 1], dtype=np.uint64)
    db = synth.from_sketches([big, big[1:], big[:2], np.array([2**64 - 1], dtype=np.uint64)])
    _check(gpu_ctx, db, 0.0)
    _check(gpu_ctx, db, 0.5)


@pytest.mark.parametrize("n,seed,mean", [(300, 1, 600), (1000, 7, 600), (2000, 11, 1500)])
def test_synthetic_flagged_pairs(gpu_ctx, n, seed, mean):
    db = synth.make_reference_db(n, seed, mean_size=mean, sd_size=mean / 3)
    _check(gpu_ctx, db, THR)


@pytest.mark.parametrize("n,seed", [(400, 3)])
def test_synthetic_all_counts(gpu_ctx, n, seed):
    # threshold 0 emits every ordered pair with a non-zero count: the whole count matrix
    db = synth.make_reference_db(n, seed, mean_size=500, sd_size=100)
    _check(gpu_ctx, db, 0.0)


def test_row_ranges_union(gpu_ctx):
    db = synth.make_reference_db(1000, 5, mean_size=600, sd_size=200)
    gpu_ctx.load_sketches(db.hashes, db.offsets)
    gpu_ctx.build_index()
    full = gpu_ctx.pairwise_flag(THR)
    b = gpu_ctx.row_partition(4)
    assert b[0] == 0 and b[-1] == db.n and all(b[k] <= b[k + 1] for k in range(4))
    parts = [gpu_ctx.pairwise_flag(THR, int(b[k]), int(b[k + 1])) for k in range(4)]
    merged = np.sort(np.concatenate(parts), order=["i", "j"])
    assert _pairs_tuple(merged) == _pairs_tuple(full)


def test_skewed_long_postings(gpu_ctx):
    # conserved-core hashes present in many genomes: long posting lists, touched-list overflow.  The final buckets
    # holding them overflow shared memory: only those buckets leave the partition path (sort-based side route) ...
    db = synth.make_reference_db(6000, 9, mean_size=60, sd_size=10, min_size=20, core_hashes=6, core_lo=0.7, core_hi=0.95)
    st, got = _check(gpu_ctx, db, 0.05, expect_path=1)
    assert 1 <= st["big_buckets"] <= 6
    # ... and with the side route switched off the whole database falls back to the general path (same answer)
    gpu_ctx.set_option("big_buckets", 0)
    try:
        gpu_ctx.load_sketches(db.hashes, db.offsets)
        st0 = gpu_ctx.build_index()
        assert st0["index_path"] == 0 and st0["big_buckets"] == 0
        assert gpu_ctx.pairwise_flag(0.05).tobytes() == got.tobytes()
    finally:
        gpu_ctx.set_option("big_buckets", 1)


def test_oversized_buckets_with_duplicates_and_two_levels(gpu_ctx):
    # two partition levels (d2 > 0), core hashes in 3 500-4 500 genomes each (> 3 072), one sketch holding a core hash twice
    db = synth.make_reference_db(5000, 23, mean_size=300, sd_size=60, min_size=100, core_hashes=3, core_lo=0.7, core_hi=0.9)
    parts = [db.sketch(g) for g in range(db.n)]
    core = np.intersect1d(np.intersect1d(parts[0], parts[1]), parts[2])
    if core.size:
        parts[1] = np.concatenate([parts[1], core[:1]])
    db2 = synth.from_sketches(parts)
    st, _ = _check(gpu_ctx, db2, 0.02, expect_path=1)
    assert st["big_buckets"] >= 1


def test_long_groups_inside_msd_buckets(gpu_ctx):
    # posting lists of ~1000 genomes that still fit a shared-memory bucket: the CTA-wide group path
    db = synth.make_reference_db(1500, 10, mean_size=60, sd_size=10, min_size=20, core_hashes=5, core_lo=0.5, core_hi=0.8)
    st, _ = _check(gpu_ctx, db, 0.05, expect_path=1)
    assert st["n_postings"] > 3000


def test_msd_two_level_partition(gpu_ctx):
    # enough hashes for two partition levels (d2 > 0) and duplicates inside sketches
    db = synth.make_reference_db(2500, 12, mean_size=1200, sd_size=300)
    parts = [db.sketch(g) for g in range(db.n)]
    parts[3] = np.concatenate([parts[3], parts[3][:40]])          # the same hash twice in one sketch
    parts[7] = np.concatenate([parts[7], parts[3][:10], parts[3][:10]])
    db2 = synth.from_sketches(parts)
    st, _ = _check(gpu_ctx, db2, THR, expect_path=1)
    assert st["has_duplicates"] == 1


def test_count_kernel_tiled_and_u16_variants(gpu_ctx):
    # the accumulator geometries that large N selects (column tiles, packed 16-bit counters), forced at small N
    db = synth.make_reference_db(1500, 14, mean_size=500, sd_size=150)
    ref = to.oracle_train(db.hashes, db.offsets, 0.1)
    gpu_ctx.load_sketches(db.hashes, db.offsets)
    gpu_ctx.build_index()
    try:
        for tile_w, u16 in [(0, 1), (401, 0), (400, 1), (64, 1), (3, 0)]:
            gpu_ctx.set_option("force_tile_w", tile_w)
            gpu_ctx.set_option("force_u16", u16)
            got = gpu_ctx.pairwise_flag(0.1)
            assert _pairs_tuple(got) == _pairs_tuple(ref.pairs), (tile_w, u16)
    finally:
        gpu_ctx.set_option("force_tile_w", 0)
        gpu_ctx.set_option("force_u16", 0)


def test_full_size_properties_85k(gpu_ctx, config_db):
    """BASELINE.json's full size (85 205 genomes): properties that do not need the oracle --
    the two independent count kernels agree, flagged pairs are consistent with the sketch sizes,
    every planted twin pair above the threshold is found, and row-range shards union to the whole."""
    db = config_db("config3")
    gpu_ctx.load_sketches(db.hashes, db.offsets)
    st = gpu_ctx.build_index()
    assert st["index_path"] == 1 and st["n_hashes"] == int(db.offsets[-1])
    assert st["n_distinct"] - st["n_singleton"] == st["n_index"]
    gpu_ctx.set_option("count_kernel", 1)
    dense = gpu_ctx.pairwise_flag(THR)
    gpu_ctx.set_option("count_kernel", 2)
    warp = gpu_ctx.pairwise_flag(THR)
    gpu_ctx.set_option("count_kernel", 0)
    assert dense.tobytes() == warp.tobytes()
    sizes = db.sizes
    i, j, c = dense["i"].astype(np.int64), dense["j"].astype(np.int64), dense["count"].astype(np.int64)
    assert np.all(i != j) and np.all(c >= 1) and np.all(c <= np.minimum(sizes[i], sizes[j]))
    assert np.all(c / sizes[i] >= THR)                                   # every emitted pair passes the reference's test
    key = i * db.n + j
    assert np.all(np.diff(key) > 0)                                      # sorted by (i, j), no duplicates
    same_cluster = db.cluster[i] == db.cluster[j]
    assert np.all(same_cluster & (db.cluster[i] >= 0))                   # random 64-bit hashes never collide across clusters
    # exact counts for a sample of flagged pairs, recomputed on the host
    rng = np.random.default_rng(0)
    for k in rng.choice(len(dense), size=200, replace=False):
        a, b = db.sketch(int(i[k])), db.sketch(int(j[k]))
        assert int(c[k]) == int(np.intersect1d(a, b, assume_unique=True).size)
    # shards
    b4 = gpu_ctx.row_partition(4)
    parts = [gpu_ctx.pairwise_flag(THR, int(b4[k]), int(b4[k + 1])) for k in range(4)]
    merged = np.sort(np.concatenate(parts), order=["i", "j"])
    assert merged.tobytes() == dense.tobytes()
    # the general (sort) index path gives the same answer at full size
    gpu_ctx.set_option("index_path", 0)
    st0 = gpu_ctx.build_index()
    alt = gpu_ctx.pairwise_flag(THR)
    gpu_ctx.set_option("index_path", 1)
    assert st0["index_path"] == 0 and alt.tobytes() == dense.tobytes()
    for f in ("n_distinct", "n_singleton", "n_index", "n_postings", "n_increments"):
        assert st0[f] == st[f], f


@pytest.mark.parametrize("per_block", [7, 32, 2000])
def test_streamed_ingest_matches_plain_load(gpu_ctx, per_block):
    # ygpu_upload_begin / _block / _finish (blocks uploaded out of order, an empty sketch, a block larger than one
    # bounce buffer when per_block = 2000) leaves the same resident sketches as ygpu_load_sketches
    db = synth.make_reference_db(2400, 21, mean_size=700, sd_size=200)
    parts = [db.sketch(g) for g in range(db.n)]
    parts[5] = np.zeros(0, dtype=np.uint64)
    db2 = synth.from_sketches(parts)
    ref = to.oracle_train(db2.hashes, db2.offsets, THR)
    gpu_ctx.load_sketches_streamed(db2.hashes, db2.offsets, per_block)
    st = gpu_ctx.build_index()
    assert (st["n_hashes"], st["n_distinct"], st["n_singleton"]) == (int(db2.offsets[-1]), ref.n_distinct, ref.n_singleton)
    assert _pairs_tuple(gpu_ctx.pairwise_flag(THR)) == _pairs_tuple(ref.pairs)


def test_random_small_databases(gpu_ctx):
    """Seeded fuzz over tiny databases drawn from a small hash universe (ties, twins, subsets, in-sketch duplicates,
    empty sketches, thresholds sitting exactly on a containment value) -- the same generator that pins the oracle to
    the reference binary in tests/test_oracle_pinned.py.  Both index paths, both count kernels, bit-exact."""
    rng = np.random.default_rng(20260102)
    for case in range(30):
        n = int(rng.integers(1, 14))
        universe = rng.integers(1, 2 ** 63, size=int(rng.integers(3, 40)), dtype=np.uint64)
        parts = []
        for g in range(n):
            k = int(rng.integers(0, min(len(universe), 12) + 1))
            s = rng.choice(universe, size=k, replace=False) if k else np.zeros(0, dtype=np.uint64)
            if k and rng.random() < 0.2:
                s = np.concatenate([s, s[: int(rng.integers(1, k + 1))]])
            if g and rng.random() < 0.25:
                src = parts[int(rng.integers(0, g))]
                s = src.copy() if rng.random() < 0.5 else src[: len(src) // 2]
            parts.append(np.asarray(s, dtype=np.uint64))
        db = synth.from_sketches(parts)
        sizes = [len(p) for p in parts if len(p)]
        thr = float(rng.choice([0.0, 1.0, 0.5, 1 / 3, THR] + ([1.0 * int(rng.integers(1, max(sizes) + 1)) / max(sizes)] if sizes else [])))
        _check(gpu_ctx, db, thr)
