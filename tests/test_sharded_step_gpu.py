"""The library's sharded train step (ygpu_comm_init / ygpu_load_sketches_sharded / ygpu_train_step_sharded): every rank
resident with its own genome range, index build split by hash range with the words and group streams stored into the
peers' buffers by the kernels, pairwise count by query rows, pair lists gathered over NCCL.  Bit-exact against the oracle
on every rank.  One rank runs anywhere; 2 / 4 / 8 ranks need that many GPUs (threads of one process here, as the drop-in
executable runs them; bench.py runs the same entry points with one process per GPU)."""
import threading

import numpy as np
import pytest

from oracle import train_oracle as to
from yacht_b200 import _lib, sharding, synth

pytestmark = pytest.mark.gpu
THR = 0.95 ** 31


def _run_ranks(db, nranks, thr, residency):
    """residency "genomes": rank r holds the sketches of its genome range (words cross NVLink in the level-1 scatter);
    "hashes": rank r holds, of every sketch, the hashes of its hash range (only work items cross NVLink)."""
    offsets = np.ascontiguousarray(db.offsets, dtype=np.uint64)
    bounds = sharding.split_rows_by_size(offsets, nranks)
    cuts = sharding.hash_cuts(int(db.hashes.max()) if len(db.hashes) else 0, nranks)
    sizes = db.sizes.astype(np.uint32)
    uid = _lib.comm_unique_id()
    out = [None] * nranks
    err = [None] * nranks

    def work(r):
        try:
            with _lib.GpuContext(r) as ctx:
                ctx.comm_init(r, nranks, uid)
                g0, g1 = int(bounds[r]), int(bounds[r + 1])
                sl = db.hashes[int(offsets[g0]):int(offsets[g1])]
                ph, po = sharding.hashrange_share(db.hashes, offsets, int(cuts[r]), int(cuts[r + 1]), last=r == nranks - 1)
                for rep in range(2):            # the second step reuses the shared exchange buffers
                    if residency == "replicated":
                        ctx.load_sketches(db.hashes, offsets)
                        st, F = ctx.train_step_replicated(thr)
                        continue
                    if residency == "genomes":
                        ctx.load_sketches_sharded(sl, offsets, g0, g1)
                    else:
                        ctx.load_sketches_hashrange(ph, po, sizes, g0, g1)
                    st, F = ctx.train_step_sharded(thr)
                out[r] = (st, ctx.pairs_host(F))
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


def _check(db, nranks, thr):
    ref = to.oracle_train(db.hashes, db.offsets, thr)
    for residency in ("hashes", "genomes", "replicated"):
        _check_one(db, nranks, thr, residency, ref)


def _check_one(db, nranks, thr, residency, ref):
    res = _run_ranks(db, nranks, thr, residency)
    for st, pairs in res:
        assert (st["n_distinct"], st["n_singleton"], st["n_index"]) == (ref.n_distinct, ref.n_singleton, ref.n_index)
        assert (st["n_postings"], st["n_increments"]) == (ref.n_postings, ref.n_increments)
        assert len(pairs) == len(ref.pairs)
        assert np.array_equal(pairs["i"], ref.pairs["i"]) and np.array_equal(pairs["j"], ref.pairs["j"])
        assert np.array_equal(pairs["count"], ref.pairs["count"])


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_sharded_step_matches_oracle(nranks):
    if _lib.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    # large enough for two partition levels and 32-bit remaining keys (the sharded path refuses tiny databases)
    db = synth.make_reference_db(6000, 11, mean_size=3000, sd_size=800)
    _check(db, nranks, THR)


@pytest.mark.parametrize("nranks", [1, 2, 8])
def test_sharded_step_larger_groups(nranks):
    """Hashes held by 5 .. 60 genomes: groups of up to 16 members travel as self-contained work items to the owners of
    the query rows, larger ones through the posting stream every rank receives."""
    if _lib.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    rng = np.random.default_rng(77)
    base = synth.make_reference_db(5000, 31, mean_size=3000, sd_size=600)
    parts = [base.sketch(g) for g in range(base.n)]
    for size in (5, 9, 16, 17, 33, 60):
        for rep in range(6):
            members = rng.choice(base.n, size=size, replace=False)
            common = rng.integers(0, synth.MAX_HASH, size=int(rng.integers(20, 200)), dtype=np.uint64)
            for m in members:
                parts[m] = np.unique(np.concatenate([parts[m], common]))
    db = synth.from_sketches(parts)
    _check(db, nranks, 0.01)
    _check(db, nranks, THR)


@pytest.mark.parametrize("nranks", [1, 2])
def test_sharded_step_edge_rows(nranks):
    if _lib.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    # empty sketches, tiny genomes (many genome boundaries per tile), threshold 0 (every overlapping pair is flagged)
    rng = np.random.default_rng(5)
    base = synth.make_reference_db(3000, 12, mean_size=4000, sd_size=500)
    parts = [base.sketch(g) for g in range(base.n)]
    parts[0] = np.zeros(0, dtype=np.uint64)
    parts[1500] = np.zeros(0, dtype=np.uint64)
    parts[-1] = np.zeros(0, dtype=np.uint64)
    for k in range(200):
        parts.append(np.unique(rng.choice(parts[7 + k], size=int(rng.integers(1, 20)), replace=False)))
    db = synth.from_sketches(parts)
    _check(db, nranks, 0.0)
    _check(db, nranks, THR)


@pytest.mark.parametrize("nranks", [1, 2])
def test_skewed_database_is_refused_unanimously_and_runs_replicated(nranks):
    """A hash held by thousands of genomes overflows a shared-memory bucket: every rank refuses the sharded step with the
    same error (nobody is left waiting in a collective), and the replicated step (full index on every rank, rows split by
    measured work) gives the oracle's answer."""
    if _lib.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    rng = np.random.default_rng(8)
    base = synth.make_reference_db(5000, 41, mean_size=2500, sd_size=500)
    parts = [base.sketch(g) for g in range(base.n)]
    members = rng.choice(base.n, size=2500, replace=False)
    common = rng.integers(0, synth.MAX_HASH, size=40, dtype=np.uint64)
    for m in members:
        parts[m] = np.unique(np.concatenate([parts[m], common]))
    db = synth.from_sketches(parts)
    ref = to.oracle_train(db.hashes, db.offsets, THR)
    for residency in ("hashes", "genomes"):
        with pytest.raises(_lib.YgpuError, match=r"\(-4\)"):
            _run_ranks(db, nranks, THR, residency)
    _check_one(db, nranks, THR, "replicated", ref)
