"""N > 1 host logic on CPU: world_size-2 gloo run of the row sharding + pair-list gather
(the same code path bench.py and multi-GPU callers use with NCCL on device tensors)."""
import os
import socket

import numpy as np
import pytest

from yacht_b200 import sharding, synth
from yacht_b200._lib import PAIR_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import train_oracle as to
        db = synth.make_reference_db(300, 17, mean_size=200, sd_size=60)
        full = to.oracle_train(db.hashes, db.offsets, 0.3).pairs
        full3 = np.zeros(len(full), dtype=PAIR_DTYPE)
        for f in ("i", "j", "count"):
            full3[f] = full[f]
        bounds = sharding.split_rows_by_work(db.sizes.astype(np.float64), world)
        mine = sharding.owned_pairs(full3, int(bounds[rank]), int(bounds[rank + 1]))
        merged = sharding.all_gather_pairs(mine, len(mine), world)
        ok = merged.tobytes() == np.sort(full3, order=["i", "j"]).tobytes()
        # a rank with nothing to report must not break the exchange
        empty = sharding.all_gather_pairs(mine if rank == 0 else mine[:0], len(mine) if rank == 0 else 0, world)
        ok = ok and len(empty) == len(sharding.owned_pairs(full3, int(bounds[0]), int(bounds[1])))
        q.put((rank, bool(ok), len(mine), len(full3)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_matches_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == res[0][3]      # the shards partition the pair list
    assert all(r[2] > 0 for r in res)


def test_split_rows_by_work_properties():
    rng = np.random.default_rng(0)
    for n, parts in [(1000, 8), (5, 8), (0, 3), (97, 1)]:
        w = rng.integers(0, 1000, size=n).astype(np.float64)
        b = sharding.split_rows_by_work(w, parts)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b.astype(np.int64)) >= 0)
        if n >= 100 * parts:
            loads = [float((w[b[k]:b[k + 1]] + sharding.ROW_CONSTANT).sum()) for k in range(parts)]
            assert max(loads) < 1.3 * (sum(loads) / parts)


def test_slice_bounds_and_size_split():
    b = sharding.slice_bounds(10, 4)
    assert b.tolist() == [0, 3, 6, 9, 10]                      # equal slices, the last one short
    assert sharding.slice_bounds(0, 3).tolist() == [0, 0, 0, 0]
    assert sharding.slice_bounds(8, 1).tolist() == [0, 8]
    off = np.array([0, 10, 10, 30, 60, 100], dtype=np.uint64)   # sizes 10, 0, 20, 30, 40
    r = sharding.split_rows_by_size(off, 2)
    assert r[0] == 0 and r[-1] == 5 and 0 < r[1] < 5
    sizes = np.diff(off.astype(np.int64))
    assert abs(int(sizes[:r[1]].sum()) - int(sizes[r[1]:].sum())) <= int(sizes.max()) + 2 * int(sharding.ROW_CONSTANT)


# ---- the sharded build's host orchestration (yacht_b200/sharding.py) with a stand-in for the library ---------------
class _FakeCtx:
    """Implements the three C-ABI calls of the hash-range sharded build in numpy (hash space split by hash % world),
    writing / reading the caller's buffers through raw pointers exactly like libyachtgpu does."""

    def __init__(self, db, fail=False):
        self.db, self.fail = db, fail
        self.finished = None

    @staticmethod
    def groups_of(db, keep):
        gid = np.repeat(np.arange(db.n, dtype=np.int64), np.diff(db.offsets.astype(np.int64)))
        order = np.lexsort((gid, db.hashes))
        h, g = db.hashes[order], gid[order]
        out = []
        start = 0
        for end in list(np.flatnonzero(h[1:] != h[:-1]) + 1) + [len(h)]:
            if end - start >= 2 and keep(int(h[start])):
                out.append(tuple(int(x) for x in g[start:end]))
            start = end
        return out

    def index_partial(self, part, nparts):
        from yacht_b200._lib import YgpuError
        if self.fail:
            raise YgpuError("does not qualify")
        groups = self.groups_of(self.db, lambda hv: hv % nparts == part)
        self.gid = np.array([x for grp in groups for x in grp], dtype=np.int32)
        self.rem = np.array([len(grp) - 1 - k for grp in groups for k in range(len(grp))], dtype=np.int16)
        st = dict(n_hashes=int(sum(1 for hv in self.db.hashes if int(hv) % nparts == part)), n_distinct=0, n_singleton=0, n_index=len(groups),
                  n_postings=len(self.gid), n_increments=sum(len(grp) ** 2 for grp in groups),
                  n_row_items=len(self.gid) - len(groups), has_duplicates=0)
        return st, len(self.gid)

    def index_stream_copy(self, gptr, rptr):
        import ctypes
        ctypes.memmove(gptr, self.gid.ctypes.data, self.gid.nbytes)
        ctypes.memmove(rptr, self.rem.ctypes.data, self.rem.nbytes)

    def index_finish(self, gptr, rptr, n_entries, rb, re, total):
        import ctypes
        gid = np.frombuffer(ctypes.string_at(gptr, 4 * n_entries), dtype=np.int32)
        rem = np.frombuffer(ctypes.string_at(rptr, 2 * n_entries), dtype=np.int16)
        groups, x = [], 0
        while x < n_entries:
            if rem[x] == 0:          # padding (or the tail of a group, consumed below)
                x += 1
                continue
            L = int(rem[x]) + 1
            assert list(rem[x:x + L]) == list(range(L - 1, -1, -1))
            groups.append(tuple(int(v) for v in gid[x:x + L]))
            x += L
        self.finished = dict(groups=sorted(groups), rows=(rb, re), total=dict(total))


def _sharded_worker(rank, world, port, q, fail_rank):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = synth.make_reference_db(120, 4, mean_size=80, sd_size=20)
        ctx = _FakeCtx(db, fail=(rank == fail_rank))
        res = sharding.build_index_sharded(ctx, db.offsets, rank, world, torch.device("cpu"))
        if res is None:
            q.put((rank, "fallback", None))
            return
        rb, re, total = res
        want = sorted(_FakeCtx.groups_of(db, lambda hv: True))
        ok = ctx.finished["groups"] == want and ctx.finished["rows"] == (rb, re)
        ok = ok and total["n_postings"] == sum(len(g) for g in want) and total["n_hashes"] == int(db.offsets[-1])
        q.put((rank, "ok" if ok else "mismatch", (rb, re)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_sharded_build_orchestration_two_ranks(fail_rank):
    """Every rank ends up with the complete group stream (own slice + the peer's, padding ignored), the summed
    statistics and its own row range; when one rank's share does not qualify, BOTH ranks take the fallback."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q, fail_rank)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    if fail_rank >= 0:
        assert [r[1] for r in res] == ["fallback", "fallback"], res
    else:
        assert [r[1] for r in res] == ["ok", "ok"], res
        assert res[0][2][0] == 0 and res[0][2][1] == res[1][2][0] and res[1][2][1] == 120      # the row ranges partition [0, n)


def _ingest_worker(rank, world, port, q):
    import ctypes
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = synth.make_reference_db(50, 8, mean_size=61, sd_size=9)          # T not a multiple of the world size
        T = int(db.offsets[-1])
        sb = sharding.slice_bounds(T, world)
        mine = torch.from_numpy(db.hashes[int(sb[rank]):int(sb[rank + 1])].view(np.int64).copy())

        class Ctx:
            def load_sketches_device(self, hptr, optr, n):
                self.h = np.frombuffer(ctypes.string_at(hptr, 8 * T), dtype=np.uint64).copy()
                self.o = np.frombuffer(ctypes.string_at(optr, 8 * (n + 1)), dtype=np.uint64).copy()

        ctx = Ctx()
        sharding.load_sketches_sharded(ctx, mine, db.offsets, T, rank, world, torch.device("cpu"))
        q.put((rank, bool(np.array_equal(ctx.h, db.hashes) and np.array_equal(ctx.o, db.offsets.astype(np.uint64)))))
    finally:
        dist.destroy_process_group()


def test_sharded_ingest_two_ranks():
    """Each rank contributes only its slice of the flat hash array; after the all-gather every rank hands the library
    the complete array (the short last slice's padding stays beyond T)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ingest_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
