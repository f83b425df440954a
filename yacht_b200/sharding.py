"""Row sharding of the all-vs-all job across ranks and the gather of the compacted pair lists.

The reference splits query rows into contiguous chunks per thread and per pass
(src/cpp/main.cpp:338-349) because every row of the count matrix is independent given the read-only
index.  Here the same split is made across GPUs (one process per GPU): rank r flags the pairs whose
smaller genome id lies in its row range (both directions of each pair are tested by that owner),
and the variable-length pair lists are exchanged with one all-gather of the counts followed by one
all-gather of the padded lists (NCCL on device tensors; gloo on CPU tensors in the tests).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from ._lib import PAIR_DTYPE

ROW_CONSTANT = 64.0   # per-row fixed cost in "increment" units (same constant as ygpu_row_partition)


def split_rows_by_work(work: np.ndarray, nparts: int) -> np.ndarray:
    """Contiguous row ranges of nearly equal (work + constant) -- the host mirror of ygpu_row_partition."""
    n = int(work.shape[0])
    bounds = np.full(nparts + 1, n, dtype=np.uint32)
    bounds[0] = 0
    if n and nparts > 1:
        cs = np.cumsum(work.astype(np.float64) + ROW_CONSTANT)
        targets = cs[-1] * np.arange(1, nparts, dtype=np.float64) / nparts
        # part k ends after the first row at which the running cost reaches k/nparts of the total
        bounds[1:nparts] = np.minimum(np.searchsorted(cs, targets, side="left") + 1, n)
    return bounds


def owned_pairs(pairs: np.ndarray, row_begin: int, row_end: int) -> np.ndarray:
    """The subset of an (i, j)-sorted ordered-pair list that the owner of rows [row_begin, row_end)
    reports: pairs whose smaller genome id falls in the range."""
    lo = np.minimum(pairs["i"], pairs["j"])
    return pairs[(lo >= row_begin) & (lo < row_end)]


def all_gather_pairs(local, n_local: int, world: int, device=None, group=None) -> np.ndarray:
    """Every rank contributes `n_local` pairs (`local`: int32 torch tensor of shape [>= 3 * n_local] on
    `device`, or a PAIR_DTYPE numpy array); returns the union sorted by (i, j) on every rank."""
    import torch
    import torch.distributed as dist

    if isinstance(local, np.ndarray):
        flat = torch.from_numpy(np.ascontiguousarray(local).view(np.int32).copy()) if n_local else torch.zeros(0, dtype=torch.int32)
        if device is not None:
            flat = flat.to(device)
    else:
        flat = local
    dev = flat.device
    if world == 1:
        merged = flat[: 3 * n_local].cpu().numpy().view(PAIR_DTYPE)
        return np.sort(merged, order=["i", "j"])
    cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
    allc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, cnt, group=group)
    sizes = [int(x) for x in allc.tolist()]
    m = max(max(sizes), 1)
    mine = torch.zeros(3 * m, dtype=torch.int32, device=dev)
    if n_local:
        mine[: 3 * n_local] = flat[: 3 * n_local]
    allp = torch.empty(world * 3 * m, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allp, mine, group=group)
    parts = [allp[r * 3 * m: r * 3 * m + 3 * sizes[r]] for r in range(world)]
    cat = torch.cat(parts).view(-1, 3)
    # rank-ordered concatenation is deterministic; ranges interleave (the owner of row a also
    # reports (b, a)), so one final ordering by (i, j) -- on the device that holds the lists
    if cat.shape[0]:
        key = (cat[:, 0].to(torch.int64) << 32) | cat[:, 1].to(torch.int64)
        cat = cat[torch.argsort(key)]
    return cat.contiguous().cpu().numpy().reshape(-1).view(PAIR_DTYPE)


def slice_bounds(total: int, world: int) -> np.ndarray:
    """Equal contiguous slices of the flat hash array (the last one may be short): slice r = [b[r], b[r+1])."""
    per = (int(total) + world - 1) // world if world else 0
    return np.minimum(np.arange(world + 1, dtype=np.int64) * per, int(total))


def load_sketches_sharded(ctx, pinned_slice, offsets: np.ndarray, total: int, rank: int, world: int, device, group=None):
    """Sharded ingest: every rank copies only ITS slice of the flat hash array from (pinned) host
    memory over its own PCIe link, the slices are all-gathered over NVLink (NCCL), and the assembled array is
    handed to the library device-to-device.  `pinned_slice` is a pinned int64 torch tensor holding
    hashes[b[rank]:b[rank+1]] with b = slice_bounds(total, world).  Returns the assembled device tensor."""
    import torch
    import torch.distributed as dist

    per = (int(total) + world - 1) // world
    full = torch.empty(max(per, 1) * world, dtype=torch.int64, device=device)
    mine = full[rank * per: (rank + 1) * per]
    n_mine = int(pinned_slice.numel())
    if n_mine:
        mine[:n_mine].copy_(pinned_slice, non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(full, mine, group=group)
    d_off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.uint64).view(np.int64)).to(device, non_blocking=True)
    _sync(device)                                        # the library works on its own stream
    ctx.load_sketches_device(full.data_ptr(), d_off.data_ptr(), int(offsets.shape[0]) - 1)
    return full


SUMMED_STATS = ("n_hashes", "n_distinct", "n_singleton", "n_index", "n_postings", "n_increments", "n_row_items", "has_duplicates")


def split_rows_by_size(offsets: np.ndarray, nparts: int) -> np.ndarray:
    """Contiguous row ranges holding nearly equal numbers of sketch hashes (known before any index exists)."""
    sizes = np.diff(np.asarray(offsets).astype(np.int64)).astype(np.float64)
    return split_rows_by_work(sizes, nparts)


def _sync(device) -> None:
    """Order torch's stream against the library's (CUDA devices only; the gloo tests run the same code on CPU tensors)."""
    import torch
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream(device).synchronize()


_stream_buffers = {}     # (device index, world) -> (gid, rem) device tensors reused across builds


def _stream_buffer(device, world: int, m: int):
    import torch
    key = (str(device), world)
    buf = _stream_buffers.get(key)
    if buf is None or buf[0].numel() < world * m:
        cap = world * (m + m // 16 + 1024)
        buf = (torch.zeros(cap, dtype=torch.int32, device=device), torch.zeros(cap, dtype=torch.int16, device=device))
        _stream_buffers[key] = buf
    return buf[0][: world * m], buf[1][: world * m]


def build_index_sharded(ctx, offsets: np.ndarray, rank: int, world: int, device, group=None, bounds: Optional[np.ndarray] = None):
    """Index build split by hash range across the ranks (include/yacht_gpu.h: ygpu_index_partial / _finish).

    Every rank holds all sketches.  Rank r partitions and groups only its share of the hash space, the ranks
    all-gather their group streams over NCCL (padded to the longest; padding entries carry follow-count 0 and
    are ignored), and each rank builds the work lists of its own query rows (`bounds`: row ranges per rank,
    default split_rows_by_size(offsets, world)).  Returns (row_begin, row_end, total statistics), or None when
    the database does not qualify (the caller then runs ctx.build_index())."""
    import torch
    import torch.distributed as dist
    from ._lib import YgpuError

    try:
        st, n_r = ctx.index_partial(rank, world)
        ok = 1
    except YgpuError:
        st, n_r, ok = {k: 0 for k in SUMMED_STATS}, 0, 0
    head = torch.tensor([ok, n_r] + [int(st[k]) for k in SUMMED_STATS], dtype=torch.int64, device=device)
    if world > 1:
        allh = torch.empty(world * head.numel(), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allh, head, group=group)
        allh = allh.view(world, -1).cpu()
    else:
        allh = head.view(1, -1).cpu()
    if int(allh[:, 0].min()) == 0:
        return None
    m = max(int(allh[:, 1].max()), 1)
    total = {k: int(allh[:, 2 + i].sum()) for i, k in enumerate(SUMMED_STATS)}
    total["has_duplicates"] = 1 if total["has_duplicates"] else 0
    gid, rem = _stream_buffer(device, world, m)
    mine_g, mine_r = gid[rank * m: (rank + 1) * m], rem[rank * m: (rank + 1) * m]
    if n_r < m:                                            # padding of this rank's slice: follow-count 0 = ignored
        mine_r[n_r:].zero_()
    _sync(device)                                          # the library writes on its own stream
    ctx.index_stream_copy(mine_g.data_ptr(), mine_r.data_ptr())
    if world > 1:
        dist.all_gather_into_tensor(gid, mine_g, group=group)
        dist.all_gather_into_tensor(rem.view(torch.uint8), mine_r.view(torch.uint8), group=group)   # NCCL has no int16
        _sync(device)
    if bounds is None:
        bounds = split_rows_by_size(offsets, world)
    rb, re = int(bounds[rank]), int(bounds[rank + 1])
    ctx.index_finish(gid.data_ptr(), rem.data_ptr(), world * m, rb, re, total)
    return rb, re, total
