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
    bounds = np.zeros(nparts + 1, dtype=np.uint32)
    cost = work.astype(np.float64) + ROW_CONSTANT
    total = float(cost.sum())
    acc = 0.0
    part = 1
    for g in range(n):
        if part >= nparts:
            break
        acc += float(cost[g])
        while part < nparts and acc >= total * part / nparts:
            bounds[part] = g + 1
            part += 1
    while part < nparts:
        bounds[part] = n
        part += 1
    bounds[nparts] = n
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
    torch.cuda.current_stream(device).synchronize()      # the library works on its own stream
    ctx.load_sketches_device(full.data_ptr(), d_off.data_ptr(), int(offsets.shape[0]) - 1)
    return full
