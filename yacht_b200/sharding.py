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
    merged = torch.cat(parts).cpu().numpy().view(PAIR_DTYPE)
    # rank-ordered concatenation is deterministic; ranges interleave (the owner of row a also
    # reports (b, a)), so one final ordering by (i, j)
    return np.sort(merged, order=["i", "j"])
