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
    """Equal contiguous slices of a flat array (the last one may be short): slice r = [b[r], b[r+1])."""
    per = (int(total) + world - 1) // world if world else 0
    return np.minimum(np.arange(world + 1, dtype=np.int64) * per, int(total))


def split_rows_by_size(offsets: np.ndarray, nparts: int) -> np.ndarray:
    """Contiguous genome ranges holding nearly equal numbers of sketch hashes: the residency of the sharded train step
    (rank r holds the sketches of genomes [b[r], b[r+1]) and counts / flags those query rows)."""
    sizes = np.diff(np.asarray(offsets).astype(np.int64)).astype(np.float64)
    return split_rows_by_work(sizes, nparts)


# ---- host mirror of the sharded index build's exchange plan (csrc/index_msd.cu: k2s_prep) -------------------------------
# Every rank histograms the leading hash bits of ITS sketches; the histograms are all-gathered and every rank derives the
# same plan: which rank owns which level-1 digit, and where in the owner's buffer the words of (source rank, digit) lie.
# The level-1 scatter then stores every word straight into its owner's buffer.  The device computes this in one CTA; this
# numpy version documents the rule and is what the CPU tests (gloo, world size 2) check.
def exchange_plan(hist_all: np.ndarray) -> dict:
    """hist_all[r, d] = hashes of rank r's sketches with leading digit d.  Returns
    owner[d]; start[r_src, d] = first slot, in owner[d]'s buffer, of the words (r_src, d); count[o] = words rank o owns."""
    hist_all = np.asarray(hist_all, dtype=np.int64)
    nranks, nb = hist_all.shape
    g = hist_all.sum(axis=0)
    total = int(g.sum())
    ex = np.concatenate([[0], np.cumsum(g)[:-1]])                      # hashes before digit d
    owner = np.minimum(nranks - 1, (ex * nranks) // max(total, 1)) if total else np.zeros(nb, dtype=np.int64)
    own_start = np.zeros(nranks, dtype=np.int64)
    for d in range(nb):
        if d == 0 or owner[d] != owner[d - 1]:
            own_start[owner[d]] = ex[d]
    pre = np.concatenate([np.zeros((1, nb), dtype=np.int64), np.cumsum(hist_all, axis=0)[:-1]], axis=0)   # words of lower ranks
    start = (ex - own_start[owner])[None, :] + pre
    count = np.array([int(g[owner == o].sum()) for o in range(nranks)], dtype=np.int64)
    return dict(owner=owner.astype(np.int64), start=start, count=count, total=total)


def exchange_capacity(total: int, nranks: int, nb: int) -> int:
    """Words a rank's exchange buffer holds (ygpu_train_step_sharded): its share plus the digit granularity of the split."""
    return total // nranks + 2 * (total // max(nb, 1)) + 3 * 4096


# ---- hash-range residency (ygpu_load_sketches_hashrange): cut every sketch at the same hash values ----------------------------
def hash_cuts(max_hash: int, nranks: int) -> np.ndarray:
    """nranks + 1 cut values: rank r holds the hashes in [cuts[r], cuts[r + 1]).  FracMinHash hashes are uniform below max_hash
    by construction, so equal-width ranges hold equal shares."""
    top = int(max_hash) + 1
    return np.array([top * r // nranks for r in range(nranks)] + [2 ** 64 - 1], dtype=np.uint64)


def hashrange_share(hashes: np.ndarray, offsets: np.ndarray, lo: int, hi: int, last: bool = False):
    """(part_hashes, part_offsets) of the hashes h with lo <= h < hi (<= when `last`), sketch by sketch, order preserved."""
    hashes = np.asarray(hashes, dtype=np.uint64)
    keep = hashes >= np.uint64(lo)
    keep &= (hashes <= np.uint64(hi)) if last else (hashes < np.uint64(hi))
    c = np.zeros(hashes.shape[0] + 1, dtype=np.uint64)
    np.cumsum(keep, out=c[1:])
    return hashes[keep], c[np.asarray(offsets).astype(np.int64)]
