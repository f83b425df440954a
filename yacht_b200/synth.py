"""Seeded synthetic FracMinHash reference databases and samples.

Hashes are uniform in ``[0, 18446744073709552)`` (scaled = 1000).  Genome sizes follow
Normal(5000, 1500) clipped to >= 50.  With probability 0.12 a genome starts a cluster of 2-8
members; each further member keeps every hash of the cluster's first genome with probability
``ANI**31`` (ANI drawn from {0.90, 0.95, 0.97, 0.99, 0.999}) and is topped up with fresh random
hashes.  The genome order is shuffled at the end.  Every sketch is sorted ascending and unique,
as sourmash writes them.

This is input generation for tests and bench.py, not part of the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

MAX_HASH = 18446744073709552
ANI_CHOICES = (0.90, 0.95, 0.97, 0.99, 0.999)


@dataclass
class SketchDB:
    """Flat layout used at the C-ABI: ``hashes[offsets[g]:offsets[g+1]]`` is sketch ``g``."""
    hashes: np.ndarray    # uint64 [T]
    offsets: np.ndarray   # uint64 [N+1]
    cluster: np.ndarray   # int64 [N] cluster id of each genome (after the shuffle), -1 = singleton

    @property
    def n(self) -> int:
        return int(self.offsets.shape[0] - 1)

    @property
    def sizes(self) -> np.ndarray:
        return np.diff(self.offsets.astype(np.int64))

    def sketch(self, g: int) -> np.ndarray:
        return self.hashes[int(self.offsets[g]):int(self.offsets[g + 1])]

    def subset(self, ids: Sequence[int]) -> "SketchDB":
        parts = [self.sketch(int(g)) for g in ids]
        return from_sketches(parts, self.cluster[np.asarray(ids, dtype=np.int64)] if len(ids) else None)


def from_sketches(parts: Sequence[np.ndarray], cluster: Optional[np.ndarray] = None) -> SketchDB:
    sizes = np.array([len(p) for p in parts], dtype=np.uint64)
    offsets = np.zeros(len(parts) + 1, dtype=np.uint64)
    np.cumsum(sizes, out=offsets[1:])
    hashes = (np.concatenate([np.asarray(p, dtype=np.uint64) for p in parts])
              if len(parts) and int(offsets[-1]) > 0 else np.zeros(0, dtype=np.uint64))
    if cluster is None:
        cluster = np.full(len(parts), -1, dtype=np.int64)
    return SketchDB(hashes=hashes, offsets=offsets, cluster=np.asarray(cluster, dtype=np.int64))


def make_reference_db(n: int, seed: int, mean_size: float = 5000.0, sd_size: float = 1500.0,
                      min_size: int = 50, p_cluster: float = 0.12, cluster_lo: int = 2,
                      cluster_hi: int = 8, ksize: int = 31, shuffle: bool = True,
                      zipf_clusters: bool = False, zipf_a: float = 1.3, zipf_cap: int = 20000,
                      core_hashes: int = 0, core_lo: float = 0.01, core_hi: float = 0.10) -> SketchDB:
    """Planted-cluster database.  ``zipf_clusters`` / ``core_hashes`` give the skewed
    "full-GTDB shape" of config 4 (Zipf(1.3) cluster sizes capped at 20 000 and a conserved core of
    hashes each present in 1-10 % of all genomes)."""
    rng = np.random.default_rng(seed)
    sizes = np.clip(np.rint(rng.normal(mean_size, sd_size, size=n)), min_size, None).astype(np.int64)

    # cluster plan: parent[g] = first genome of the cluster, or -1
    parent = np.full(n, -1, dtype=np.int64)
    cluster = np.full(n, -1, dtype=np.int64)
    g = 0
    cid = 0
    starts = rng.random(n) < p_cluster
    while g < n:
        if starts[g] and g + 1 < n:
            if zipf_clusters:
                s = int(min(max(2, rng.zipf(zipf_a) + 1), zipf_cap))
            else:
                s = int(rng.integers(cluster_lo, cluster_hi + 1))
            s = min(s, n - g)
            cluster[g:g + s] = cid
            parent[g + 1:g + s] = g
            cid += 1
            g += s
        else:
            g += 1

    parts: List[Optional[np.ndarray]] = [None] * n
    roots = np.flatnonzero(parent < 0)
    flat = rng.integers(0, MAX_HASH, size=int(sizes[roots].sum()), dtype=np.uint64)
    pos = 0
    for r in roots:
        parts[r] = np.unique(flat[pos:pos + sizes[r]])
        pos += int(sizes[r])
    del flat
    children = np.flatnonzero(parent >= 0)
    ani = rng.choice(np.array(ANI_CHOICES), size=len(children))
    for c, a in zip(children, ani):
        ph = parts[int(parent[c])]
        kept = ph[rng.random(ph.shape[0]) < a ** ksize]
        n_fresh = max(0, int(sizes[c]) - kept.shape[0])
        fresh = rng.integers(0, MAX_HASH, size=n_fresh, dtype=np.uint64)
        parts[c] = np.unique(np.concatenate([kept, fresh]))

    if core_hashes > 0:
        core = rng.integers(0, MAX_HASH, size=core_hashes, dtype=np.uint64)
        frac = rng.uniform(core_lo, core_hi, size=core_hashes)
        add: List[List[int]] = [[] for _ in range(n)]
        for h, f in zip(core, frac):
            members = np.flatnonzero(rng.random(n) < f)
            for m in members:
                add[m].append(int(h))
        for m in range(n):
            if add[m]:
                parts[m] = np.unique(np.concatenate([parts[m], np.array(add[m], dtype=np.uint64)]))

    order = rng.permutation(n) if shuffle else np.arange(n)
    parts = [parts[int(o)] for o in order]
    return from_sketches(parts, cluster[order])


def make_skewed_db(n: int, seed: int, mean_size: float = 5000.0, sd_size: float = 1500.0, min_size: int = 50,
                   p_cluster: float = 0.12, zipf_a: float = 1.3, zipf_cap: int = 20000, core_hashes: int = 200,
                   core_lo: float = 0.01, core_hi: float = 0.10, ksize: int = 31) -> SketchDB:
    """The "full-GTDB shape" of BASELINE.json configs[3] at full size (Zipf(1.3) cluster sizes capped at
    `zipf_cap`, conserved core hashes each present in 1-10 % of all genomes), generated cluster by cluster with numpy
    (make_reference_db(zipf_clusters=True) draws child by child in Python: 11 minutes at 400 000 genomes).  Same model,
    its own random stream; the seeded golden cases keep using make_reference_db."""
    rng = np.random.default_rng(seed)
    sizes = np.clip(np.rint(rng.normal(mean_size, sd_size, size=n)), min_size, None).astype(np.int64)
    cluster = np.full(n, -1, dtype=np.int64)
    # every sketch is built as a row of a (genome, slot) table of width max size, then masked down to its own size
    blocks: List[np.ndarray] = []      # per genome: sorted unique hashes
    g = 0
    cid = 0
    starts = rng.random(n) < p_cluster
    sketches: List[Optional[np.ndarray]] = [None] * n
    while g < n:
        if starts[g] and g + 1 < n:
            s = int(min(max(2, rng.zipf(zipf_a) + 1), zipf_cap, n - g))
            cluster[g:g + s] = cid
            cid += 1
            parent = np.unique(rng.integers(0, MAX_HASH, size=int(sizes[g]), dtype=np.uint64))
            sketches[g] = parent
            m = s - 1
            ani = rng.choice(np.array(ANI_CHOICES), size=m) ** ksize
            for c0 in range(0, m, 2048):                      # chunks of children: the keep matrix stays below ~100 MB
                c1 = min(m, c0 + 2048)
                keep = rng.random((c1 - c0, parent.shape[0])) < ani[c0:c1, None]
                for k in range(c1 - c0):
                    child = g + 1 + c0 + k
                    kept = parent[keep[k]]
                    n_fresh = max(0, int(sizes[child]) - kept.shape[0])
                    fresh = rng.integers(0, MAX_HASH, size=n_fresh, dtype=np.uint64)
                    merged = np.concatenate([kept, fresh])
                    merged.sort()
                    sketches[child] = merged                  # (a 55-bit collision between kept and fresh hashes has probability ~1e-9 per sketch)
            g += s
        else:
            sketches[g] = np.unique(rng.integers(0, MAX_HASH, size=int(sizes[g]), dtype=np.uint64))
            g += 1
    if core_hashes > 0:
        core = np.sort(rng.integers(0, MAX_HASH, size=core_hashes, dtype=np.uint64))
        frac = rng.uniform(core_lo, core_hi, size=core_hashes)
        member = rng.random((core_hashes, n)) < frac[:, None]          # 200 x 400 000 booleans
        for m in np.flatnonzero(member.any(axis=0)):
            sketches[m] = np.unique(np.concatenate([sketches[m], core[member[:, m]]]))
    order = rng.permutation(n)
    return from_sketches([sketches[int(o)] for o in order], cluster[order])


def make_sample(db: SketchDB, seed: int, n_present: int, total_hashes: int,
                cov_lo: float = 0.01, cov_hi: float = 1.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Config-5 style metagenome sample: the union of ``n_present`` random reference genomes, each
    subsampled at a coverage drawn from U(cov_lo, cov_hi), plus random noise hashes up to
    ``total_hashes``.  Returns (sorted unique hashes, present genome ids, their coverages)."""
    rng = np.random.default_rng(seed)
    n_present = min(n_present, db.n)
    present = np.sort(rng.choice(db.n, size=n_present, replace=False))
    cov = rng.uniform(cov_lo, cov_hi, size=n_present)
    chunks = []
    for g, c in zip(present, cov):
        sk = db.sketch(int(g))
        chunks.append(sk[rng.random(sk.shape[0]) < c])
    have = int(sum(len(c) for c in chunks))
    if total_hashes > have:
        chunks.append(rng.integers(0, MAX_HASH, size=total_hashes - have, dtype=np.uint64))
    sample = np.unique(np.concatenate(chunks)) if chunks else np.zeros(0, dtype=np.uint64)
    return sample, present, cov


def workload_counts(db: SketchDB) -> dict:
    """Implementation-independent T, U, U2, P, W of SURVEY.md section 8 (sort + run-length on the
    host).  Used by bench.py to compute the algorithmic bytes of the pairwise-count kernel."""
    T = int(db.offsets[-1])
    if T == 0:
        return dict(T=0, U=0, U2=0, P=0, W=0)
    s = np.sort(db.hashes)
    head = np.empty(T, dtype=bool)
    head[0] = True
    np.not_equal(s[1:], s[:-1], out=head[1:])
    starts = np.flatnonzero(head)
    lens = np.diff(np.append(starts, T))
    shared = lens[lens >= 2].astype(np.int64)
    return dict(T=T, U=int(len(lens)), U2=int(len(shared)), P=int(shared.sum()),
                W=int((shared * shared).sum()))
