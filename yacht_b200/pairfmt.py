"""Pair-file lines and result digests of the train path.

The reference core writes one line per flagged ordered pair, ``i,j,jaccard,c_ij,c_ji`` with the three doubles at the
default ostream precision (src/cpp/main.cpp:296-305; identical to ``printf("%g")``).  The
digests are what tests/golden/config_digests.json stores for BASELINE.json's configurations -- produced there by the
UNMODIFIED reference binary -- so a run at full size can be checked without holding the reference's output.
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence

import numpy as np


def pair_lines(pairs: np.ndarray, sizes: np.ndarray) -> List[str]:
    """Lines of the <pass>_<thread>.txt files for `pairs` (fields i, j, count), sizes = sketch sizes."""
    i = pairs["i"].astype(np.int64)
    j = pairs["j"].astype(np.int64)
    m = pairs["count"].astype(np.int64)
    ni = np.asarray(sizes, dtype=np.int64)[i]
    nj = np.asarray(sizes, dtype=np.int64)[j]
    # the reference's expressions in double: 1.0*m/(ni+nj-m), 1.0*m/ni, 1.0*m/nj  (main.cpp:296-298)
    jac = m.astype(np.float64) / (ni + nj - m).astype(np.float64)
    cij = m.astype(np.float64) / ni.astype(np.float64)
    cji = m.astype(np.float64) / nj.astype(np.float64)
    return ["%d,%d,%g,%g,%g" % t for t in zip(i.tolist(), j.tolist(), jac.tolist(), cij.tolist(), cji.tolist())]


def digest_lines(lines: Sequence[str]) -> str:
    return hashlib.sha256("\n".join(sorted(lines)).encode()).hexdigest()


def digest_ij(pairs: np.ndarray) -> str:
    """sha256 of the (i, j) pairs as little-endian int32, sorted by (i, j)."""
    key = (pairs["i"].astype(np.int64) << 32) | pairs["j"].astype(np.int64)
    order = np.argsort(key, kind="stable")
    ij = np.stack([pairs["i"][order], pairs["j"][order]], axis=1).astype("<i4")
    return hashlib.sha256(np.ascontiguousarray(ij).tobytes()).hexdigest()


def digest_ids(ids: Sequence[int]) -> str:
    return hashlib.sha256("\n".join(str(int(g)) for g in ids).encode()).hexdigest()
