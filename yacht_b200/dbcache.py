"""Packed reference-sketch cache: the flat uint64 hash array + CSR offsets of a trained
database, stored next to the reference's own intermediate files so that `yacht run` hands them to the device
without parsing ~66k JSON signature files again.

Nothing the reference writes is touched: the cache is a new sub-directory `ygpu_cache/` of
`<prefix>_intermediate_files/` (the directory make_training_data_from_sketches.py:112-155 creates and
run_YACHT.py:121 reads back through `_config.json`).  It is keyed by the manifest's md5sum column in row order
(row order = genome id on the device) and ignored -- then rebuilt -- when that list differs, when a file is
missing or has the wrong size, or when YACHT_DB_CACHE=0.

Layout:  meta.json {version, genomes, hashes, md5_digest}   hashes.npy (uint64[T])   offsets.npy (uint64[n+1])
"""
from __future__ import annotations

import hashlib
import json
import os
import tempfile
from typing import Optional, Sequence, Tuple

import numpy as np

VERSION = 1
DIRNAME = "ygpu_cache"


def enabled() -> bool:
    return os.environ.get("YACHT_DB_CACHE", "1") not in ("0", "false", "no")


def _digest(md5sums: Sequence[str]) -> str:
    h = hashlib.sha256()
    for m in md5sums:
        h.update(str(m).encode())
        h.update(b"\n")
    return h.hexdigest()


def cache_dir(path_to_genome_temp_dir: str) -> str:
    return os.path.join(path_to_genome_temp_dir, DIRNAME)


def load(path_to_genome_temp_dir: str, md5sums: Sequence[str]) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """(hashes, offsets) when a valid cache for exactly this genome list exists, else None.  `hashes` is a
    read-only memory map: the only pass over it is the copy to the device."""
    if not enabled():
        return None
    d = cache_dir(path_to_genome_temp_dir)
    try:
        with open(os.path.join(d, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("version") != VERSION or meta.get("genomes") != len(md5sums) or meta.get("md5_digest") != _digest(md5sums):
            return None
        offsets = np.load(os.path.join(d, "offsets.npy"))
        hashes = np.load(os.path.join(d, "hashes.npy"), mmap_mode="r")
        if offsets.dtype != np.uint64 or hashes.dtype != np.uint64 or offsets.shape != (len(md5sums) + 1,):
            return None
        if int(offsets[0]) != 0 or int(offsets[-1]) != hashes.shape[0] or meta.get("hashes") != hashes.shape[0]:
            return None
        if np.any(np.diff(offsets.astype(np.int64)) < 0):
            return None
        return hashes, offsets
    except (OSError, ValueError, KeyError, json.JSONDecodeError):
        return None


def store(path_to_genome_temp_dir: str, md5sums: Sequence[str], hashes: np.ndarray, offsets: np.ndarray) -> bool:
    """Write the cache atomically (temporary directory + rename of each file, meta.json last).  Failures are not
    errors: the cache is an optimisation."""
    if not enabled():
        return False
    d = cache_dir(path_to_genome_temp_dir)
    try:
        os.makedirs(d, exist_ok=True)
        meta_path = os.path.join(d, "meta.json")
        if os.path.exists(meta_path):
            os.remove(meta_path)                      # invalid while the arrays are being replaced
        for name, arr in (("hashes.npy", np.ascontiguousarray(hashes, dtype=np.uint64)),
                          ("offsets.npy", np.ascontiguousarray(offsets, dtype=np.uint64))):
            fd, tmp = tempfile.mkstemp(dir=d, suffix=".tmp")
            with os.fdopen(fd, "wb") as f:
                np.save(f, arr)
            os.replace(tmp, os.path.join(d, name))
        fd, tmp = tempfile.mkstemp(dir=d, suffix=".tmp")
        with os.fdopen(fd, "w") as f:
            json.dump({"version": VERSION, "genomes": len(md5sums), "hashes": int(np.asarray(offsets)[-1]) if len(offsets) else 0,
                       "md5_digest": _digest(md5sums)}, f)
        os.replace(tmp, meta_path)
        return True
    except OSError:
        return False
