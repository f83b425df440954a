"""TEST INFRASTRUCTURE ONLY -- Python handle on oracle/sketch_oracle.c (FracMinHash sketching as sourmash publishes it).

``read_records`` is a deliberately plain line-by-line FASTA/FASTQ reader (independent of the product's reader in
yacht_b200/sketch.py); ``sketch_records`` / ``sketch_file`` give the sorted distinct hashes and their abundances, i.e. the
``mins`` / ``abundances`` arrays of the signature ``sourmash sketch dna -p k=K,scaled=S,abund`` writes.
Parity status: pinned by four counts of the reference's checked-in workbook + the hash function's KAT (header of
sketch_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes
import gzip
import os
from typing import Iterable, List, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libsketch_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle port` (or __graft_entry__.build())")
        lib = ctypes.CDLL(LIB_PATH)
        lib.so_murmur3_x64_128.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]
        lib.so_murmur3_verification.restype = ctypes.c_uint32
        lib.so_sketch_record.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64,
                                         ctypes.c_void_p, ctypes.c_uint64]
        lib.so_sketch_record.restype = ctypes.c_uint64
        _lib = lib
    return _lib


def max_hash_for_scaled(scaled: int) -> int:
    """sourmash: 0 -> 0, 1 -> 2^64 - 1, else round((2^64 - 1) / scaled) in double arithmetic, capped at 2^64 - 1."""
    if scaled == 0:
        return 0
    if scaled == 1:
        return 2 ** 64 - 1
    return min(int(round((2 ** 64 - 1) / scaled, 0)), 2 ** 64 - 1)


def murmur3_x64_128(data: bytes, seed: int = 42) -> Tuple[int, int]:
    out = (ctypes.c_uint64 * 2)()
    buf = ctypes.create_string_buffer(data, len(data))
    _load().so_murmur3_x64_128(buf, len(data), seed, out)
    return int(out[0]), int(out[1])


def murmur3_verification() -> int:
    return int(_load().so_murmur3_verification())


def read_records(path: str) -> List[Tuple[str, bytes]]:
    """[(name, sequence)] of a FASTA or FASTQ file (optionally gzip), line by line."""
    opener = gzip.open if path.endswith(".gz") else open
    records: List[Tuple[str, bytes]] = []
    with opener(path, "rb") as f:
        lines = [ln.rstrip(b"\r\n") for ln in f]
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(b">"):
            name = ln[1:].decode()
            i += 1
            parts = []
            while i < len(lines) and not lines[i].startswith(b">"):
                parts.append(lines[i])
                i += 1
            records.append((name, b"".join(parts)))
        elif ln.startswith(b"@"):
            records.append((ln[1:].decode(), lines[i + 1]))
            i += 4
        elif not ln:
            i += 1
        else:
            raise ValueError(f"{path}: line {i + 1} starts neither a FASTA nor a FASTQ record")
    return records


def sketch_records(seqs: Iterable[bytes], ksize: int, scaled: int, seed: int = 42) -> Tuple[np.ndarray, np.ndarray]:
    """(mins ascending, abundances) of all records together."""
    lib = _load()
    max_hash = max_hash_for_scaled(scaled)
    kept = []
    for seq in seqs:
        n = len(seq)
        if n < ksize:
            continue
        cap = n - ksize + 1
        out = np.empty(cap, dtype=np.uint64)
        buf = np.frombuffer(seq, dtype=np.uint8)
        got = lib.so_sketch_record(buf.ctypes.data, n, ksize, seed, max_hash, out.ctypes.data, cap)
        kept.append(out[:got].copy())
    if not kept:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    mins, ab = np.unique(np.concatenate(kept), return_counts=True)
    return mins.astype(np.uint64), ab.astype(np.uint32)


def sketch_file(path: str, ksize: int = 31, scaled: int = 1000, seed: int = 42) -> Tuple[np.ndarray, np.ndarray]:
    return sketch_records([s for _, s in read_records(path)], ksize, scaled, seed)
