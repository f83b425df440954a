"""CPU suite of the sketching row: the oracle against its pins, the device hash code compiled for the host
against the oracle, the host-side reader and signature writer.  No GPU needed."""
import ctypes
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import sketch_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "sketch_golden.json")


def test_murmur3_matches_the_published_verification_value():
    # SMHasher's VerificationTest value for MurmurHash3_x64_128
    assert so.murmur3_verification() == 0x6384BA69


def test_max_hash_for_scaled():
    assert so.max_hash_for_scaled(1000) == 18446744073709552        # "max_hash" of every signature the reference ships
    assert so.max_hash_for_scaled(1) == 2 ** 64 - 1
    assert so.max_hash_for_scaled(0) == 0


def test_golden_was_pinned_by_the_reference_workbook():
    """tests/golden/make_sketch_golden.py sketched the reference's demo genomes with the oracle and compared with the counts
    in the reference's checked-in workbook; the committed file records both."""
    g = json.load(open(GOLD))
    assert g["workbook"]["GCF_018918235.1"] == {"num_unique_kmers_in_genome_sketch": 2319, "num_total_kmers_in_genome_sketch": 2323}
    assert g["workbook"]["GCF_018918045.1"] == {"num_unique_kmers_in_genome_sketch": 2452, "num_total_kmers_in_genome_sketch": 2453}
    assert len(g["workbook"]) == 5
    for name, want in g["workbook"].items():
        got = g["demo_genomes"][name]
        assert got["n_unique"] == want["num_unique_kmers_in_genome_sketch"]
        assert got["n_total"] == want["num_total_kmers_in_genome_sketch"]
    assert len(g["demo_genomes"]) == 15


@pytest.mark.skipif(not os.path.isdir("/root/reference/demo/ref_genomes"), reason="reference tree not present")
def test_oracle_reproduces_the_workbook_counts_live():
    for name, (nu, nt) in {"GCF_018918235.1": (2319, 2323), "GCF_018918045.1": (2452, 2453)}.items():
        mins, ab = so.sketch_file(f"/root/reference/demo/ref_genomes/{name}_genomic.fna.gz", 31, 1000)
        assert (len(mins), int(ab.sum())) == (nu, nt)


def _random_sequence(rng, n, p_bad=0.01, p_lower=0.2):
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    lower = rng.random(n) < p_lower
    seq = np.where(lower, seq | 0x20, seq).astype(np.uint8)
    bad = rng.random(n) < p_bad
    seq = np.where(bad, rng.choice(np.frombuffer(b"NnRYKM-*\n.", dtype=np.uint8), size=n), seq).astype(np.uint8)
    return seq.tobytes()


@pytest.fixture(scope="module")
def host_hash(tmp_path_factory):
    """yacht_b200/csrc/sketch_hash.cuh compiled for the host (the functions the kernel calls)."""
    out = tmp_path_factory.mktemp("hh") / "libsketch_hash_host.so"
    src = os.path.join(HERE, "harness", "sketch_hash_host.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", str(out), src], check=True)
    lib = ctypes.CDLL(str(out))
    lib.hh_hash_windows.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]
    lib.hh_hash_windows.restype = ctypes.c_uint64
    return lib


@pytest.mark.parametrize("k", [1, 2, 7, 8, 9, 15, 16, 17, 21, 24, 31, 32, 33, 47, 48, 51, 64, 100])
def test_device_hash_code_matches_oracle(host_hash, k):
    rng = np.random.default_rng(k)
    seq = _random_sequence(rng, 20000)
    lib = so._load()
    n = len(seq)
    buf = np.frombuffer(seq, dtype=np.uint8)
    exp = np.empty(n, dtype=np.uint64)
    n_exp = lib.so_sketch_record(buf.ctypes.data, n, k, 42, 2 ** 64 - 1, exp.ctypes.data, n)     # max_hash = all ones: every hash kept, in order
    got = np.empty(n, dtype=np.uint64)
    n_got = host_hash.hh_hash_windows(buf.ctypes.data, n, k, 42, got.ctypes.data)
    assert n_got == n_exp and n_exp > 1000
    assert np.array_equal(got[:n_got], exp[:n_exp])


def test_device_hash_code_palindromes_and_seed(host_hash):
    lib = so._load()
    for seq in [b"ACGT" * 10, b"AATT" * 8, b"GAATTC" * 6, b"A" * 40, b"T" * 40, b"acgtnACGT" * 9]:
        for k, seed in [(4, 42), (6, 42), (8, 7), (16, 0), (31, 42)]:
            n = len(seq)
            buf = np.frombuffer(seq, dtype=np.uint8)
            exp = np.empty(n, dtype=np.uint64)
            got = np.empty(n, dtype=np.uint64)
            ne = lib.so_sketch_record(buf.ctypes.data, n, k, seed, 2 ** 64 - 1, exp.ctypes.data, n)
            ng = host_hash.hh_hash_windows(buf.ctypes.data, n, k, seed, got.ctypes.data)
            assert ne == ng and np.array_equal(got[:ng], exp[:ne]), (seq, k, seed)


def test_reader_and_oracle_reader_agree(tmp_path):
    from yacht_b200 import sketch
    rng = np.random.default_rng(3)
    recs = [(f"contig_{i} some description", _random_sequence(rng, int(rng.integers(1, 3000)), p_bad=0.02).replace(b"\n", b"N").replace(b"*", b"N"))
            for i in range(12)]
    fa = tmp_path / "g.fna"
    with open(fa, "wb") as f:
        for name, seq in recs:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + b"\n")
    with open(fa, "rb") as f, gzip.open(str(fa) + ".gz", "wb") as g:
        g.write(f.read())
    fq = tmp_path / "r.fq"
    with open(fq, "wb") as f:
        for name, seq in recs:
            f.write(b"@" + name.encode() + b"\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n")
    for path in (str(fa), str(fa) + ".gz", str(fq)):
        a = so.read_records(path)
        b = sketch.read_records(path)
        assert [(n, s) for n, s in a] == [(n, s) for n, s in b], path
        assert [(n, s) for n, s in a] == recs, path
