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
    lib.hh_hash_windows_packed.argtypes = lib.hh_hash_windows.argtypes
    lib.hh_hash_windows_packed.restype = ctypes.c_uint64
    lib.hh_tile_walk_packed.argtypes = lib.hh_hash_windows.argtypes
    lib.hh_tile_walk_packed.restype = ctypes.c_uint64
    for fn in (lib.hh_hash_windows_packed2, lib.hh_tile_walk_packed2):
        fn.argtypes = lib.hh_hash_windows.argtypes
        fn.restype = ctypes.c_uint64
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


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_device_packed_hash_code_matches_oracle(host_hash, k):
    """the k <= 32 path of the kernel: window = one 64-bit word of 2-bit codes"""
    rng = np.random.default_rng(1000 + k)
    seq = _random_sequence(rng, 6000) + b"ACGT" * 12 + b"AATT" * 10 + b"A" * 40 + b"GAATTC" * 8
    lib = so._load()
    n = len(seq)
    buf = np.frombuffer(seq, dtype=np.uint8)
    exp = np.empty(n, dtype=np.uint64)
    got = np.empty(n, dtype=np.uint64)
    for seed in (42, 0, 0xFFFFFFFF):
        n_exp = lib.so_sketch_record(buf.ctypes.data, n, k, seed, 2 ** 64 - 1, exp.ctypes.data, n)
        n_got = host_hash.hh_hash_windows_packed(buf.ctypes.data, n, k, seed, got.ctypes.data)
        assert n_got == n_exp and n_exp > 100
        assert np.array_equal(got[:n_got], exp[:n_exp]), (k, seed)


@pytest.mark.parametrize("k,n", [(31, 20000), (32, 9000), (21, 4096 * 2), (1, 4097), (16, 4095), (31, 4096 + 30), (31, 31), (31, 30), (7, 12345)])
def test_packed_tile_walk_matches_oracle(host_hash, k, n):
    """the packed kernel's shared-memory layout and per-thread window extraction, walked on the host with the kernel's own helpers"""
    rng = np.random.default_rng(k * 100003 + n)
    for p_bad in (0.0, 0.01, 0.2):
        seq = _random_sequence(rng, n, p_bad=p_bad)
        lib = so._load()
        buf = np.frombuffer(seq, dtype=np.uint8)
        exp = np.empty(n + 1, dtype=np.uint64)
        got = np.empty(n + 1, dtype=np.uint64)
        n_exp = lib.so_sketch_record(buf.ctypes.data, n, k, 42, 2 ** 64 - 1, exp.ctypes.data, n)
        n_got = host_hash.hh_tile_walk_packed(buf.ctypes.data, n, k, 42, got.ctypes.data)
        assert n_got == n_exp, (k, n, p_bad)
        assert np.array_equal(got[:n_got], exp[:n_exp]), (k, n, p_bad)


@pytest.mark.parametrize("k", list(range(33, 65)))
def test_device_packed2_hash_code_matches_oracle(host_hash, k):
    """the 33 <= k <= 64 path of the kernel: window = two 64-bit words of 2-bit codes"""
    rng = np.random.default_rng(2000 + k)
    seq = _random_sequence(rng, 8000, p_bad=0.004) + b"ACGT" * 40 + b"AATT" * 40 + b"A" * 100 + b"GAATTC" * 30
    lib = so._load()
    n = len(seq)
    buf = np.frombuffer(seq, dtype=np.uint8)
    exp = np.empty(n, dtype=np.uint64)
    got = np.empty(n, dtype=np.uint64)
    for seed in (42, 7):
        n_exp = lib.so_sketch_record(buf.ctypes.data, n, k, seed, 2 ** 64 - 1, exp.ctypes.data, n)
        n_got = host_hash.hh_hash_windows_packed2(buf.ctypes.data, n, k, seed, got.ctypes.data)
        assert n_got == n_exp and n_exp > 100
        assert np.array_equal(got[:n_got], exp[:n_exp]), (k, seed)


@pytest.mark.parametrize("k,n", [(51, 20000), (64, 9000), (33, 4096 * 2), (48, 4097), (47, 4095), (51, 4096 + 50), (51, 51), (51, 50), (63, 12345)])
def test_packed2_tile_walk_matches_oracle(host_hash, k, n):
    rng = np.random.default_rng(k * 100003 + n)
    for p_bad in (0.0, 0.005, 0.1):
        seq = _random_sequence(rng, n, p_bad=p_bad)
        lib = so._load()
        buf = np.frombuffer(seq, dtype=np.uint8)
        exp = np.empty(n + 1, dtype=np.uint64)
        got = np.empty(n + 1, dtype=np.uint64)
        n_exp = lib.so_sketch_record(buf.ctypes.data, n, k, 42, 2 ** 64 - 1, exp.ctypes.data, n)
        n_got = host_hash.hh_tile_walk_packed2(buf.ctypes.data, n, k, 42, got.ctypes.data)
        assert n_got == n_exp, (k, n, p_bad)
        assert np.array_equal(got[:n_got], exp[:n_exp]), (k, n, p_bad)


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


class _OracleBackedContext:
    """Stands in for the GPU context so that the HOST logic of yacht_b200/sketch.py (record separators, batching, merging of
    partial sketches, writers) can be exercised on CPU; the device call itself is covered by tests/test_sketch_gpu.py."""
    def __init__(self):
        self.calls = []

    def sketch_sequences(self, bases, offsets, ksize, max_hash, seed=42):
        bases = bytes(bases)
        self.calls.append(len(bases))
        hs, ab, off = [], [], [0]
        lib = so._load()
        for s in range(len(offsets) - 1):
            chunk = bases[int(offsets[s]):int(offsets[s + 1])]
            n = len(chunk)
            kept = np.empty(max(n, 1), dtype=np.uint64)
            buf = np.frombuffer(chunk, dtype=np.uint8) if n else np.zeros(1, np.uint8)
            got = lib.so_sketch_record(buf.ctypes.data, n, ksize, seed, max_hash, kept.ctypes.data, n) if n >= ksize else 0
            m, a = np.unique(kept[:got], return_counts=True)
            hs.append(m.astype(np.uint64)); ab.append(a.astype(np.uint32)); off.append(off[-1] + len(m))
        return np.concatenate(hs), np.concatenate(ab), np.array(off, np.uint64), 0


def test_host_batching_and_merging(monkeypatch):
    from yacht_b200 import sketch
    fake = _OracleBackedContext()
    monkeypatch.setattr(sketch, "_context", lambda: fake)
    monkeypatch.setattr(sketch, "BATCH_BASES", 50_000)
    rng = np.random.default_rng(23)
    unit = _random_sequence(rng, 30_000, p_bad=0.0)
    groups = [[unit, unit[:10_000], _random_sequence(rng, 45_000), unit], [_random_sequence(rng, 5_000)], [], [_random_sequence(rng, 70_000)],
              [b"ACGT"], [_random_sequence(rng, 20_000), _random_sequence(rng, 20_000), _random_sequence(rng, 20_000)]]
    got = sketch.sketch_record_groups(groups, 31, 50)
    assert len(fake.calls) >= 4 and len(got) == len(groups)
    for g, (mins, ab) in zip(groups, got):
        em, ea = so.sketch_records([s.replace(b"\n", b"N") for s in g], 31, 50)
        assert np.array_equal(mins, em) and np.array_equal(ab, ea)


def test_sketch_ref_and_sample_writers(tmp_path, monkeypatch):
    import argparse
    from yacht_b200 import sigio, sketch, sketch_ref_genomes, sketch_sample
    monkeypatch.setattr(sketch, "_context", lambda: _OracleBackedContext())
    rng = np.random.default_rng(5)
    folder = tmp_path / "g"
    folder.mkdir()
    recs = {}
    for g, ext in enumerate([".fna.gz", ".fa", ".fasta"]):
        rs = [(f"c{g}_{c}", _random_sequence(rng, 30_000).replace(b"\n", b"N").replace(b"*", b"N")) for c in range(2)]
        path = folder / f"G{g}{ext}"
        opener = gzip.open if ext.endswith(".gz") else open
        with opener(path, "wb") as f:
            for n, s in rs:
                f.write(b">" + n.encode() + b"\n" + s + b"\n")
        recs[f"G{g}"] = rs
    out = tmp_path / "ref.sig.zip"
    monkeypatch.setattr(sketch, "CHUNK_FILE_BYTES", 40_000)          # forces several read-ahead chunks
    sketch_ref_genomes.main(argparse.Namespace(infile=str(folder), kmer=31, scaled=100, outfile=str(out)))
    sigs = {s.name: s for s in sigio.read_sig_zip(str(out))}
    assert sorted(sigs) == sorted(recs)
    for name, rs in recs.items():
        em, ea = so.sketch_records([s for _, s in rs], 31, 100)
        assert np.array_equal(np.asarray(sigs[name].mins, np.uint64), em)
        assert np.array_equal(np.asarray(sigs[name].abundances, np.uint32), ea)
        assert sigs[name].scaled == 100 and sigs[name].ksize == 31
    # the zip a `yacht train` run starts from: loadable per k-mer size like any sourmash database
    fq = tmp_path / "s.fq"
    with open(fq, "wb") as f:
        for n, s in recs["G1"]:
            f.write(b"@" + n.encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    out2 = tmp_path / "sample.sig.zip"
    sketch_sample.main(argparse.Namespace(infile=[str(fq)], kmer=31, scaled=100, outfile=str(out2)))
    sig = sigio.load_signature_with_ksize(str(out2), 31)
    em, ea = so.sketch_records([s for _, s in recs["G1"]], 31, 100)
    assert np.array_equal(np.asarray(sig.mins, np.uint64), em) and np.array_equal(np.asarray(sig.abundances, np.uint32), ea)
    with pytest.raises(ValueError):
        sketch_sample.main(argparse.Namespace(infile=["a", "b", "c"], kmer=31, scaled=100, outfile=str(out2)))
