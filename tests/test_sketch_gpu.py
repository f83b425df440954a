"""GPU parity of the sketching row: ygpu_sketch_sequences and the `yacht sketch` mirrors against the CPU
oracle (oracle/sketch_oracle.c, pinned by the reference's workbook counts and the hash KAT -- tests/test_sketch_oracle.py).
Bar: bit-exact (hashes, abundances, sketch boundaries).  All device calls go through the C ABI."""
import gzip
import os

import numpy as np
import pytest

from oracle import sketch_oracle as so

pytestmark = pytest.mark.gpu


def _random_sequence(rng, n, p_bad=0.003, p_lower=0.2):
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    seq = np.where(rng.random(n) < p_lower, seq | 0x20, seq).astype(np.uint8)
    bad = rng.random(n) < p_bad
    return np.where(bad, rng.choice(np.frombuffer(b"NnRYKM-*.", dtype=np.uint8), size=n), seq).astype(np.uint8)


def _expect(bases: np.ndarray, offsets, k, scaled, seed=42):
    hs, ab, off = [], [], [0]
    for s in range(len(offsets) - 1):
        m, a = so.sketch_records([bases[int(offsets[s]):int(offsets[s + 1])].tobytes()], k, scaled, seed)
        hs.append(m)
        ab.append(a)
        off.append(off[-1] + len(m))
    return np.concatenate(hs) if hs else np.zeros(0, np.uint64), np.concatenate(ab) if ab else np.zeros(0, np.uint32), np.array(off, np.uint64)


def _check(ctx, bases, offsets, k, scaled, seed=42):
    got_h, got_a, got_o, n_kmers = ctx.sketch_sequences(bases, offsets, k, so.max_hash_for_scaled(scaled), seed)
    exp_h, exp_a, exp_o = _expect(bases, offsets, k, scaled, seed)
    assert np.array_equal(got_o, exp_o), (k, scaled)
    assert np.array_equal(got_h, exp_h), (k, scaled)
    assert np.array_equal(got_a, exp_a), (k, scaled)
    return n_kmers


@pytest.mark.parametrize("k,scaled", [(31, 1000), (21, 100), (51, 1000), (16, 50), (32, 10), (7, 1), (1, 1), (100, 20)])
def test_one_sketch_matches_oracle(gpu_ctx, k, scaled):
    rng = np.random.default_rng(100 + k)
    n = 700_000 if scaled >= 50 else 60_000
    bases = _random_sequence(rng, n)
    n_kmers = _check(gpu_ctx, bases, [0, n], k, scaled)
    assert 0 < n_kmers <= n - k + 1


@pytest.mark.parametrize("k,scaled", [(31, 100), (21, 10), (32, 50), (5, 1), (51, 100), (33, 10), (64, 20), (48, 50)])
def test_bytewise_kernel_forced_for_small_k(gpu_ctx, k, scaled):
    """k <= 64 normally takes the packed-word kernels; the byte-wise one (any k) must agree with them and with the oracle"""
    rng = np.random.default_rng(300 + k)
    n = 200_000 if scaled >= 10 else 30_000
    bases = _random_sequence(rng, n, p_bad=0.01)
    offsets = [0, 1234, 1234, 77_777 if n > 77_777 else 5000, n]
    packed = gpu_ctx.sketch_sequences(bases, offsets, k, so.max_hash_for_scaled(scaled))
    gpu_ctx.set_option("sketch_kernel", 2)
    try:
        bytewise = gpu_ctx.sketch_sequences(bases, offsets, k, so.max_hash_for_scaled(scaled))
        _check(gpu_ctx, bases, offsets, k, scaled)
    finally:
        gpu_ctx.set_option("sketch_kernel", 0)
    for a, b in zip(packed, bytewise):
        assert np.array_equal(a, b)
    _check(gpu_ctx, bases, offsets, k, scaled)


def test_repeats_give_abundances(gpu_ctx):
    rng = np.random.default_rng(7)
    unit = _random_sequence(rng, 50_000, p_bad=0.0)
    bases = np.concatenate([unit, np.frombuffer(b"N", np.uint8), unit, np.frombuffer(b"\n", np.uint8), unit[:20_000]])
    got_h, got_a, got_o, _ = gpu_ctx.sketch_sequences(bases, [0, len(bases)], 31, so.max_hash_for_scaled(100))
    assert got_a.max() == 3 and got_a.min() >= 2
    _check(gpu_ctx, bases, [0, len(bases)], 31, 100)


def test_many_sketches_with_empty_and_short_ranges(gpu_ctx):
    rng = np.random.default_rng(11)
    lens = [0, 5, 30, 31, 32, 4095, 4096, 4097, 0, 12_345, 1, 250_000, 8192, 31, 0]
    parts, offsets = [], [0]
    for ln in lens:
        seq = _random_sequence(rng, ln)
        if ln:
            seq[-1] = ord("\n")                   # what the host puts behind every record
        parts.append(seq)
        offsets.append(offsets[-1] + ln)
    bases = np.concatenate(parts)
    _check(gpu_ctx, bases, offsets, 31, 20)
    _check(gpu_ctx, bases, offsets, 21, 1)
    _check(gpu_ctx, bases, offsets, 51, 10)
    _check(gpu_ctx, bases, offsets, 64, 1)


def test_windows_do_not_cross_sketch_boundaries(gpu_ctx):
    # no separator between the ranges: a window that starts in one range and ends in the next belongs to neither
    rng = np.random.default_rng(13)
    bases = _random_sequence(rng, 40_000, p_bad=0.0)
    offsets = [0, 1000, 1010, 20_000, 40_000]
    _check(gpu_ctx, bases, offsets, 31, 1)


@pytest.mark.parametrize("n", [4096 * 3, 4096 * 3 + 1, 4096 * 3 - 1, 4096 * 2 + 30, 31, 30])
def test_tile_edges(gpu_ctx, n):
    rng = np.random.default_rng(n)
    bases = _random_sequence(rng, n, p_bad=0.0)
    _check(gpu_ctx, bases, [0, n], 31, 1)


@pytest.mark.parametrize("n", [4096 * 3, 4096 * 3 + 1, 4096 * 3 - 1, 4096 * 2 + 50, 4096 * 2 + 63, 51, 50, 64])
def test_tile_edges_two_word_windows(gpu_ctx, n):
    rng = np.random.default_rng(n + 7)
    bases = _random_sequence(rng, n, p_bad=0.0)
    _check(gpu_ctx, bases, [0, n], 51, 1)
    if n >= 64:
        _check(gpu_ctx, bases, [0, n], 64, 1)


def test_second_pass_when_survivors_exceed_the_estimate(gpu_ctx):
    # every window is the same k-mer and it survives: far more kept hashes than windows x (max_hash / 2^64)
    k = 31
    h = so.murmur3_x64_128(b"A" * k, 42)[0]
    n = 300_000
    bases = np.full(n, ord("A"), dtype=np.uint8)
    got_h, got_a, got_o, n_kmers = gpu_ctx.sketch_sequences(bases, [0, n], k, h, 42)
    assert got_h.tolist() == [h] and got_a.tolist() == [n - k + 1] and n_kmers == n - k + 1


def test_bad_arguments_are_refused(gpu_ctx):
    from yacht_b200._lib import YgpuError
    bases = np.frombuffer(b"ACGT" * 100, dtype=np.uint8)
    with pytest.raises(YgpuError):
        gpu_ctx.sketch_sequences(bases, [0, 400], 0, 2 ** 64 - 1)
    with pytest.raises(YgpuError):
        gpu_ctx.sketch_sequences(bases, [0, 400], 257, 2 ** 64 - 1)
    with pytest.raises(YgpuError):
        gpu_ctx.sketch_sequences(bases, [0, 300], 31, 2 ** 64 - 1)          # offsets must end at n_bases
    with pytest.raises(YgpuError):
        gpu_ctx.sketch_sequences(bases, [0, 300, 200, 400], 31, 2 ** 64 - 1)


def _write_fasta(path, records, width=80, gz=False):
    opener = gzip.open if gz else open
    with opener(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + b"\n")


def test_sketch_ref_folder_and_single_file(tmp_path):
    """`yacht sketch ref`: folder -> one sketch per file named after it; single file -> one sketch per record."""
    import argparse
    from yacht_b200 import sigio, sketch_ref_genomes
    rng = np.random.default_rng(21)
    folder = tmp_path / "genomes"
    folder.mkdir()
    files = {}
    for g in range(5):
        recs = [(f"contig{g}_{c} desc", _random_sequence(rng, int(rng.integers(20_000, 120_000))).tobytes().replace(b"*", b"N")) for c in range(1 + g % 3)]
        ext = [".fna.gz", ".fa", ".fasta", ".fna", ".fa.gz"][g]
        path = folder / f"GCF_{g:05d}.1_genomic{ext}"
        _write_fasta(str(path), recs, gz=ext.endswith(".gz"))
        files[f"GCF_{g:05d}.1_genomic"] = (str(path), recs)
    out = tmp_path / "ref.sig.zip"
    sketch_ref_genomes.main(argparse.Namespace(infile=str(folder), kmer=31, scaled=100, outfile=str(out)))
    sigs = {s.name: s for s in sigio.read_sig_zip(str(out))}
    assert sorted(sigs) == sorted(files)
    assert os.path.exists(folder / "dataset.csv")
    for name, (path, recs) in files.items():
        mins, ab = so.sketch_records([s for _, s in recs], 31, 100)
        s = sigs[name]
        assert s.ksize == 31 and s.scaled == 100
        assert np.array_equal(np.asarray(s.mins, dtype=np.uint64), mins), name
        assert np.array_equal(np.asarray(s.abundances, dtype=np.uint32), ab), name
        assert s.md5sum == sigio.compute_md5sum(31, [int(h) for h in mins])
    # single file: --singleton
    name0 = sorted(files)[2]
    path0, recs0 = files[name0]
    out1 = tmp_path / "single.sig.zip"
    sketch_ref_genomes.main(argparse.Namespace(infile=path0, kmer=21, scaled=50, outfile=str(out1)))
    sigs1 = sigio.read_sig_zip(str(out1))
    assert sorted(s.name for s in sigs1) == sorted(n for n, _ in recs0)
    by_name = {s.name: s for s in sigs1}
    for n, seq in recs0:
        mins, ab = so.sketch_records([seq], 21, 50)
        assert np.array_equal(np.asarray(by_name[n].mins, dtype=np.uint64), mins)
        assert np.array_equal(np.asarray(by_name[n].abundances, dtype=np.uint32), ab)


def test_sketch_sample_single_and_paired(tmp_path):
    import argparse
    from yacht_b200 import sigio, sketch_sample
    rng = np.random.default_rng(22)
    genome = _random_sequence(rng, 200_000, p_bad=0.0).tobytes()

    def reads(n, seed):
        r = np.random.default_rng(seed)
        out = []
        for i in range(n):
            p = int(r.integers(0, len(genome) - 150))
            out.append((f"read{i}/1", genome[p:p + 150]))
        return out
    r1, r2 = reads(4000, 1), reads(4000, 2)
    fq1, fq2 = tmp_path / "s_1.fq", tmp_path / "s_2.fq"
    for path, rs in ((fq1, r1), (fq2, r2)):
        with open(path, "wb") as f:
            for name, seq in rs:
                f.write(b"@" + name.encode() + b"\n" + seq + b"\n+\n" + b"F" * len(seq) + b"\n")
    out = tmp_path / "sample.sig.zip"
    sketch_sample.main(argparse.Namespace(infile=[str(fq1)], kmer=31, scaled=100, outfile=str(out)))
    sig = sigio.load_signature_with_ksize(str(out), 31)
    mins, ab = so.sketch_records([s for _, s in r1], 31, 100)
    assert np.array_equal(np.asarray(sig.mins, dtype=np.uint64), mins) and np.array_equal(np.asarray(sig.abundances, dtype=np.uint32), ab)
    assert ab.max() > 1
    out2 = tmp_path / "paired.sig.zip"
    sketch_sample.main(argparse.Namespace(infile=[str(fq1), str(fq2)], kmer=31, scaled=100, outfile=str(out2)))
    sig2 = sigio.load_signature_with_ksize(str(out2), 31)
    mins2, ab2 = so.sketch_records([s for _, s in r1 + r2], 31, 100)
    assert np.array_equal(np.asarray(sig2.mins, dtype=np.uint64), mins2) and np.array_equal(np.asarray(sig2.abundances, dtype=np.uint32), ab2)


def test_groups_larger_than_a_batch_are_merged(tmp_path, monkeypatch):
    from yacht_b200 import sketch
    rng = np.random.default_rng(23)
    monkeypatch.setattr(sketch, "BATCH_BASES", 50_000)
    unit = _random_sequence(rng, 30_000, p_bad=0.0).tobytes()
    groups = [[unit, unit[:10_000], _random_sequence(rng, 45_000).tobytes(), unit], [_random_sequence(rng, 5_000).tobytes()], [],
              [_random_sequence(rng, 70_000).tobytes()]]
    got = sketch.sketch_record_groups(groups, 31, 50)
    for g, (mins, ab) in zip(groups, got):
        em, ea = so.sketch_records(g, 31, 50)
        assert np.array_equal(mins, em) and np.array_equal(ab, ea)
