"""BASELINE.json configs[0] in shape: sequence files -> `yacht sketch ref` -> `yacht train` -> `yacht sketch sample` ->
`yacht run`, every step through this repo's CLI (`python -m yacht_b200 ...`, same arguments as the reference's README
demo) on synthetic genomes, checked against the CPU oracles chained the same way (sketch_oracle -> train_oracle -> run_oracle).
The reference's demo data itself cannot travel to the GPU box; its sketch counts pin the sketching oracle
(tests/golden/sketch_golden.json)."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from oracle import run_oracle as ro
from oracle import sketch_oracle as so
from oracle import train_oracle as to
from yacht_b200 import sigio, xlsx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, SCALED, ANI = 31, 100, 0.95


def _yacht(*argv):
    return subprocess.run([sys.executable, "-m", "yacht_b200", *argv], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def _mutate(rng, seq: np.ndarray, rate: float) -> np.ndarray:
    out = seq.copy()
    pos = np.flatnonzero(rng.random(len(seq)) < rate)
    out[pos] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=len(pos))
    return out


def test_sequence_files_to_result_workbook(tmp_path):
    rng = np.random.default_rng(2024)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    folder = tmp_path / "ref_genomes"
    folder.mkdir()
    genomes = {}
    base = [rng.choice(acgt, size=int(rng.integers(90_000, 150_000))) for _ in range(10)]
    seqs = base + [_mutate(rng, base[0], 0.002), _mutate(rng, base[7], 0.002)]          # two near-duplicates (ANI ~ 0.998)
    for g, seq in enumerate(seqs):
        name = f"GCF_{g:06d}.1"
        cut = len(seq) // 3
        contigs = [seq[:cut].tobytes(), seq[cut:].tobytes().lower() if g % 4 == 1 else seq[cut:].tobytes()]
        with gzip.open(folder / f"{name}_genomic.fna.gz", "wb") as f:
            for c, s in enumerate(contigs):
                f.write(f">{name}_contig{c} synthetic\n".encode())
                for i in range(0, len(s), 80):
                    f.write(s[i:i + 80] + b"\n")
        genomes[f"{name}_genomic"] = contigs
    d = str(tmp_path)
    ref_zip = os.path.join(d, "ref.sig.zip")
    res = _yacht("sketch", "ref", "--infile", str(folder), "--kmer", str(K), "--scaled", str(SCALED), "--outfile", ref_zip)
    assert res.returncode == 0 and os.path.exists(ref_zip), res.stderr[-2000:]

    res = _yacht("train", "--force", "--ref_file", ref_zip, "--ksize", str(K), "--prefix", "demo", "--ani_thresh", str(ANI),
                 "--outdir", d, "--num_threads", "4")
    assert res.returncode == 0, res.stderr[-3000:]
    cfg_path = os.path.join(d, "demo_config.json")
    cfg = json.load(open(cfg_path))
    assert cfg["ksize"] == K and cfg["scale"] == SCALED
    manifest = pd.read_csv(os.path.join(d, "demo_processed_manifest.tsv"), sep="\t")

    # the same chain on the CPU oracles: sketches -> all-vs-all containment + greedy selection
    names = sorted(genomes)
    sk = {n: so.sketch_records(genomes[n], K, SCALED) for n in names}
    for n in names:
        row = manifest[manifest["organism_name"] == n]
        if len(row):
            assert int(row["num_unique_kmers_in_genome_sketch"].iloc[0]) == len(sk[n][0])
            assert int(row["num_total_kmers_in_genome_sketch"].iloc[0]) == int(sk[n][1].sum())
    assert len(manifest) == 10                                        # one genome of each near-duplicate pair is dropped
    for a, b in (("GCF_000000.1_genomic", "GCF_000010.1_genomic"), ("GCF_000007.1_genomic", "GCF_000011.1_genomic")):
        assert (a in set(manifest["organism_name"])) != (b in set(manifest["organism_name"]))
    # train order inside the tool is its own file order; the oracle is asked about the same SET of retained genomes
    order = [n for n in pd.read_csv(os.path.join(d, "demo_intermediate_files", "training_sig_files.tsv"), header=None)[0]]
    assert len(order) == 12
    by_md5 = {os.path.basename(p)[:-4]: p for p in order}
    md5_of = {n: sigio.compute_md5sum(K, [int(h) for h in sk[n][0]]) for n in names}
    train_names = [next(n for n in names if md5_of[n] == os.path.basename(p)[:-4]) for p in order]
    hashes = np.concatenate([sk[n][0] for n in train_names])
    offsets = np.zeros(len(train_names) + 1, np.uint64)
    np.cumsum([len(sk[n][0]) for n in train_names], out=offsets[1:])
    ref = to.oracle_train(hashes, offsets, ANI ** K)
    kept = sorted(train_names[int(g)] for g in ref.selected)
    assert kept == sorted(manifest["organism_name"])
    assert len(by_md5) == 12

    # sample: reads from three of the retained genomes
    present = [n for n in manifest["organism_name"] if n in ("GCF_000003.1_genomic", "GCF_000004.1_genomic", "GCF_000005.1_genomic")]
    assert len(present) == 3
    reads = []
    for n in present:
        whole = b"".join(genomes[n]).upper()
        for _ in range(int(3 * len(whole) / 150)):
            p = int(rng.integers(0, len(whole) - 150))
            reads.append(whole[p:p + 150])
    fq = os.path.join(d, "sample.fq")
    with open(fq, "wb") as f:
        for i, r in enumerate(reads):
            f.write(f"@read{i}\n".encode() + r + b"\n+\n" + b"F" * len(r) + b"\n")
    sample_zip = os.path.join(d, "sample.sig.zip")
    res = _yacht("sketch", "sample", "--infile", fq, "--kmer", str(K), "--scaled", str(SCALED), "--outfile", sample_zip)
    assert res.returncode == 0 and os.path.exists(sample_zip), res.stderr[-2000:]

    res = _yacht("run", "--json", cfg_path, "--sample_file", sample_zip, "--significance", "0.99", "--min_coverage_list", "1", "0.5", "0.1",
                 "--outdir", d, "--show_all", "--num_threads", "4")
    assert res.returncode == 0, res.stderr[-3000:]
    sheets = xlsx.read_xlsx(os.path.join(d, "results", "result.xlsx"))
    assert list(sheets.keys()) == ["min_coverage1.0", "min_coverage0.5", "min_coverage0.1"]
    df = sheets["min_coverage0.5"]
    found = sorted(df[df["in_sample_est"].astype(str) == "True"]["organism_name"])
    assert found == sorted(present)

    # exclusive hashes / matches per organism against the run oracle on the oracle's sketches
    m_names = list(manifest["organism_name"])
    mh = np.concatenate([sk[n][0] for n in m_names])
    mo = np.zeros(len(m_names) + 1, np.uint64)
    np.cumsum([len(sk[n][0]) for n in m_names], out=mo[1:])
    sample_mins, sample_ab = so.sketch_records(reads, K, SCALED)
    exp = ro.exclusive_counts(mh, mo, sample_mins)
    assert len(df) == int((exp["n_overlap"] > 0).sum())
    for _, row in df.iterrows():
        g = m_names.index(row["organism_name"])
        assert int(row["num_exclusive_kmers_to_genome"]) == int(exp["n_exclusive"][g])
        assert int(row["num_matches"]) == int(exp["n_match"][g])
        assert int(row["num_total_kmers_in_sample_sketch"]) == int(sample_ab.sum())
