"""Host-side mirror of the reference's run-side compute (reference
src/yacht/hypothesis_recovery_src.py) on top of the C ABI: same function names, arguments, return
shapes and error behaviour; the arithmetic runs on the GPU.

reference function                      here
--------------------------------------  -----------------------------------------------------------
get_organisms_with_nonzero_overlap :30  K5 phase A (ygpu_exclusive_hashes: n_overlap > 0) instead of
                                        the `sourmash scripts multisearch ... -t 0` subprocess; the
                                        same side files are written (sample_sig_file.txt,
                                        organism_sig_file.txt, sample_multisearch_result.csv with a
                                        ``match_name`` column)
get_exclusive_hashes              :116  K5 phase B (ygpu_exclusive_hashes with the nontrivial mask)
get_alt_mut_rate                  :209  ygpu_alt_mut_rate
single_hyp_test                   :233  ygpu_hyp_test (one row)
hypothesis_recovery               :309  one K5 pass + one K6 launch for all coverages
"""
from __future__ import annotations

import glob
import os
import shutil
import zipfile
from typing import List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import _lib, dbcache, sigio
from .utils import _log, decompress_all_sig_files, load_signature_with_ksize

SIG_SUFFIX = ".sig"

_ctx: Optional[_lib.GpuContext] = None
_loaded_key = None


def _context() -> _lib.GpuContext:
    """One lazily created context on the device named by YACHT_DEVICE (default 0)."""
    global _ctx
    if _ctx is None:
        _ctx = _lib.GpuContext(int(os.environ.get("YACHT_DEVICE", "0")))
    return _ctx


_last_counts = None      # (loaded database key, sample key, counts) of the last overlap probe


def _load_manifest_sketches(manifest: pd.DataFrame, path_to_genome_temp_dir: str, num_threads: int, ksize: int = 0) -> None:
    """Make the manifest's genomes (row order = genome id) resident on the device; cached per manifest."""
    global _loaded_key
    paths = [os.path.join(path_to_genome_temp_dir, "signatures", md5sum + SIG_SUFFIX) for md5sum in manifest["md5sum"]]
    key = (path_to_genome_temp_dir, tuple(manifest["md5sum"]))
    if _loaded_key == key:
        return
    md5sums = list(manifest["md5sum"])
    cached = dbcache.load(path_to_genome_temp_dir, md5sums)
    if cached is not None:
        hashes, offsets = cached
        _log("INFO", f"Packed sketch cache found ({len(md5sums)} genomes, {int(offsets[-1])} hashes): signature files are not parsed")
    else:
        # like the reference (load_signature_with_ksize per manifest row, hypothesis_recovery_src.py:154-172): the sub-signature of
        # THIS k-mer size, exactly one per file -- a file with several ksizes must not contribute whichever comes first
        hashes, offsets, n_bad = _lib.read_signatures(paths, max(1, int(num_threads)), ksize=int(ksize))
        if n_bad:
            # the reference loads every manifest signature through load_signature_with_ksize and raises when one is missing
            # (utils.py:43-50); an absent file must not silently become an empty sketch here
            raise ValueError(f"{n_bad} reference signature file(s) of the manifest could not be opened under "
                             f"{os.path.join(path_to_genome_temp_dir, 'signatures')}")
        dbcache.store(path_to_genome_temp_dir, md5sums, hashes, offsets)
    _context().load_sketches(hashes, offsets)
    _loaded_key = key


def get_organisms_with_nonzero_overlap(manifest: pd.DataFrame, sample_file: str, scale: int, ksize: int, num_threads: int,
                                       path_to_genome_temp_dir: str, path_to_sample_temp_dir: str) -> List[str]:
    """reference :30-113 -- names of the organisms whose sketch shares at least one hash with the sample."""
    _log("INFO", "Unzipping the sample signature zip file")
    with zipfile.ZipFile(sample_file, "r") as sample_zip_file:
        sample_zip_file.extractall(path_to_sample_temp_dir)
    all_gz_files = glob.glob(f"{path_to_sample_temp_dir}/signatures/*.sig.gz")
    _log("INFO", f"Decompressing {len(all_gz_files)} .sig.gz files using {num_threads} threads.")
    decompress_all_sig_files(all_gz_files, num_threads)

    sample_sig_files = [os.path.join(path_to_sample_temp_dir, "signatures", f)
                        for f in os.listdir(os.path.join(path_to_sample_temp_dir, "signatures"))]
    pd.DataFrame(sample_sig_files).to_csv(os.path.join(path_to_sample_temp_dir, "sample_sig_file.txt"), header=False, index=False)
    organism_sig_file = pd.DataFrame([os.path.join(path_to_genome_temp_dir, "signatures", md5sum + SIG_SUFFIX)
                                      for md5sum in manifest["md5sum"]])
    organism_sig_file.to_csv(os.path.join(path_to_sample_temp_dir, "organism_sig_file.txt"), header=False, index=False)

    # every sample signature of this ksize is a query, like multisearch's query list
    result_file = os.path.join(path_to_sample_temp_dir, "sample_multisearch_result.csv")
    _load_manifest_sketches(manifest, path_to_genome_temp_dir, num_threads, ksize)
    ctx = _context()
    names: List[str] = []
    rows = []
    global _last_counts
    _last_counts = None
    for sf in sample_sig_files:
        for sig in sigio.parse_signature_json(sigio._open_text(sf), sf):
            if sig.ksize != ksize or sig.scaled != scale:
                continue
            counts = ctx.exclusive_hashes(sig.mins)
            # one probe of the reference serves both steps: with the nontrivial genomes = those with overlap (what
            # hypothesis_recovery passes on), the exclusive-hash counts of get_exclusive_hashes are already in `counts`
            _last_counts = (_loaded_key, sig.mins.tobytes() if len(sig.mins) < (1 << 22) else (len(sig.mins), int(sig.mins[0]), int(sig.mins[-1])), counts)
            hit = np.flatnonzero(counts["n_overlap"] > 0)
            for g in hit:
                rows.append((sig.name, sig.md5sum, manifest["organism_name"].iloc[int(g)], manifest["md5sum"].iloc[int(g)],
                             int(counts["n_overlap"][g]) / max(len(sig), 1), int(counts["n_overlap"][g])))
    if not rows:
        open(result_file, "w").close()
        print("ERROR: Multisearch file is empty. Likely there are no microorganisms in your sample, or something went wrong", flush=True)
        exit(0)
    res = pd.DataFrame(rows, columns=["query_name", "query_md5", "match_name", "match_md5", "containment", "intersect_hashes"])
    res.to_csv(result_file, index=False)
    res = res.drop_duplicates().reset_index(drop=True)
    names = res["match_name"].to_list()
    return names


def get_exclusive_hashes(manifest: pd.DataFrame, nontrivial_organism_names: List[str], sample_sig, ksize: int,
                         path_to_genome_temp_dir: str, num_threads: int = 1) -> Tuple[List[Tuple[int, int]], pd.DataFrame]:
    """reference :116-206 -- [(n exclusive hashes, n exclusive hashes in the sample)] per nontrivial organism, in
    sub-manifest order, and the sub-manifest."""
    keep = manifest["organism_name"].isin(nontrivial_organism_names)
    sub_manifest = manifest.loc[keep, :].reset_index(drop=True)
    _load_manifest_sketches(manifest, path_to_genome_temp_dir, num_threads, ksize)
    mask = keep.to_numpy().astype(np.uint8)
    sample_hashes = sample_sig.mins if hasattr(sample_sig, "mins") else np.asarray(list(sample_sig.minhash.hashes), dtype=np.uint64)
    sample_hashes = np.ascontiguousarray(sample_hashes, dtype=np.uint64)
    skey = sample_hashes.tobytes() if len(sample_hashes) < (1 << 22) else (len(sample_hashes), int(sample_hashes[0]), int(sample_hashes[-1]))
    if (_last_counts is not None and _last_counts[0] == _loaded_key and _last_counts[1] == skey
            and np.array_equal(_last_counts[2]["nontrivial"].astype(np.uint8), mask)):
        counts = _last_counts[2]          # same sample, same nontrivial set: the overlap probe already computed them
    else:
        counts = _context().exclusive_hashes(sample_hashes, mask)
    ids = np.flatnonzero(mask)
    info = [(int(counts["n_exclusive"][g]), int(counts["n_match"][g])) for g in ids]
    return info, sub_manifest


def get_alt_mut_rate(nu: int, thresh: int, ksize: int, significance: float = 0.99) -> float:
    """reference :209-230."""
    return float(_context().alt_mut_rate([int(nu)], [int(thresh)], ksize, significance)[0])


def _row_tuple(r) -> Tuple[bool, float, int, int, int, float, float, float]:
    return (bool(r["in_sample_est"]), float(r["p_val"]), int(r["num_exclusive_kmers"]), int(r["num_exclusive_kmers_coverage"]),
            int(r["num_matches"]), float(r["acceptance_threshold_with_coverage"]), float(r["actual_confidence_with_coverage"]),
            float(r["alt_confidence_mut_rate_with_coverage"]))


def single_hyp_test(exclusive_hashes_info_org: Tuple[int, int], ksize: int, significance: float = 0.99,
                    ani_thresh: float = 0.95, min_coverage: float = 1) -> Tuple[bool, float, int, int, int, float, float, float]:
    """reference :233-306 -- the 8-tuple, same order as the reference returns it."""
    rows = _context().hyp_test([int(exclusive_hashes_info_org[0])], [int(exclusive_hashes_info_org[1])], ksize, significance,
                               ani_thresh, [float(min_coverage)])
    return _row_tuple(rows[0, 0])


GIVEN_COLUMNS = [
    "in_sample_est", "p_vals", "num_exclusive_kmers_to_genome", "num_exclusive_kmers_to_genome_coverage", "num_matches",
    "acceptance_threshold_with_coverage", "actual_confidence_with_coverage", "alt_confidence_mut_rate_with_coverage",
]


def hypothesis_recovery(manifest: pd.DataFrame, sample_info_set, path_to_genome_temp_dir: str, min_coverage_list: List[float],
                        scale: int, ksize: int, significance: float = 0.99, ani_thresh: float = 0.95, num_threads: int = 16):
    """reference :309-417 -- one DataFrame per min_coverage (manifest columns + the 8 result columns)."""
    sample_file, sample_sig = sample_info_set
    sample_dir = os.path.dirname(sample_file)
    sample_name = os.path.basename(sample_file).replace(".sig.zip", "")
    path_to_sample_temp_dir = os.path.join(sample_dir, f"sample_{sample_name}_intermediate_files")
    if os.path.exists(path_to_sample_temp_dir):
        _log("INFO", f"Removing existing temporary directory: {path_to_sample_temp_dir}")
        shutil.rmtree(path_to_sample_temp_dir)
    os.makedirs(path_to_sample_temp_dir)

    nontrivial_organism_names = get_organisms_with_nonzero_overlap(manifest, sample_file, scale, ksize, num_threads,
                                                                   path_to_genome_temp_dir, path_to_sample_temp_dir)
    exclusive_hashes_info, manifest = get_exclusive_hashes(manifest, nontrivial_organism_names, sample_sig, ksize,
                                                           path_to_genome_temp_dir, num_threads)
    ne = [e[0] for e in exclusive_hashes_info]
    nm = [e[1] for e in exclusive_hashes_info]
    covs = [float(c) for c in min_coverage_list]
    rows = _context().hyp_test(ne, nm, ksize, significance, ani_thresh, covs) if len(ne) and len(covs) else None

    manifest_list = []
    for c, min_coverage in enumerate(min_coverage_list):
        _log("INFO", f"Computing hypothesis recovery for min_coverage={min_coverage}")
        results = pd.DataFrame([_row_tuple(rows[c, r]) for r in range(len(ne))] if rows is not None else [], columns=GIVEN_COLUMNS)
        manifest["min_coverage"] = min_coverage
        manifest_list.append(pd.concat([manifest, results], axis=1))
    return manifest_list
