"""Pin of the sketching oracle (oracle/sketch_oracle.c) by the reference's own checked-in numbers.

Run in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_sketch_golden.py

The reference ships no sequence together with its sourmash sketch, but it does ship
    demo/ref_genomes/*.fna.gz                                         (15 genomes) and
    tests/testdata/standardize_output_testdata/results/result.xlsx    (a `yacht run` result on a database trained from them,
                                                                       k = 31, scaled = 1000)
whose rows record, per organism, num_unique_kmers_in_genome_sketch and num_total_kmers_in_genome_sketch -- the number of
distinct hashes and the sum of abundances of that genome's sourmash sketch (utils.py:54-75, 89-110).  This script sketches
the demo genomes with the oracle, checks those numbers for every organism the workbook lists, and commits
tests/golden/sketch_golden.json: the workbook numbers, and for all 15 genomes (n_unique, n_total, sha256 of the ascending
mins as decimal lines) so that later changes of the oracle are noticed even where /root/reference is absent.
"""
import glob
import hashlib
import json
import os
import re
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("YACHT_REFERENCE", "/root/reference")

from oracle import sketch_oracle as so  # noqa: E402


def workbook_rows(path):
    """{organism_name: {column: value}} over all sheets (inline strings; openpyxl is not installed)."""
    z = zipfile.ZipFile(path)
    out = {}
    for sheet in sorted(n for n in z.namelist() if n.startswith("xl/worksheets/sheet")):
        xml = z.read(sheet).decode()
        rows = []
        for rm in re.finditer(r"<row [^>]*>(.*?)</row>", xml, re.S):
            cells = []
            for cm in re.finditer(r'<c r="[A-Z]+\d+"[^>]*?(?:/>|>(.*?)</c>)', rm.group(1), re.S):
                body = cm.group(1) or ""
                m = re.search(r"<t[^>]*>(.*?)</t>|<v>(.*?)</v>", body, re.S)
                cells.append((m.group(1) if m.group(1) is not None else m.group(2)) if m else None)
            rows.append(cells)
        header = rows[0]
        for r in rows[1:]:
            rec = dict(zip(header, r))
            prev = out.setdefault(r[0], rec)
            for col in ("num_unique_kmers_in_genome_sketch", "num_total_kmers_in_genome_sketch"):
                assert prev[col] == rec[col], (sheet, r[0], col)          # the sheets (one per min_coverage) agree
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found")
    wb = workbook_rows(os.path.join(REF, "tests/testdata/standardize_output_testdata/results/result.xlsx"))
    out = {"ksize": 31, "scaled": 1000, "seed": 42, "max_hash": so.max_hash_for_scaled(1000),
           "murmur3_x64_128_verification": hex(so.murmur3_verification()),
           "workbook": {}, "demo_genomes": {}}
    for path in sorted(glob.glob(os.path.join(REF, "demo/ref_genomes/*.fna.gz"))):
        name = os.path.basename(path).replace("_genomic.fna.gz", "")
        mins, ab = so.sketch_file(path, 31, 1000)
        digest = hashlib.sha256("".join(f"{int(h)}\n" for h in mins).encode()).hexdigest()
        out["demo_genomes"][name] = {"n_unique": int(len(mins)), "n_total": int(ab.sum()), "mins_sha256": digest}
        if name in wb:
            want = {"num_unique_kmers_in_genome_sketch": int(wb[name]["num_unique_kmers_in_genome_sketch"]),
                    "num_total_kmers_in_genome_sketch": int(wb[name]["num_total_kmers_in_genome_sketch"])}
            out["workbook"][name] = want
            assert (len(mins), int(ab.sum())) == (want["num_unique_kmers_in_genome_sketch"], want["num_total_kmers_in_genome_sketch"]), (name, len(mins), int(ab.sum()), want)
            print(f"{name}: {len(mins)} distinct / {int(ab.sum())} total == workbook")
    assert len(out["workbook"]) == 5, "the workbook rows were not found"
    with open(os.path.join(HERE, "sketch_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote sketch_golden.json: {len(out['demo_genomes'])} genomes, {len(out['workbook'])} pinned by the workbook")


if __name__ == "__main__":
    main()
