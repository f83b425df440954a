"""Extract the run-path golden vectors from the reference's checked-in result workbooks.

Run in the build container (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_run_golden.py

Every data row of every sheet of
    use_case_examples/low_abundance_samples/result_k51_ani0.95.xlsx      (k=51, ANI 0.95, significance 0.95)
    use_case_examples/MAG_fishing/result_k51_ani0.95_SRR32008482.xlsx    (same parameters)
    use_case_examples/low_abundance_samples/result_k31_ani0.90.xlsx      (k=31, ANI 0.90, significance 0.90)
    tests/testdata/standardize_output_testdata/results/result.xlsx       (k=31, ANI 0.95, significance 0.99)
is one evaluation of the reference's single_hyp_test (hypothesis_recovery_src.py:233-306):
inputs (num_exclusive_kmers_to_genome, min_coverage, num_matches) and outputs
(num_exclusive_kmers_to_genome_coverage, acceptance threshold, actual confidence, alt mutation
rate, p-value, in_sample_est).  The (ksize, ani, significance) of each workbook are not stored
in it; they are the ones SURVEY.md 8c determined (every row reproduces under them).  Rows are
de-duplicated on the input triple and written to tests/golden/run_golden.json.gz.

The workbooks are read with zipfile + regular expressions (openpyxl is not installed).
The last workbook's p_vals column was produced by an older formula (full n instead of n*cov):
its p-values are kept only for min_coverage == 1 rows (p_ok flag).
"""
import gzip
import json
import os
import re
import sys
import zipfile

REF = os.environ.get("YACHT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

BOOKS = [
    ("use_case_examples/low_abundance_samples/result_k51_ani0.95.xlsx", 51, 0.95, 0.95, True),
    ("use_case_examples/MAG_fishing/result_k51_ani0.95_SRR32008482.xlsx", 51, 0.95, 0.95, True),
    ("use_case_examples/low_abundance_samples/result_k31_ani0.90.xlsx", 31, 0.90, 0.90, True),
    ("tests/testdata/standardize_output_testdata/results/result.xlsx", 31, 0.95, 0.99, False),
]

ROW_RE = re.compile(r"<row [^>]*>(.*?)</row>", re.S)
CELL_RE = re.compile(r'<c r="([A-Z]+)\d+"(?: s="\d+")?(?: t="(\w+)")?\s*(?:/>|>(.*?)</c>)', re.S)
VAL_RE = re.compile(r"<v>(.*?)</v>|<t[^>]*>(.*?)</t>", re.S)


def read_sheet(xml: str, shared):
    rows = []
    for rm in ROW_RE.finditer(xml):
        cells = {}
        for col, typ, body in CELL_RE.findall(rm.group(1)):
            if not body:
                continue
            vm = VAL_RE.search(body)
            if not vm:
                continue
            txt = vm.group(1) if vm.group(1) is not None else vm.group(2)
            if typ == "s":
                val = shared[int(txt)]
            elif typ in ("inlineStr", "str"):
                val = txt
            elif typ == "b":
                val = bool(int(txt))
            else:
                val = float(txt)
            cells[col] = val
        rows.append(cells)
    return rows


def col_letters(k):
    s = ""
    k += 1
    while k:
        k, r = divmod(k - 1, 26)
        s = chr(65 + r) + s
    return s


def main():
    out = {"books": []}
    total = 0
    for rel, k, ani, sig, p_ok_all in BOOKS:
        path = os.path.join(REF, rel)
        z = zipfile.ZipFile(path)
        shared = []
        if "xl/sharedStrings.xml" in z.namelist():
            shared = [re.sub(r"<[^>]+>", "", m) for m in re.findall(r"<si>(.*?)</si>", z.read("xl/sharedStrings.xml").decode(), re.S)]
        sheets = sorted(n for n in z.namelist() if n.startswith("xl/worksheets/sheet"))
        seen = {}
        nrows = 0
        for sh in sheets:
            rows = read_sheet(z.read(sh).decode(), shared)
            if not rows:
                continue
            header = {v: c for c, v in rows[0].items()}
            def col(*names):
                for nm in names:
                    if nm in header:
                        return header[nm]
                raise KeyError(names)
            c_cov = col("min_coverage"); c_in = col("in_sample_est"); c_p = col("p_vals")
            c_ne = col("num_exclusive_kmers_to_genome"); c_nc = col("num_exclusive_kmers_to_genome_coverage")
            c_m = col("num_matches")
            c_thr = col("acceptance_threshold_with_coverage", "acceptance_threshold_wo_coverage")
            c_conf = col("actual_confidence_with_coverage", "actual_confidence_wo_coverage")
            c_alt = col("alt_confidence_mut_rate_with_coverage", "alt_confidence_mut_rate_wo_coverage")
            for r in rows[1:]:
                if c_ne not in r:
                    continue
                nrows += 1
                cov = float(r[c_cov])
                key = (int(r[c_ne]), cov, int(r[c_m]))
                p_ok = bool(p_ok_all or cov == 1.0)
                rec = [key[0], cov, key[2], int(r[c_nc]), float(r[c_thr]), float(r[c_conf]), float(r[c_alt]),
                       float(r[c_p]), bool(r[c_in]), p_ok]
                if key in seen:
                    assert seen[key][:7] == rec[:7], (rel, key, seen[key], rec)
                else:
                    seen[key] = rec
        recs = sorted(seen.values())
        total += len(recs)
        print(f"{rel}: {nrows} rows -> {len(recs)} distinct evaluations", file=sys.stderr)
        out["books"].append({"source": rel, "ksize": k, "ani_thresh": ani, "significance": sig, "n_rows": nrows,
                             "columns": ["n_excl", "min_coverage", "num_matches", "n_cov", "thr", "conf", "alt", "p_val", "in_sample", "p_ok"],
                             "rows": recs})
    dst = os.path.join(HERE, "run_golden.json.gz")
    with gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print(f"wrote {dst}: {total} evaluations, {os.path.getsize(dst)} bytes", file=sys.stderr)


if __name__ == "__main__":
    main()
