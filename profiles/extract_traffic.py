#!/usr/bin/env python
"""ncu raw CSV (`ncu -i X.ncu-rep --page raw --csv`) -> profiles/traffic.json: DRAM bytes (read + write) per launch of
every kernel of the train hot path, keyed the way bench.py's `rooflines` are.
usage: python profiles/extract_traffic.py raw.csv genomes seed "<source description>" > profiles/traffic.json"""
import csv, json, sys

KEYS = [("k2_hist1", "part_hist1"), ("k2_scatter<1", "part_scatter1"), ("k2_scatter<(int)1", "part_scatter1"), ("k2_hist2", "part_hist2"),
        ("k2_scatter<2", "part_scatter2"), ("k2_scatter<(int)2", "part_scatter2"), ("k2_group", "index_grouping"), ("k3_count_flag", "pairwise_count")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
per, out = {}, {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = next((k for pat, k in KEYS if pat in name), None)
    if key is None or key in out:
        continue
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    ms = float(r[ix["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[ix["gpu__time_duration.sum"]]]
    out[key] = int(rd + wr)
    per[name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")] = {"ms_under_ncu": round(ms, 3), "dram_read": int(rd), "dram_write": int(wr)}
if all(k in out for k in ("part_hist1", "part_scatter1", "part_hist2", "part_scatter2")):
    out["index_partition"] = out["part_hist1"] + out["part_scatter1"] + out["part_hist2"] + out["part_scatter2"]
print(json.dumps({"source": sys.argv[4] if len(sys.argv) > 4 else sys.argv[1], "genomes": int(sys.argv[2]), "seed": int(sys.argv[3]),
                  "dram_bytes_per_launch": out, "per_kernel": per}, indent=1))
