"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[r[ki][:70]].append(v)
for k, v in agg.items():
    print(f"{k:70s} n={len(v):4d} mean={sum(v) / len(v):10.1f} us  min={min(v):10.1f}")
