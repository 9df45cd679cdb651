"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. usage: launch_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h = r
        start = i + 1
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) > vi:
        try:
            d[r[ki][:70]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):4d} avg={sum(v) / len(v) / 1000:10.1f} us  share={sum(v) / tot:.3f}")
