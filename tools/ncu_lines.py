#!/usr/bin/env python
"""Per-source-line executed warp instructions / stall samples from an .ncu-rep (needs -lineinfo and
--import-source on).  usage: python tools/ncu_lines.py rep.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, ""])
cur_file = ""
hdr = None
first_kernel_done = False
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Kernel Name":
        if first_kernel_done:
            break
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr and r[0].isdigit() and len(r) > ie and r[ie].isdigit():
        first_kernel_done = True
        k = (cur_file, int(r[0]))
        agg[k][0] += int(r[ie])
        agg[k][1] += int(r[si]) if r[si].isdigit() else 0
        agg[k][2] = r[1]
tot = sum(v[0] for v in agg.values())
stot = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot}, stall samples {stot}")
byfile = collections.Counter()
for (f, l), v in agg.items():
    byfile[f] += v[0]
for f, n in byfile.most_common():
    print(f"  {f:<28} {100 * n / tot:5.1f}%")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{v[0]:>10} {100 * v[0] / tot:5.1f}%  smp {100 * v[1] / max(stot, 1):5.1f}%  {f}:{l}: {v[2].strip()[:95]}")
