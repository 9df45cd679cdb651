#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
Prints, per captured launch: duration, DRAM bytes, L2/L1/SM throughput, occupancy, registers,
then the opcode mix and the 25 SASS lines with most stall samples of the first launch."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "lts__t_sectors.sum.per_second", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "dram__bytes.sum.per_second",
        "sm__warps_active.avg.per_cycle_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k} = {r[i]} {units[i]}")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return
    h = rows[hi[0]]
    body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
    si, ie, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    ops, smp = collections.Counter(), collections.Counter()
    for r in body:
        if not r[ie].isdigit():
            continue
        t = re.sub(r"^@!?U?P\d\s+", "", r[src].strip())
        op = t.split()[0].split(".")[0]
        ops[op] += int(r[ie])
        smp[op] += int(r[si])
    tot, stot = sum(ops.values()), sum(smp.values())
    print(f"\nopcode mix of launch 0 (warp instructions {tot}, stall samples {stot}):")
    for op, n in ops.most_common(18):
        print(f"  {op:<10} {n:>12} {100 * n / tot:5.1f}%  samples {smp[op]:>7} {100 * smp[op] / max(stot, 1):5.1f}%")
    print("\ntop SASS lines by stall samples (launch 0):")
    for r in sorted(body, key=lambda r: -int(r[si]) if r[si].isdigit() else 0)[:25]:
        print(f"  {r[si]:>7} {r[ie]:>10}  {r[src].strip()[:100]}")


if __name__ == "__main__":
    main()
