#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per kernel (stdin -> stdout)."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
print(f"{'kernel':58s} " + " ".join(f"{n:>9s}" for _, n in WANT))
for r in rows[2:]:
    vals = []
    for key, _ in WANT:
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        u = units[i]
        if u == "ns" or u == "nsecond":
            v /= 1e3
        if u == "ms" or u == "msecond":
            v *= 1e3
        if u == "byte":
            v /= 1e6
        if u == "Kbyte":
            v /= 1e3
        if u == "Gbyte":
            v *= 1e3
        vals.append(v)
    name = r[ik].replace("<unnamed>::", "").replace("void ", "")
    print(f"{name[:58]:58s} " + " ".join(f"{v:9.2f}" for v in vals))
