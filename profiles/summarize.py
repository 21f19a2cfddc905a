#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per kernel (stdin -> stdout).

    ncu -i gpurun_out/X.ncu-rep --page raw --csv | python profiles/summarize.py > profiles/rNN_ncu_full_....txt
    ... | python profiles/summarize.py --traffic tc_conv2_bwd=conv3x3_bwd_tc_kernel:0 --source profiles/rNN_....txt

--traffic NAME=KERNEL_SUBSTRING:K  also records dram__bytes_read.sum + dram__bytes_write.sum of the K-th launch whose name
contains KERNEL_SUBSTRING under NAME in profiles/ncu_traffic.json -- the file bench.py reads `roofline.traffic` from."""
import argparse
import csv
import json
import os
import sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
ap = argparse.ArgumentParser()
ap.add_argument("--traffic", action="append", default=[])
ap.add_argument("--source", default="")
args = ap.parse_args()
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
print(f"{'kernel':58s} " + " ".join(f"{n:>9s}" for _, n in WANT))
table = []
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
    table.append((name, vals))
    print(f"{name[:58]:58s} " + " ".join(f"{v:9.2f}" for v in vals))
if args.traffic:
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    for spec in args.traffic:
        name, sel = spec.split("=")
        sub, k = sel.rsplit(":", 1)
        hits = [v for n, v in table if sub in n]
        v = hits[int(k)]
        db[name] = {"dram_bytes": (v[1] + v[2]) * 1e6, "source": args.source,
                    "note": f"{sub} launch {k}: dram__bytes_read.sum {v[1]:.2f} MB + dram__bytes_write.sum {v[2]:.2f} MB per launch, "
                            f"{v[0]:.1f} us under ncu --set full --clock-control none"}
    json.dump(db, open(path, "w"), indent=1)
