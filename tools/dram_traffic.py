#!/usr/bin/env python
"""Turns an ncu launch list that carries `dram__bytes_read.sum` and `dram__bytes_write.sum` (beside gpu__time_duration.sum)
into profiles/rNN_dram_traffic.json: DRAM bytes per launch of each tensor kernel, averaged over the captured launches.
bench.py copies the dominant kernel's figure into `roofline.traffic`.

Usage: python tools/dram_traffic.py launches.csv "<the ncu command>" > profiles/r02_dram_traffic.json"""
import csv
import json
import re
import sys
from collections import defaultdict

KERNELS = ["conv_stack3_kernel", "conv_halo2_kernel", "conv_halo_kernel", "wgrad_halo_kernel", "conv_wgrad_kernel", "conv_fprop_kernel"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path, source = sys.argv[1], " ".join(sys.argv[2:])
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    header, per_launch = None, defaultdict(dict)
    for r in csv.reader(lines):
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        row = dict(zip(header, r))
        m = row.get("Metric Name", "")
        if not m.startswith("dram__bytes"):
            continue
        val = float(row["Metric Value"].replace(",", "")) * SCALE.get(row.get("Metric Unit", "byte"), 1.0)
        per_launch[(row["ID"], row["Kernel Name"])][m] = val
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for (_id, name), vals in per_launch.items():
        for k in KERNELS:
            if re.search(r"\b%s\b" % k, name):
                agg[k][0] += 1
                agg[k][1] += vals.get("dram__bytes_read.sum", 0.0)
                agg[k][2] += vals.get("dram__bytes_write.sum", 0.0)
                break
    out = {}
    for k, (n, rd, wr) in agg.items():
        out[k] = {"dram_bytes_per_launch": (rd + wr) / n, "read": rd / n, "write": wr / n, "launches_captured": n,
                  "source": source}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
