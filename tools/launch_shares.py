#!/usr/bin/env python
"""Aggregates an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`) into a
per-kernel share table, normalised per training step (a step is recognised by `diversity_finalize_kernel`, which runs
once per step).  Usage: python tools/launch_shares.py launches.csv [title...] > profiles/rNN_launch_shares_*.txt"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:62]


def main():
    path = sys.argv[1]
    title = " ".join(sys.argv[2:])
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(dict(zip(header, r)))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += us
    steps = max(1, agg.get("diversity_finalize_kernel", [1])[0])
    total = sum(v[1] for v in agg.values())
    print(title)
    print("launches %d, steps %d, total %.1f us = %.1f us/step (cold-cache, serialised: compare SHARES)" %
          (sum(v[0] for v in agg.values()), steps, total, total / steps))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s n/step=%6.1f %10.1f us/step %5.1f%%  avg %7.1f" % (k, n / steps, us / steps, 100 * us / total, us / n))
    fam = defaultdict(lambda: [0, 0.0])
    for k, (n, us) in agg.items():
        f = re.sub(r"<.*$", "", k)
        fam[f][0] += n
        fam[f][1] += us
    print("\nby kernel family (template arguments merged):")
    for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:24]:
        print("%-62s n/step=%6.1f %10.1f us/step %5.1f%%  avg %7.1f" % (k, n / steps, us / steps, 100 * us / total, us / n))
    tensor = sum(us for k, (n, us) in fam.items() if k.startswith(("conv_", "wgrad_halo")) and "epilogue" not in k
                 and "tanh" not in k)
    print("tensor-core kernels: %.1f%% of the serialised kernel time" % (100 * tensor / total))


if __name__ == "__main__":
    main()
