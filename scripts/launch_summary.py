#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel the number of
launches, the total and mean device time and its share of the listed launches.  The per-launch
times are cold-cache and serialised, so only the SHARES are comparable with bench.py's stage
timers.  Usage: launch_summary.py launches.csv [skip_first_n] > profiles/<name>.md"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.replace("void ", "").replace("cg::", "").strip()


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, gi, bi = (hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"),
                      hdr.index("Block Size"))
    for r in rd:
        rows.append((short(r[ki]), float(r[vi].replace(",", "")), r[gi], r[bi]))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"launches listed: {len(rows)} (first {skip} skipped), total device time "
          f"{total / 1e6:.3f} ms\n")
    print("| kernel | launches | total ms | mean us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, (n, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.1f} | {100 * ns / total:.1f}% | "
              f"{grid} | {block} |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
