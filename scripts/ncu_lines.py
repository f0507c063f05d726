#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep
(compiled with -lineinfo).  Usage: ncu_lines.py file.ncu-rep <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(path, kernel, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name",
                          f"regex:{kernel}", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    li, si = 0, 1
    inst, thr, samp = defaultdict(float), defaultdict(float), defaultdict(float)
    text = {}
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or r[0] == "Line No":
            continue
        try:
            line = int(r[li])
        except ValueError:
            continue
        text.setdefault(line, r[si])
        def f(name):
            try:
                return float(r[col[name]])
            except (ValueError, KeyError):
                return 0.0
        inst[line] += f("Instructions Executed")
        thr[line] += f("Thread Instructions Executed")
        samp[line] += f("# Samples")
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
    for line in sorted(inst, key=lambda k: -samp[k])[:top]:
        print(f"{line:5d} inst {100 * inst[line] / max(tot_i, 1):5.1f}%  samples {100 * samp[line] / max(tot_s, 1):5.1f}%"
              f"  thr/inst {thr[line] / max(inst[line], 1):4.1f}  {text[line].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
