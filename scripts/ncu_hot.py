#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel in an .ncu-rep (source page).
Usage: ncu_hot.py file.ncu-rep kernel_regex [N]"""
import csv
import io
import subprocess
import sys

path, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name",
                      "regex:" + rx], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:
    if len(r) < len(hdr) or r[0].startswith("Kernel") or r[0] == "Address":
        break
    try:
        data.append((int(r[ix["# Samples"]] or 0), r))
    except ValueError:
        break
tot = sum(s for s, _ in data) or 1
print(f"total samples {tot}, instructions {len(data)}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for s, r in sorted(data, key=lambda t: -t[0])[:n]:
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100 * s / tot:5.1f}%  thr {r[ix['Avg. Threads Executed']]:>5s}  "
          f"{r[ix['Source']][:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")

if len(sys.argv) > 4:   # context: dump instructions [a, b) by position with samples
    a, b = map(int, sys.argv[4].split(":"))
    for i, (s, r) in enumerate(data[a:b], start=a):
        print(f"{i:4d} {100 * s / tot:5.1f}% thr {r[ix['Avg. Threads Executed']]:>5s} {r[ix['Source']][:90]}")
