#!/usr/bin/env python
"""Print the interesting fields of bench.py JSON lines.  Usage: show_bench.py file..."""
import json
import sys
for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(f"== {f}: {d.get('impl', 'ours')} value {d.get('value'):.4g} {d.get('unit')} "
              f"ms/step {d.get('ms_per_step'):.4g}")
        for k in ("integrate", "merge", "project_submaps", "two_jobs_in_flight", "per_frame_call", "e2e", "gpu_launches", "roofline", "stages_ms_per_step",
                  "per_step", "cpu_baseline", "clocks"):
            if k in d:
                print("  ", k, d[k])
