#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-launch duration, DRAM traffic,
occupancy, issue rate and the dominant stall reasons.  Usage: ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "sm__cycles_elapsed.max"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h
             and h.endswith("_per_issue_active.ratio") or
             (h.startswith("smsp__average_warps_issue_stalled") and h.endswith(".ratio"))]
    for r in rows[2:]:
        print("-" * 70)
        for w in WANT:
            if w in idx:
                print(f"{w:62s} {r[idx[w]]} {units[idx[w]]}")
        st = []
        for h in stall:
            try:
                st.append((float(r[idx[h]]), h))
            except ValueError:
                pass
        for v, h in sorted(st, reverse=True)[:6]:
            print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''):50s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
