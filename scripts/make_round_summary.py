#!/usr/bin/env python
"""Collect the bench lines of the round (gpurun_out/*.json written by the final runs) into
profiles/: <label>_bench_lines.json (every line, trimmed of nothing), <label>_scaling.md (the
multi-GPU figures side by side) and <label>_parity_margins.json (copied).
Usage: make_round_summary.py <label>  (reads gpurun_out/final_*.json)"""
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
label = sys.argv[1] if len(sys.argv) > 1 else "r2"
go, out = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
lines = {}
for path in sorted(glob.glob(os.path.join(go, "final_*.json"))):
    rows = [l for l in open(path).read().splitlines() if l.startswith("{")]
    if rows:
        lines[os.path.basename(path)[len("final_"):-len(".json")]] = json.loads(rows[-1])
with open(os.path.join(out, f"{label}_bench_lines.json"), "w") as f:
    json.dump(lines, f, indent=1)
pm = os.path.join(go, "parity_margins.json")
if os.path.exists(pm):
    shutil.copy(pm, os.path.join(out, f"{label}_parity_margins.json"))


def g(d, *ks, default=None):
    for k in ks:
        if not isinstance(d, dict) or k not in d:
            return default
        d = d[k]
    return d


md = [f"# {label}: multi-GPU figures (bench.py lines in {label}_bench_lines.json)\n",
      "`value` = points/s over all ranks (pipelined C2 steps / C3 steps), max-over-ranks step time; "
      "sharded = the server's global merge over all ranks (project into partial layers + exchange + "
      "fold, whole call timed on the device, max over ranks): `native` = cg_project_submaps_sharded "
      "(owner-pull over NVLink peer memory behind the C ABI), `packed` = pack + NCCL all_to_all through "
      "torch.distributed + fold.\n",
      "| run | N | value | ms/step | per-rank ms/step | e2e | sharded native (ms, G voxels/s) | sharded packed (ms) | global blocks | parity vs single-process oracle fold |",
      "|---|---|---|---|---|---|---|---|---|---|"]
for name, d in lines.items():
    sh = g(d, "project_submaps", "sharded") or (g(d, "project_submaps") if g(d, "project_submaps", "native") else None)
    par = (sh or {}).get("parity") or g(d, "project_submaps", "parity")
    md.append("| {} | {} | {:.3g} {} | {:.3f} | {} | {:.3g} | {} | {} | {} | {} |".format(
        name, d.get("n_gpus"), d.get("value", 0), d.get("unit", ""), d.get("ms_per_step", 0),
        ", ".join(f"{x:.2f}" for x in (d.get("per_rank_ms_per_step") or [])) or "—",
        g(d, "e2e", "value", default=0),
        "{:.3f}, {:.1f}".format(g(sh, "native", "ms", default=0), g(sh, "native", "value", default=0) / 1e9) if sh and g(sh, "native") else
        ("{:.3f}, {:.1f} (one GPU)".format(g(d, "project_submaps", "ms", default=0), g(d, "project_submaps", "value", default=0) / 1e9) if g(d, "project_submaps", "ms") else "—"),
        "{:.3f}".format(g(sh, "packed", "ms", default=0)) if sh and g(sh, "packed") else "—",
        g(sh, "native", "global_blocks") if sh else g(d, "project_submaps", "global_blocks", default="—"),
        ("within tolerance: max d err/tol {:.4f}, colour LSB hist {}".format(par.get("max_distance_err_over_tol", 0), par.get("colour_lsb_hist")) if par and par.get("within_tolerance") else (str(par) if par else "—"))))
with open(os.path.join(out, f"{label}_scaling.md"), "w") as f:
    f.write("\n".join(md) + "\n")
print("wrote", [n for n in sorted(os.listdir(out)) if n.startswith(label)])
