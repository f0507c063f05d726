#!/usr/bin/env python
"""Turn the artefacts of scripts/profile_round.sh (gpurun_out/) into the committed summaries under
profiles/: the launch list, the ncu --set full summary of the top kernels, and traffic.json (DRAM
bytes per launch per pipeline stage, read by bench.py for roofline.traffic).
Usage: make_profiles.py <tag> [<round-label>] [<second tag> ...]
With several tags (one capture per robot of the C2 shape: the steps alternate between two scenes of
different cost) traffic.json holds the mean over the captures and the per-capture values."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
label = sys.argv[2] if len(sys.argv) > 2 else tag
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

# kernel -> bench.py stage name
STAGE = {"k_point_keys": "point_keys", "k_gather_sorted": "gather_sorted", "k_fold": "fold_wide",
         "k_bundle_order": "bundle_order", "k_walk_segments": "walk_segments",
         "k_block_accumulate": "block_accumulate", "k_voxel_update": "voxel_update",
         "k_long_finish": "replay_wide", "k_finalize_blocks": "finalize",
         "k_resample_merge": "merge_resample", "k_visit_precompute": "visits"}

head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                      text=True).stdout.strip()
ls = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_summary.py"),
                     os.path.join(go, f"launches_{tag}.csv")], capture_output=True, text=True).stdout
with open(os.path.join(out_dir, f"{label}_launches.md"), "w") as f:
    f.write(f"# {label}: launch list of C2 steps (commit {head})\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include cg_step/ "
            "-k 'regex:^k_|^Device' python bench.py --steps 2 --warmup 10 --profile-mode` — the NVTX "
            "range brackets exactly one un-instrumented C2 step (bench.py).\n\n"
            "Per-launch times under ncu are cold-cache and serialised: compare the SHARES with "
            "`stages_ms_per_step` of the bench line, not the absolutes.\n\n" + ls)

def raw_page(name):
    """--page raw --csv of a capture: the exported CSV if the report itself did not travel."""
    csv_path = os.path.join(go, f"prof_{name}.raw.csv")
    if os.path.exists(csv_path):
        return open(csv_path).read()
    return subprocess.run(["ncu", "-i", os.path.join(go, f"prof_{name}.ncu-rep"), "--page", "raw",
                           "--csv"], capture_output=True, text=True).stdout


raw = raw_page(tag)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__inst_executed.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def summarise(rows, hdr, units, traffic):
    ix = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in rows:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        short = name.split("<")[0]  # template arguments of the library kernels run to a page
        lines.append(f"## {short if short.startswith('cub::') else name}\n")
        name = short.replace("cg::", "")
        lines.append("| metric | value |\n|---|---|")
        for w in want:
            if w in ix:
                lines.append(f"| {w} | {r[ix[w]]} {units[ix[w]]} |")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[ix[h]]), h.replace("smsp__average_warps_issue_stalled_", "")
                                   .replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:4])
        lines.append(f"| top stalls (warps per issue) | {top} |\n")
        if traffic is None:
            continue
        dram = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
            to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        if name in STAGE and STAGE[name] not in traffic:
            traffic[STAGE[name]] = dram
        # library kernels, in launch order: the first radix sort of the step (histogram + its
        # onesweep passes) is the bundle sort, the second the (voxel, ray) pair sort; the first
        # select compacts the bundle heads, the second the update-list heads
        if "DeviceRadixSortHistogramKernel" in name:
            traffic["_sorts"] = traffic.get("_sorts", 0) + 1
        if "DeviceRadixSort" in name:
            key = "bundle_sort" if traffic.get("_sorts", 1) <= 1 else "pair_sort"
            traffic[key] = traffic.get(key, 0.0) + dram
        if "DeviceSelectSweepKernel" in name:
            traffic["_selects"] = traffic.get("_selects", 0) + 1
            key = "bundle_scan" if traffic["_selects"] == 1 else "segments"
            traffic[key] = traffic.get(key, 0.0) + dram
    return lines


traffic = {}
lines = summarise(rows[2:], hdr, units, traffic)
# second capture: the server-side kernels (40-submap projection, re-projection, meshing)
server = [os.path.join(go, f"prof_{k}_{tag}.ncu-rep") for k in ("merge", "mesh")]
if all(os.path.exists(r) for r in server):
    lines.append("# server side: `ncu --set full ... python scripts/merge_probe.py 40` — the first "
                 "projection of 40 C2 submaps (`-k regex:k_project_batch|k_mark_batch|k_rank_hist|"
                 "k_list_candidates -c 4`) and the meshing of the projected map (`-k k_mesh_blocks "
                 "-c 2`: count pass, write pass)\n")
    for rep2 in server:
        raw2 = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True,
                              text=True).stdout
        rows2 = list(csv.reader(io.StringIO(raw2)))
        lines += summarise(rows2[2:], rows2[0], rows2[1], None)
    plain = os.path.join(go, f"merge_plain_{tag}.log")
    if os.path.exists(plain):
        lines.append("Un-instrumented run of the same script:\n\n```\n" + open(plain).read().strip() +
                     "\n```\n")
with open(os.path.join(out_dir, f"{label}_ncu_top_kernels.md"), "w") as f:
    f.write(f"# {label}: ncu --set full of the library's own kernels (commit {head})\n\n"
            "`ncu --set full --clock-control none --import-source on --nvtx --nvtx-include cg_step/ "
            "-k regex:<kernels> python bench.py --steps 2 --warmup 10 --profile-mode` — the kernels of "
            "one C2 step (25 x 640x480 frames).\n\n" + "\n".join(lines))
traffic = {k: v for k, v in traffic.items() if not k.startswith("_")}
per_capture = {tag: traffic}
for extra in sys.argv[3:]:
    rows2 = list(csv.reader(io.StringIO(raw_page(extra))))
    t2 = {}
    extra_lines = summarise(rows2[2:], rows2[0], rows2[1], t2)
    per_capture[extra] = {k: v for k, v in t2.items() if not k.startswith("_")}
    with open(os.path.join(out_dir, f"{label}_ncu_top_kernels_{extra}.md"), "w") as f:
        f.write(f"# {label}: ncu --set full, capture {extra} (commit {head})\n\n" + "\n".join(extra_lines))
    ls2 = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_summary.py"),
                          os.path.join(go, f"launches_{extra}.csv")], capture_output=True, text=True).stdout
    with open(os.path.join(out_dir, f"{label}_launches_{extra}.md"), "w") as f:
        f.write(f"# {label}: launch list of one C2 step, capture {extra} (commit {head})\n\n" + ls2)
mean = {}
for k in sorted({k for t in per_capture.values() for k in t}):
    vals = [t[k] for t in per_capture.values() if k in t]
    mean[k] = sum(vals) / len(vals)
with open(os.path.join(out_dir, "traffic.json"), "w") as f:
    json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of one step per pipeline "
                           "stage (a stage = one kernel or one library call; the passes of a "
                           f"library sort are summed), ncu --set full, profiles/{label}_ncu_top_kernels*.md; "
                           "mean over the captures below (one per robot of the C2 shape)",
               **mean, "_per_capture": per_capture}, f, indent=1)
print("wrote", sorted(os.listdir(out_dir)))
