#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): plain bench, ncu launch list of one step, ncu --set full
# of the library's own top kernels.  Outputs land in gpurun_out/ with the tag given as $1.
#   gpurun --timeout 1200 -- 'bash scripts/profile_round.sh r1'
tag=${1:-r1}
mkdir -p gpurun_out
# enough warm-up steps for the context's scratch buffers and bundle-key box to settle
# STEPS=2 ends on robot 1's submap, STEPS=3 on robot 0's (the steps alternate between the robots)
ARGS="--steps ${STEPS:-2} --warmup 10 --profile-mode"
if [ -z "$SKIP_STEP" ]; then
python bench.py $ARGS > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "cg_step/" -k "regex:^k_|^Device" --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k "regex:k_block_accumulate|k_walk_segments|k_fold|k_point_keys|k_voxel_update|k_visit_precompute|k_long_finish|k_gather_sorted|k_resample_merge|k_bundle_order|k_finalize_blocks|DeviceRadixSortOnesweepKernel|DeviceRadixSortHistogramKernel|DeviceSelectSweepKernel" \
    --nvtx --nvtx-include "cg_step/" -o gpurun_out/prof_$tag -f python bench.py $ARGS > gpurun_out/ncu_full_$tag.log 2>&1
# only the raw page travels back (gpurun_out/ is limited to 64 MiB): KEEP_REP=1 keeps the report
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$tag.raw.csv 2>/dev/null
[ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_$tag.ncu-rep
fi
if [ -n "$SKIP_SERVER" ]; then echo profile_round done "(step only)"; exit 0; fi
# server side: projection of 40 submaps, incremental re-projection, meshing (scripts/merge_probe.py)
python scripts/merge_probe.py 40 > gpurun_out/merge_plain_$tag.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k "regex:k_project_batch|k_mark_batch|k_rank_hist|k_list_candidates" \
    -c 4 -o gpurun_out/prof_merge_$tag -f python scripts/merge_probe.py 40 > gpurun_out/ncu_merge_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_mesh_blocks" \
    -c 2 -o gpurun_out/prof_mesh_$tag -f python scripts/merge_probe.py 40 > gpurun_out/ncu_mesh_$tag.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$tag.csv
echo profile_round done
