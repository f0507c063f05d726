ARGS="--steps 3 --warmup 10 --profile-mode"
python bench.py $ARGS > gpurun_out/plain_b.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_b.log; exit 1; }
ncu --set full --clock-control none --import-source on \
    -k "regex:k_voxel_update|k_long_finish|k_visit_precompute" \
    --nvtx --nvtx-include "cg_step/" -o gpurun_out/prof_b -f python bench.py $ARGS > gpurun_out/ncu_full_b.log 2>&1
tail -3 gpurun_out/ncu_full_b.log
