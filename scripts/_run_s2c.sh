# memcheck of the round's new kernels, then the profile refresh at the final code
# compute-sanitizer is closed on this GPU pool (runs under it left GPUs needing a reset)

SKIP_SERVER=1 STEPS=2 bash scripts/profile_round.sh r2_robot1
SKIP_SERVER=1 STEPS=3 bash scripts/profile_round.sh r2_robot0
ls -la gpurun_out/ | tail -20
