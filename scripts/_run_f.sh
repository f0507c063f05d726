SKIP_SERVER=1 STEPS=2 bash scripts/profile_round.sh r2_robot1
SKIP_SERVER=1 STEPS=3 bash scripts/profile_round.sh r2_robot0
ls -la gpurun_out/
