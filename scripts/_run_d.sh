timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --config C4 --cpu-seconds 10 > gpurun_out/c4_full.json 2> gpurun_out/c4_full.err; tail -c 600 gpurun_out/c4_full.err
timeout 600 python bench.py --config C5 --steps 5 --warmup 3 --cpu-seconds 10 > gpurun_out/c5_full.json 2> gpurun_out/c5_full.err; tail -c 600 gpurun_out/c5_full.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
