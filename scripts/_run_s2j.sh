CG_TRACE_COMM=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --config C3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/trace_c3_n2.json 2> gpurun_out/trace_c3_n2.err
grep "cg comm rank 0" gpurun_out/trace_c3_n2.err | tail -4
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/trace_c3_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
p=d.get("project_submaps") or {}
print({k:p.get(k) for k in ("native","packed","ms","submaps","global_blocks","blocks_in","blocks_out")})
PY
