timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/r2_p.json 2> gpurun_out/r2_p.err; tail -c 600 gpurun_out/r2_p.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_p.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","plain_calls","e2e","e2e_loops_ms_per_step","per_frame_call","gpu_launches","stages_ms_per_step"): print(k, d.get(k))
print({k:v for k,v in d["roofline"].items() if k!="note"})
PY
