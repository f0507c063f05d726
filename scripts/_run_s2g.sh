timeout 900 python -m pytest tests/test_gpu_integrate.py tests/test_gpu_edge.py tests/test_golden.py tests/test_gpu_parity_scale.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --project-submaps 0 > gpurun_out/r2_q.json 2> gpurun_out/r2_q.err; tail -c 300 gpurun_out/r2_q.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_q.json").read().strip().splitlines()[-1])
st=d["stages_ms_per_step"]
print("value ms", d["ms_per_step"], "plain", d["plain_calls"]["ms_per_step"], "per_frame", d["per_frame_call"]["ms"], d["per_frame_call"].get("queued_ms"), "e2e", d["e2e"]["ms_per_step"])
print({k: round(v,3) for k,v in st.items()})
print(d["two_jobs_in_flight"])
PY
