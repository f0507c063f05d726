timeout 600 python -m pytest tests/test_gpu_parity_scale.py tests/test_gpu_multi.py tests/test_gpu_reproject.py tests/test_gpu_scale.py tests/test_mesh_recover.py -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --config C4 > gpurun_out/final_c4_n1.json 2> gpurun_out/final_c4.err; tail -c 300 gpurun_out/final_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_c4_n1.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}); print(d.get("e2e")); print(d.get("layer")); print({k:v for k,v in d["roofline"].items() if k!="note"})
PY
cat gpurun_out/parity_margins.json | head -80
