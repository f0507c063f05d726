timeout 600 python -m pytest tests -m gpu -x -q tests/test_gpu_integrate.py tests/test_golden.py 2>&1 | tail -3
for c in 4 8; do
CG_FOLD_SHORT_CTAS=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --project-submaps 0 > gpurun_out/r2_q$c.json 2> gpurun_out/r2_q.err; tail -c 300 gpurun_out/r2_q.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_q$c.json").read().strip().splitlines()[-1])
st=d["stages_ms_per_step"]
print($c, "value ms", d["ms_per_step"], "plain", d["plain_calls"]["ms_per_step"], "fold_wide", st["fold_wide"], "fold", st["fold_bundles"], "order", st["bundle_order"], "per_frame", d["per_frame_call"]["ms"])
print(d["two_jobs_in_flight"])
PY
done
