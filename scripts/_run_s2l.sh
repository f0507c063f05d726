timeout 1100 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/final_c2_n1.json 2> gpurun_out/final_c2_n1.err; tail -c 300 gpurun_out/final_c2_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_c2_n1.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "plain", d["plain_calls"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["step_frac"])
print("per_frame", d["per_frame_call"]["ms"], d["per_frame_call"]["queued_ms"], d["per_frame_call"]["host_pageable"])
PY
