# whole GPU suite (new parity cases included) + the C2 line with the queued per-frame leg
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
timeout 400 python bench.py > gpurun_out/final_c2_n1.json 2> gpurun_out/final_c2_n1.err; tail -c 300 gpurun_out/final_c2_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_c2_n1.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "plain", d["plain_calls"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print("per_frame", {k:v for k,v in d["per_frame_call"].items() if k!="host_pageable"})
print("esdf", d["project_submaps"]["esdf"])
PY
