# session-2 state check: GPU tests, then the bench lines of C2 / C1 / C3 / C5 and the reference arm
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/final_c2_n1.json 2> gpurun_out/final_c2_n1.err; tail -c 300 gpurun_out/final_c2_n1.err
timeout 300 python bench.py --config C1 --steps 10 --warmup 3 > gpurun_out/final_c1_n1.json 2> gpurun_out/final_c1.err; tail -c 300 gpurun_out/final_c1.err
timeout 300 python bench.py --config C3 > gpurun_out/final_c3_n1.json 2> gpurun_out/final_c3.err; tail -c 300 gpurun_out/final_c3.err
timeout 500 python bench.py --config C5 --steps 5 --warmup 3 > gpurun_out/final_c5_n1.json 2> gpurun_out/final_c5.err; tail -c 300 gpurun_out/final_c5.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_c2_reference.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_c2_n1.json").read().strip().splitlines()[-1])
st=d["stages_ms_per_step"]
print("value ms", d["ms_per_step"], "plain", d["plain_calls"]["ms_per_step"], "per_frame", d["per_frame_call"], "e2e", d["e2e"])
print({k: round(v,3) for k,v in st.items()})
print({k:v for k,v in d["roofline"].items() if k!="note"})
PY
ls -la gpurun_out | head -30
