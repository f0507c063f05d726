timeout 400 python bench.py > gpurun_out/final_c2_n1.json 2> gpurun_out/final_c2_n1.err; tail -c 300 gpurun_out/final_c2_n1.err
timeout 300 python bench.py --config C1 --steps 10 --warmup 3 > gpurun_out/final_c1_n1.json 2> gpurun_out/final_c1.err; tail -c 300 gpurun_out/final_c1.err
timeout 300 python bench.py --config C3 > gpurun_out/final_c3_n1.json 2> gpurun_out/final_c3.err; tail -c 300 gpurun_out/final_c3.err
timeout 500 python bench.py --config C5 --steps 5 --warmup 3 > gpurun_out/final_c5_n1.json 2> gpurun_out/final_c5.err; tail -c 300 gpurun_out/final_c5.err
timeout 600 python bench.py --config C4 > gpurun_out/final_c4_n1.json 2> gpurun_out/final_c4.err; tail -c 300 gpurun_out/final_c4.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_c2_reference.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.err
python - <<'PY'
import json
for f in ("final_c2_n1","final_c1_n1","final_c3_n1","final_c5_n1","final_c4_n1","final_c2_reference"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("step_frac"))
    except Exception as ex: print("ERR", f, ex)
PY
