timeout 600 python -m pytest tests/test_gpu_esdf.py tests/test_host_cpp.py -m gpu -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/s2b_c2.json 2> gpurun_out/s2b_c2.err; tail -c 400 gpurun_out/s2b_c2.err
timeout 500 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2b_c5.json 2> gpurun_out/s2b_c5.err; tail -c 400 gpurun_out/s2b_c5.err
python - <<'PY'
import json
for f in ("s2b_c2","s2b_c5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e = d.get("esdf") or d.get("project_submaps",{}).get("esdf")
        print(f, d["ms_per_step"], e)
    except Exception as ex: print("ERR", f, ex)
PY
ncu --set full --clock-control none --import-source on -k "regex:k_esdf_sweep" -c 3 -o gpurun_out/prof_esdf -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_esdf.log 2>&1
ncu -i gpurun_out/prof_esdf.ncu-rep --page raw --csv > gpurun_out/prof_esdf.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_esdf.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_esdf.src.csv 2>/dev/null; ls -la gpurun_out | tail -8
