set -x
timeout 300 python bench.py --config C1 --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/c1.json 2> gpurun_out/c1.err; tail -c 800 gpurun_out/c1.err
timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --c3-submaps 16 --cpu-seconds 5 > gpurun_out/c3.json 2> gpurun_out/c3.err; tail -c 800 gpurun_out/c3.err
timeout 300 python bench.py --config C5 --steps 5 --warmup 3 --c3-submaps 8 --cpu-seconds 5 > gpurun_out/c5.json 2> gpurun_out/c5.err; tail -c 800 gpurun_out/c5.err
timeout 300 python bench.py --config C4 --c4-blocks 60000 --cpu-seconds 5 > gpurun_out/c4.json 2> gpurun_out/c4.err; tail -c 800 gpurun_out/c4.err
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 3 > gpurun_out/c2.json 2> gpurun_out/c2.err; tail -c 800 gpurun_out/c2.err
for c in c1 c2 c3 c4 c5; do echo == $c; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$c.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("metric","value","ms_per_step","gpu_launches") if k in d})
    print("e2e",d.get("e2e")); print("roofline",{k:v for k,v in d["roofline"].items() if k!="note"})
    for k in ("layer","project_submaps","per_step","cpu_baseline","reproject_10pct_moved"):
        if k in d: print(k, d[k])
except Exception as ex: print("ERR", ex)
PY
done
