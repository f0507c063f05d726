N=${1:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --config C3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/final_c3_n$N.json 2> gpurun_out/c3_n$N.err; tail -c 600 gpurun_out/c3_n$N.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_c2_n$N.json 2> gpurun_out/c2_n$N.err; tail -c 600 gpurun_out/c2_n$N.err
python - <<PY
import json
for c in ("final_c2_n$N","final_c3_n$N"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{c}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(c,{k:d.get(k) for k in ("value","ms_per_step","per_rank_ms_per_step")}); print(" e2e",d.get("e2e"))
        p=d.get("project_submaps") or {}
        print(" sharded", p.get("sharded") or {k:p.get(k) for k in ("native","packed","parity","exchange","ms","value","submaps_total")})
    except Exception as ex: print("ERR",c,ex)
PY
