set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/c2_n2.json 2> gpurun_out/c2_n2.err; tail -c 1500 gpurun_out/c2_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config C3 --steps 10 --warmup 5 --c3-submaps 32 > gpurun_out/c3_n2.json 2> gpurun_out/c3_n2.err; tail -c 1500 gpurun_out/c3_n2.err
python - <<'PY'
import json
for c in ("c2_n2","c3_n2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{c}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(c,{k:d.get(k) for k in ("value","ms_per_step","per_rank_ms_per_step")}); print(" e2e",d.get("e2e"))
        p=d.get("project_submaps") or {}
        print(" sharded", p.get("sharded") or {k:p.get(k) for k in ("native","packed","parity","exchange","ms","value")})
    except Exception as ex: print("ERR",c,ex)
PY
