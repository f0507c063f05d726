for t in 0 3 0 3 6; do echo "CG_STAGE_THREADS=$t"; CG_STAGE_THREADS=$t ./build/host_api_check time 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_integrate.py tests/test_host_cpp.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import time, numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
import torch
ctx = Context(0); L = Layer(ctx, 0.05, max_blocks=4096)
integ = TsdfIntegrator(TsdfIntegratorConfig(default_truncation_distance=0.16, use_const_weight=1, method=1), L)
fr = synth.submap_frames(0, 0, 25, device=torch.device("cuda", 0))
poses = np.stack([T for (T,_,_) in fr]).astype(np.float32)
pts = torch.cat([p for (_,p,_) in fr]).cpu().numpy().copy(); cols = torch.cat([c for (_,_,c) in fr]).cpu().numpy().copy()
offs = np.cumsum([0]+[len(p) for (_,p,_) in fr]).astype(np.uint64)
for rep in range(4):
    L.clear(); torch.cuda.synchronize(); t0=time.perf_counter()
    integ.integrateBatch(poses, pts, cols, offs)
    torch.cuda.synchronize(); print("pageable 25-frame batch (123 MB): %.2f ms" % ((time.perf_counter()-t0)*1e3), "threads", os.environ.get("CG_STAGE_THREADS","3"))
PY
CG_STAGE_THREADS=0 python - <<'PY'
import time, numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
import torch
ctx = Context(0); L = Layer(ctx, 0.05, max_blocks=4096)
integ = TsdfIntegrator(TsdfIntegratorConfig(default_truncation_distance=0.16, use_const_weight=1, method=1), L)
fr = synth.submap_frames(0, 0, 25, device=torch.device("cuda", 0))
poses = np.stack([T for (T,_,_) in fr]).astype(np.float32)
pts = torch.cat([p for (_,p,_) in fr]).cpu().numpy().copy(); cols = torch.cat([c for (_,_,c) in fr]).cpu().numpy().copy()
offs = np.cumsum([0]+[len(p) for (_,p,_) in fr]).astype(np.uint64)
for rep in range(4):
    L.clear(); torch.cuda.synchronize(); t0=time.perf_counter()
    integ.integrateBatch(poses, pts, cols, offs)
    torch.cuda.synchronize(); print("pageable 25-frame batch (123 MB): %.2f ms" % ((time.perf_counter()-t0)*1e3), "threads 0")
PY
