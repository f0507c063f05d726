timeout 200 python -m pytest tests/test_host_cpp.py tests/test_gpu_edge.py -m gpu -q 2>&1 | tail -2
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
    torch.cuda.synchronize(); print("pageable 25-frame batch: %.2f ms" % ((time.perf_counter()-t0)*1e3))
p1, c1 = pts[:307200].copy(), cols[:307200].copy()
for rep in range(3):
    L.clear(); torch.cuda.synchronize(); t0=time.perf_counter()
    for k in range(20): integ.integratePointCloud(poses[0], p1, c1)
    torch.cuda.synchronize(); print("pageable single frame: %.3f ms per call" % ((time.perf_counter()-t0)*1e3/20))
PY
