"""Stage timers of per-frame integratePointCloud calls (the live path)."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
dev = torch.device("cuda", 0)
ctx = Context(0)
cfg = TsdfIntegratorConfig(default_truncation_distance=0.16, use_const_weight=1, method=1)
fr = synth.submap_frames(0, 0, 25, device=dev)
L = Layer(ctx, 0.05, max_blocks=4096)
integ = TsdfIntegrator(cfg, L)
for rep in range(2):
    L.clear()
    if rep == 1:
        ctx.reset_profile(); ctx.set_profiling(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for (T, p, c) in fr:
        integ.integratePointCloud(T, p, c)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*1e3/25:.3f} ms per frame (wall)")
ctx.set_profiling(False)
prof = ctx.profile()
tot = sum(v[0] for v in prof.values())
print("sum of stages per frame: %.3f ms" % (tot / 25))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if v[0] > 0: print(f"  {k:18s} {v[0]/25*1e3:8.1f} us")
