"""Per-frame integratePointCloud calls over the submaps rank `R` of `W` fuses in bench.py: wall time
per frame per submap, with CG_TRACE_KEYS=1 the bundle-key box retries show on stderr."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig
R, W = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
ctx = Context(0)
cfg = TsdfIntegratorConfig(**bench.CFG)
L = Layer(ctx, 0.05, max_blocks=4096)
integ = TsdfIntegrator(cfg, L)
for s in range(12):
    robot, sm = bench.submap_of_step(s, R, W)
    poses, pts, cols = bench.host_frames(robot, sm, 25, dev)
    L.clear()
    ctx.reset_profile(); ctx.set_profiling(True)
    n0 = ctx.kernel_launches
    per = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for f in range(25):
        t1 = time.perf_counter()
        st = integ.integratePointCloud(poses[f], pts[f], cols[f])
        per.append((time.perf_counter() - t1) * 1e3)
        if per[-1] > 5:
            print(f"   slow frame {f}: {per[-1]:.1f} ms rays {st.rays} updates {st.voxel_updates} "
                  f"general {st.general_updates} touched {st.blocks_touched}", flush=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ctx.set_profiling(False)
    prof = ctx.profile()
    print(f"step {s} robot {robot} submap {sm}: {dt*1e3/25:.3f} ms per frame, stage sum "
          f"{sum(v[0] for v in prof.values())/25:.3f} ms, {(ctx.kernel_launches - n0)/25:.1f} launches/frame, "
          f"slowest frames {sorted(np.round(per, 2))[-3:]}, blocks {L.num_blocks}", flush=True)
