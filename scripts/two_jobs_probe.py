"""Two fusion jobs in flight on one GPU (two contexts, two host threads) against one at a time:
DESIGN.md §10(4).  The C ABI releases no lock of its own; ctypes drops the GIL during the calls."""
import sys, os, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig, mergeLayerAintoLayerB, synth
dev = torch.device("cuda", 0)
cfg = TsdfIntegratorConfig(**bench.CFG)
LANES = int(sys.argv[1]) if len(sys.argv) > 1 else 2
jobs = []
for robot in range(LANES):
    ctx = Context(0)
    sub, glob = Layer(ctx, 0.05, max_blocks=4096), Layer(ctx, 0.05, max_blocks=32768)
    ents = []
    for sm in range(4):
        poses, pts, cols = bench.host_frames(robot % 2, sm + 4 * (robot // 2), 25, dev)
        ents.append((poses, torch.cat(pts).contiguous(), torch.cat(cols).contiguous(),
                     np.cumsum([0] + [len(p) for p in pts]).astype(np.uint64)))
    jobs.append((ctx, sub, glob, TsdfIntegrator(cfg, sub), ents, synth.robot_map_offset(robot % 2)))
torch.cuda.synchronize()

def run(job, steps):
    ctx, sub, glob, integ, ents, T = job
    for k in range(steps):
        e = ents[k % len(ents)]
        sub.clear()
        integ.integrateBatch(e[0], e[1], e[2], e[3])
        mergeLayerAintoLayerB(sub, T, glob)
    ctx.synchronize()

for j in jobs:
    run(j, 6)                                     # warm-up: scratch buffers, key box
STEPS = 20
t0 = time.perf_counter()
for j in jobs:
    run(j, STEPS)
seq = time.perf_counter() - t0
th = [threading.Thread(target=run, args=(j, STEPS)) for j in jobs]
t0 = time.perf_counter()
for t in th: t.start()
for t in th: t.join()
par = time.perf_counter() - t0
pts = LANES * STEPS * 7.68e6
print(f"one job at a time: {seq/(LANES*STEPS)*1e3:.3f} ms per submap ({pts/seq/1e9:.2f} G points/s); "
      f"{LANES} in flight: {par/(LANES*STEPS)*1e3:.3f} ms per submap ({pts/par/1e9:.2f} G points/s)")
