"""Timing probe of the server-side global merge (getProjectedMap) at the C2 / C5 shapes."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coxgraph_b200 import (Context, Layer, TsdfIntegrator, TsdfIntegratorConfig, getProjectedMap,
                           reprojectSubmaps, synth)

dev = torch.device("cuda", 0)
ctx = Context(0)
cfg = TsdfIntegratorConfig(default_truncation_distance=0.16, use_const_weight=1, method=1)
subs, poses = [], []
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for k in range(nsub):
    robot, sm = k % 2, k // 2
    fr = synth.submap_frames(robot, sm % 20, 25, device=dev)
    L = Layer(ctx, 0.05, max_blocks=1024)
    P = np.stack([T for (T, _, _) in fr]).astype(np.float32)
    pts = torch.cat([p for (_, p, _) in fr]).contiguous()
    cols = torch.cat([c for (_, _, c) in fr]).contiguous()
    offs = np.cumsum([0] + [len(p) for (_, p, _) in fr]).astype(np.uint64)
    TsdfIntegrator(cfg, L).integrateBatch(P, pts, cols, offs)
    subs.append(L)
    poses.append(synth.robot_map_offset(robot))
    del fr, pts, cols
glob = Layer(ctx, 0.05, max_blocks=65536)
poses = np.stack(poses)
blocks_in = sum(L.num_blocks for L in subs)
for it in range(3):
    glob.removeAllBlocks()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    getProjectedMap(subs, poses, glob)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if it == 2:
        ctx.reset_profile(); ctx.set_profiling(True)
        glob.removeAllBlocks(); getProjectedMap(subs, poses, glob); ctx.set_profiling(False)
        print({k: round(v[0], 3) for k, v in ctx.profile().items() if v[0] > 0})
    print(f"project {nsub} submaps ({blocks_in} blocks in, {glob.num_blocks} global): {dt*1e3:.2f} ms, "
          f"{4096*blocks_in/dt/1e9:.2f} G voxels/s, {dt*1e3/nsub:.3f} ms/submap")

# incremental re-projection after a pose-graph update that moved a fraction of the submaps
rng = np.random.default_rng(1)
for frac in (0.05, 0.25, 1.0):
    best = None
    for it in range(3):
        glob.removeAllBlocks()
        getProjectedMap(subs, poses, glob)
        new = poses.copy()
        moved = rng.choice(nsub, max(1, int(frac * nsub)), replace=False)
        for k in moved:
            new[k] = synth.perturb_pose(poses[k], rng)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        changed, st = reprojectSubmaps(subs, poses, new, glob)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"reproject {len(moved)}/{nsub} moved: {best*1e3:.2f} ms (dirty {st.blocks_dirty} blocks, "
          f"{st.candidates} candidates, {st.blocks_folded} folds, {st.blocks_removed} removed)")

# meshing the device-resident global map (N4) against downloading it
glob.removeAllBlocks()
getProjectedMap(subs, poses, glob)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mi, mb, mv, mn, mc = glob.generateMesh()
    dt = time.perf_counter() - t0
t0 = time.perf_counter()
glob.download()
dl = time.perf_counter() - t0
print(f"generateMesh: {glob.num_blocks} blocks -> {len(mv)//3} triangles in {dt*1e3:.2f} ms "
      f"({(len(mv) * 28) / 1e6:.1f} MB out); downloading the layer instead: {dl*1e3:.2f} ms "
      f"({glob.num_blocks * 49152 / 1e6:.1f} MB)")
