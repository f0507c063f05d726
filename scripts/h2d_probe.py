"""Host->device rate of the staged transfer alone and overlapped with the fusion (bench.py's e2e leg)."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = Context(0, stream=st.cuda_stream)
cfg = TsdfIntegratorConfig(**bench.CFG)
sub = Layer(ctx, 0.05, max_blocks=4096)
integ = TsdfIntegrator(cfg, sub)
ents = []
for s in range(4):
    poses, pts, cols = bench.host_frames(s % 2, s, 25, dev)
    d_pts, d_cols = torch.cat(pts).contiguous(), torch.cat(cols).contiguous()
    offs = np.cumsum([0] + [len(p) for p in pts]).astype(np.uint64)
    h_pts = torch.empty(d_pts.shape, dtype=d_pts.dtype, pin_memory=True).copy_(d_pts).numpy()
    h_cols = torch.empty(d_cols.shape, dtype=d_cols.dtype, pin_memory=True).copy_(d_cols).numpy()
    ents.append((poses, offs, h_pts, h_cols, d_pts, d_cols))
torch.cuda.synchronize()
nbytes = ents[0][2].nbytes + ents[0][3].nbytes
# (a) staging alone
for rep in range(2):
    ctx.synchronize(); t0 = time.perf_counter()
    for k in range(20):
        e = ents[k % 4]
        integ.stageBatch(k % 2, e[2], e[3])
    ctx.synchronize(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"stage alone: {dt*1e3:.3f} ms per {nbytes/1e6:.1f} MB = {nbytes/dt/1e9:.1f} GB/s")
# (b) fusion alone from staged slots (no copy in flight)
integ.stageBatch(0, ents[0][2], ents[0][3]); ctx.synchronize(); torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(20):
    sub.clear(); integ.integrateStaged(0, ents[0][0], ents[0][1])
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"fusion alone (staged input): {dt*1e3:.3f} ms")
# (c) overlapped as in bench.run_e2e, without merge / download
integ.stageBatch(0, ents[0][2], ents[0][3])
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(20):
    nxt = ents[(k + 1) % 4]
    integ.stageBatch((k + 1) % 2, nxt[2], nxt[3])
    sub.clear(); e = ents[k % 4]
    integ.integrateStaged(k % 2, e[0], e[1])
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"overlapped: {dt*1e3:.3f} ms per step = {nbytes/dt/1e9:.1f} GB/s of input")
# (d) the same with the library's stage timers on: GPU time of the stages vs wall time per step
for mode in ("alone", "overlapped"):
    ctx.reset_profile(); ctx.set_profiling(True)
    integ.stageBatch(0, ents[0][2], ents[0][3])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(20):
        nxt = ents[(k + 1) % 4]
        if mode == "overlapped":
            integ.stageBatch((k + 1) % 2, nxt[2], nxt[3])
        sub.clear(); e = ents[k % 4]
        integ.integrateStaged((k % 2) if mode == "overlapped" else 0, ents[0][0] if mode == "alone" else e[0],
                              ents[0][1] if mode == "alone" else e[1])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    ctx.set_profiling(False)
    prof = ctx.profile()
    tot = sum(v[0] for v in prof.values()) / 20
    top = sorted(((v[0] / 20, k) for k, v in prof.items()), reverse=True)[:6]
    print(f"{mode}: wall {dt*1e3:.3f} ms per step, stage sum {tot:.3f} ms; " +
          ", ".join(f"{k} {v:.3f}" for v, k in top))
