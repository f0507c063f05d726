import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coxgraph_b200 import Context, Layer, TsdfIntegrator, TsdfIntegratorConfig, mergeLayerAintoLayerB, synth, VOXEL_DTYPE
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = Context(0, stream=st.cuda_stream)
cfg = TsdfIntegratorConfig(default_truncation_distance=0.16, use_const_weight=1, method=1)
fr = synth.submap_frames(0, 0, 25, device=dev)
P = np.stack([T for (T, _, _) in fr]).astype(np.float32)
pts = torch.cat([p for (_, p, _) in fr]).contiguous(); cols = torch.cat([c for (_, _, c) in fr]).contiguous()
offs = np.cumsum([0] + [len(p) for (_, p, _) in fr]).astype(np.uint64)
sub, glob = Layer(ctx, 0.05, max_blocks=4096), Layer(ctx, 0.05, max_blocks=32768)
integ = TsdfIntegrator(cfg, sub)
out_idx = torch.empty((4096, 3), dtype=torch.int32, pin_memory=True).numpy()
out_vox = torch.empty((4096, 4096 * 12), dtype=torch.uint8, pin_memory=True).numpy().view(VOXEL_DTYPE).reshape(4096, 4096)
out_flags = torch.empty((4096,), dtype=torch.uint8, pin_memory=True).numpy()
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("clear            %.3f ms" % t(lambda: sub.clear()))
print("clear+integrate  %.3f ms" % t(lambda: (sub.clear(), integ.integrateBatch(P, pts, cols, offs))))
print("merge            %.3f ms" % t(lambda: mergeLayerAintoLayerB(sub, synth.robot_map_offset(1), glob)))
print("download pinned  %.3f ms (%d blocks)" % (t(lambda: sub.download(out=(out_idx, out_vox, out_flags))), sub.num_blocks))
