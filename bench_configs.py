"""bench_configs.py — the BASELINE.json configs other than the headline C2 (`bench.py --config`).

  C1  single client: 100 synthetic 640x480 frames fused into one 5 cm submap (trunc 0.15 m).
      step = clear the layer + one 100-frame job (30.72 M points).
  C3  8 robots x 64 submaps x 10 frames, one robot per GPU (`--gpus 8`; at fewer GPUs a rank
      takes robot = rank), unbounded corridor scene so that the submaps lie along trajectories.
      step = clear + fuse one 10-frame submap + merge it into the rank's global layer; afterwards
      the server's global merge of all submaps (sharded over the ranks when N > 1).
  C4  2 cm voxels, trunc 0.06 m, max_ray 3 m, 1280x720, long corridor trajectory until ~2 M blocks
      are allocated (a 98 GB block pool).  step = one 4-frame job into the growing layer; the
      steps alternate between device-resident inputs (-> value) and host buffers (-> e2e).
  C5  the 512 submaps of C3 (8 robots x 64) on one GPU, every pose perturbed (sigma_t 5 cm,
      sigma_yaw 1 deg): step = the full re-projection into an empty global layer.

Same JSON contract as bench.py (one line on rank 0).  Synthetic inputs are rendered with torch on
the GPU outside every timed region; timing is CUDA events on the context's stream, max over ranks.
"""
import json
import os
import time

import numpy as np

import bench as B

C1_CFG = dict(default_truncation_distance=0.15, max_ray_length_m=5.0, min_ray_length_m=0.1,
              use_const_weight=1, method=1)
C3_CFG = dict(default_truncation_distance=0.16, max_ray_length_m=5.0, min_ray_length_m=0.1,
              use_const_weight=1, method=1)
C4_CFG = dict(default_truncation_distance=0.06, max_ray_length_m=3.0, min_ray_length_m=0.1,
              use_const_weight=1, method=1)


class Env:
    """Rank plumbing shared by the configs."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from coxgraph_b200 import Context
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx = Context(local_rank, stream=self.stream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def close(self):
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def make_entry(env, frames, T_M_S=None, pinned=True):
    """[(pose, points, colours)] (device tensors) -> one job's inputs, resident in HBM and in
    pinned host memory."""
    torch = env.torch
    poses = np.stack([T for (T, _, _) in frames]).astype(np.float32)
    d_pts = torch.cat([p for (_, p, _) in frames]).contiguous()
    d_cols = torch.cat([c for (_, _, c) in frames]).contiguous()
    offs = np.cumsum([0] + [len(p) for (_, p, _) in frames]).astype(np.uint64)
    e = dict(poses=poses, d_pts=d_pts, d_cols=d_cols, offs=offs, n=int(offs[-1]), T_M_S=T_M_S,
             frames=len(frames))
    if pinned:
        e["h_pts"] = torch.empty(d_pts.shape, dtype=d_pts.dtype, pin_memory=True).copy_(d_pts).numpy()
        e["h_cols"] = torch.empty(d_cols.shape, dtype=d_cols.dtype, pin_memory=True).copy_(d_cols).numpy()
    return e


def account(env, integ, layer, e, glob=None):
    """Untimed pass: per-frame B_touched (the byte model's definition), job statistics, merge
    block counts."""
    from coxgraph_b200 import mergeLayerAintoLayerB
    layer.clear()
    touched = 0
    for f in range(e["frames"]):
        a, b = int(e["offs"][f]), int(e["offs"][f + 1])
        touched += integ.integratePointCloud(e["poses"][f], e["d_pts"][a:b], e["d_cols"][a:b]).blocks_touched
    e["bytes_integrate"] = 16 * e["n"] + 2 * B.BLOCK_BYTES * touched
    layer.clear()
    st = integ.integrateBatch(e["poses"], e["d_pts"], e["d_cols"], e["offs"])
    e.update(rays=int(st.rays), pairs=int(st.voxel_updates), general=int(st.general_updates),
             blocks_in=layer.num_blocks, bytes_merge=0, voxels_in=0, b_out=0)
    if glob is not None:
        glob.clear()
        ms = mergeLayerAintoLayerB(layer, e["T_M_S"], glob)
        e["bytes_merge"] = B.BLOCK_BYTES * (ms.blocks_in + 2 * ms.blocks_out)
        e["voxels_in"] = 4096 * ms.blocks_in
        e["b_out"] = int(ms.blocks_out)


def cpu_baseline(e, cfg_fields, voxel, seconds):
    """The oracle's voxblox-style multi-thread integrator on the host cores, bounded."""
    from oracle import oracle_py as orc
    threads = os.cpu_count() or 1
    ocfg = orc.default_config(**cfg_fields)
    ol = orc.Layer(voxel)
    t0 = time.perf_counter()
    done = frames = 0
    for f in range(e["frames"]):
        a, b = int(e["offs"][f]), int(e["offs"][f + 1])
        ol.integrate(ocfg, e["poses"][f], e["h_pts"][a:b], e["h_cols"][a:b], threads=threads)
        done += b - a
        frames += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "points/s", "cores": threads, "kind": "port",
            "sample": f"first {frames} frames ({done} points) of one job, {threads}-thread "
                      f"voxblox-style integrator (oracle port)"}


def fusion_loop(env, spec, pool, submap, glob, integ):
    """The three legs of a fusion config: device-resident, profiled, host buffers."""
    from coxgraph_b200 import mergeLayerAintoLayerB
    args, torch, ctx, stream = env.args, env.torch, env.ctx, env.stream
    pool_n = len(pool)
    cap = submap.max_blocks
    from coxgraph_b200 import VOXEL_DTYPE
    out_idx = torch.empty((cap, 3), dtype=torch.int32, pin_memory=True).numpy()
    out_vox = torch.empty((cap, 4096 * 12), dtype=torch.uint8, pin_memory=True).numpy() \
        .view(VOXEL_DTYPE).reshape(cap, 4096)
    out_flags = torch.empty((cap,), dtype=torch.uint8, pin_memory=True).numpy()

    def step_device(e):
        submap.clear()
        integ.integrateBatch(e["poses"], e["d_pts"], e["d_cols"], e["offs"])
        if glob is not None:
            mergeLayerAintoLayerB(submap, e["T_M_S"], glob)

    def run_e2e(entries):
        d2h = 0
        if entries:
            integ.stageBatch(0, entries[0]["h_pts"], entries[0]["h_cols"])
        for k, e in enumerate(entries):
            if k + 1 < len(entries):
                nxt = entries[k + 1]
                integ.stageBatch((k + 1) % 2, nxt["h_pts"], nxt["h_cols"])
            submap.clear()
            integ.integrateStaged(k % 2, e["poses"], e["offs"])
            if glob is not None:
                mergeLayerAintoLayerB(submap, e["T_M_S"], glob)
            idx, _, _ = submap.download(out=(out_idx, out_vox, out_flags))
            d2h += len(idx) * (B.BLOCK_BYTES + 13)
        return d2h

    results = {}
    for leg in ("device", "profiled", "e2e"):
        if glob is not None:
            glob.clear()
        if leg != "e2e":
            for s in range(args.warmup):
                step_device(pool[s % pool_n])
        else:
            run_e2e([pool[s % pool_n] for s in range(args.warmup)])
        sampler = B.ClockSampler(env.local_rank)
        ev_a, ev_b = env.event(), env.event()
        launches0 = ctx.kernel_launches
        if leg == "profiled":
            ctx.reset_profile()
            ctx.set_profiling(True)
        sampler.start()
        env.barrier()
        ev_a.record(stream)
        d2h = 0
        used = [pool[(args.warmup + k) % pool_n] for k in range(args.steps)]
        if leg != "e2e":
            for e in used:
                step_device(e)
        else:
            d2h = run_e2e(used)
        ev_b.record(stream)
        env.barrier()
        clocks = sampler.stop()
        ctx.set_profiling(False)
        res = dict(total_ms=env.reduce(ev_a.elapsed_time(ev_b)), clocks=clocks,
                   launches=ctx.kernel_launches - launches0,
                   points=env.reduce(sum(e["n"] for e in used), "sum"),
                   h2d=sum(16 * e["n"] for e in used) / args.steps, d2h=d2h / args.steps,
                   bytes=sum(e["bytes_integrate"] + e["bytes_merge"] for e in used), used=used)
        if leg == "profiled":
            res["profile"] = ctx.profile()
        results[leg] = res
    return results


def fusion_line(env, spec, results, extra):
    args = env.args
    dv, ee, prof = results["device"], results["e2e"], results["profiled"]["profile"]
    used = dv["used"]
    n = float(len(used))
    per_step = dict(n_pts=sum(e["n"] for e in used) / n, rays=sum(e["rays"] for e in used) / n,
                    pairs=sum(e["pairs"] for e in used) / n,
                    general=sum(e["general"] for e in used) / n,
                    blocks=sum(e["blocks_in"] for e in used) / n,
                    b_in=sum(e["blocks_in"] for e in used) / n if spec["merge"] else 0,
                    b_out=sum(e["b_out"] for e in used) / n)
    line = {
        "metric": "tsdf_points_integrated_per_s", "value": dv["points"] / (dv["total_ms"] * 1e-3),
        "unit": "points/s", "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dv["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": spec["workload"], "parallelism": spec["parallelism"],
                   "l2": "inputs of one step exceed the 126 MB L2; distinct job per step",
                   "pool_jobs": len({id(e) for e in used})},
        "e2e": {"value": ee["points"] / (ee["total_ms"] * 1e-3), "unit": "points/s",
                "ms_per_step": ee["total_ms"] / args.steps, "h2d_bytes_per_step": ee["h2d"],
                "d2h_bytes_per_step": ee["d2h"]},
        "gpu_launches": dv["launches"],
        "roofline": B.roofline_record(prof, args.steps, dv["total_ms"] / args.steps,
                                      dv["bytes"] / args.steps, per_step, traffic_file=None),
        "stages_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
        "per_step": {k: per_step[k] for k in ("rays", "pairs", "general", "blocks")},
        "clocks": dv["clocks"], "clocks_e2e": ee["clocks"],
    }
    line.update(extra)
    return line


# ----------------------------------------------------------------------------- C1
def run_c1(env):
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    args = env.args
    spec = dict(workload="C1 single client: per step 100 x 640x480 depth frames (30.72 M points) fused "
                         "into one cleared 5 cm submap, 15 cm truncation, merged integrator, carving",
                parallelism=f"one client per GPU, {env.world} GPU(s), no collective", merge=False)
    submap = Layer(env.ctx, 0.05, max_blocks=8192)
    integ = TsdfIntegrator(TsdfIntegratorConfig(**C1_CFG), submap)
    pool = []
    for k in range(2):
        fr = synth.submap_frames((env.rank + k) % 2, k, 100, device=env.dev)
        pool.append(make_entry(env, fr))
        account(env, integ, submap, pool[-1])
    results = fusion_loop(env, spec, pool, submap, None, integ)
    extra = {}
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        extra["cpu_baseline"] = cpu_baseline(pool[0], C1_CFG, 0.05, args.cpu_seconds)
    if env.rank == 0:
        print(json.dumps(fusion_line(env, spec, results, extra)), flush=True)
    submap.close()


# ----------------------------------------------------------------------------- C3 / C5 submaps
def corridor_submap(env, robot, sm, frames=10, advance=0.2, with_host=False):
    """Submap `sm` of `robot` on the corridor scene: robots start 128 m apart, a submap covers
    frames * advance metres, consecutive submaps follow one another along the trajectory."""
    from coxgraph_b200 import synth
    fr = synth.corridor_frames(sm * frames, frames, robot=robot, advance=advance,
                               start_x=128.0 * robot, device=env.dev)
    return make_entry(env, fr, T_M_S=synth.robot_map_offset(0), pinned=with_host)


def fuse_submaps(env, robot_submaps, cfg_fields, max_blocks=768):
    """-> [Layer] (one fused submap each), resident on the device."""
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig
    cfg = TsdfIntegratorConfig(**cfg_fields)
    out = []
    for (robot, sm) in robot_submaps:
        e = corridor_submap(env, robot, sm)
        L = Layer(env.ctx, 0.05, max_blocks=max_blocks)
        TsdfIntegrator(cfg, L).integrateBatch(e["poses"], e["d_pts"], e["d_cols"], e["offs"])
        out.append(L)
    return out


def time_projection(env, subs, poses, glob, repeats=4):
    from coxgraph_b200 import getProjectedMap
    times, st = [], None
    for it in range(repeats):
        glob.clear()
        a, b = env.event(), env.event()
        a.record(env.stream)
        st = getProjectedMap(subs, poses, glob, want_stats=(it == repeats - 1))
        b.record(env.stream)
        env.torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return min(times[1:]), st


def run_c3(env):
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    args = env.args
    robot = env.rank % 8
    spec = dict(workload="C3 eight robots x 64 submaps x 10 frames (640x480, 3.07 M points per "
                         "step), one robot per GPU, corridor scene, 5 cm voxels, 16 cm truncation; "
                         "per step one submap is fused into a cleared layer and merged into the "
                         "rank's global TSDF; then the server's global merge of all submaps",
                parallelism=f"robot r -> rank r mod {env.world}; global merge sharded by block owner "
                            f"over {env.world} GPU(s)", merge=True)
    submap = Layer(env.ctx, 0.05, max_blocks=2048)
    glob = Layer(env.ctx, 0.05, max_blocks=131072)
    integ = TsdfIntegrator(TsdfIntegratorConfig(**C3_CFG), submap)
    pool = []
    for k in range(min(args.pool, args.warmup + args.steps)):
        pool.append(corridor_submap(env, robot, 3 * k, with_host=True))
        account(env, integ, submap, pool[-1], glob)
    results = fusion_loop(env, spec, pool, submap, glob, integ)
    # the server's global merge: all submaps of this rank's robot(s) into the global layer
    mine = [(robot, sm) for sm in range(args.c3_submaps)]  # one robot per GPU (weak scaling)
    subs = fuse_submaps(env, mine, C3_CFG)
    rng = np.random.default_rng(1000 + env.rank)
    poses = np.stack([synth.perturb_pose(synth.robot_map_offset(0), rng) for _ in subs])
    glob.clear()
    vox = env.reduce(4096.0 * sum(L.num_blocks for L in subs), "sum")
    project = {"submaps_total": len(subs) * env.world, "unit": "voxels/s"}
    if env.world == 1:
        ms, st = time_projection(env, subs, poses, glob)
        project.update(ms=ms, value=vox / (ms * 1e-3), blocks_in=int(st.blocks_in),
                       blocks_out=int(st.blocks_out), global_blocks=glob.num_blocks,
                       hbm_frac=B.BLOCK_BYTES * (st.blocks_in + 2 * st.blocks_out) / (ms * 1e-3) / 1e9 /
                       B.measured_peak_gbs()[0])
    else:
        partial = Layer(env.ctx, 0.05, max_blocks=131072)
        owned = Layer(env.ctx, 0.05, max_blocks=131072)
        project.update(B.time_sharded(subs, poses, partial, owned, env.ctx, env.barrier,
                                      lambda x: env.reduce(x, "max"), lambda x: env.reduce(x, "sum"),
                                      vox, env.stream))
        project["parity"] = B.sharded_parity(subs[:args.parity_submaps], poses[:args.parity_submaps],
                                             partial, owned, env.rank, env.world)
        partial.close()
        owned.close()
    extra = {"project_submaps": project}
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        extra["cpu_baseline"] = cpu_baseline(pool[0], C3_CFG, 0.05, args.cpu_seconds)
    if env.rank == 0:
        print(json.dumps(fusion_line(env, spec, results, extra)), flush=True)
    for L in subs:
        L.close()
    submap.close()
    glob.close()


def run_c5(env):
    """512 submaps re-projected after a pose-graph update, one GPU."""
    from coxgraph_b200 import Layer, reprojectSubmaps, synth
    args = env.args
    n_sub = 8 * args.c3_submaps
    t0 = time.perf_counter()
    subs = fuse_submaps(env, [(r, sm) for r in range(8) for sm in range(args.c3_submaps)], C3_CFG)
    build_s = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    base = [synth.robot_map_offset(0) for _ in subs]
    poses_old = np.stack(base)
    poses_new = np.stack([synth.perturb_pose(T, rng, sigma_t=0.05, sigma_yaw_deg=1.0) for T in base])
    glob = Layer(env.ctx, 0.05, max_blocks=262144)
    blocks_in = sum(L.num_blocks for L in subs)
    sampler = B.ClockSampler(env.local_rank)
    for _ in range(max(1, args.warmup // 2)):
        time_projection(env, subs, poses_new, glob, repeats=2)
    sampler.start()
    launches0 = env.ctx.kernel_launches
    times = []
    st = None
    for it in range(args.steps):
        glob.clear()
        a, b = env.event(), env.event()
        a.record(env.stream)
        from coxgraph_b200 import getProjectedMap
        st = getProjectedMap(subs, poses_new, glob, want_stats=(it == args.steps - 1))
        b.record(env.stream)
        env.torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    clocks = sampler.stop()
    launches = env.ctx.kernel_launches - launches0
    ms = float(np.mean(times))
    bytes_step = B.BLOCK_BYTES * (st.blocks_in + 2 * st.blocks_out)
    peak, peak_src = B.measured_peak_gbs()
    # stage times of the same call
    env.ctx.reset_profile()
    env.ctx.set_profiling(True)
    glob.clear()
    getProjectedMap(subs, poses_new, glob)
    env.ctx.set_profiling(False)
    prof = env.ctx.profile()
    top = max((k for k in prof if prof[k][0] > 0), key=lambda k: prof[k][0])
    # e2e: the poses come from the host every call and the mesh of the merged map goes back (what
    # saveAndPubCombinedMesh publishes, server_visualizer.cpp:123-126); the submaps are resident
    t_e2e = []
    d2h = 0
    for it in range(3):
        env.torch.cuda.synchronize()
        t0 = time.perf_counter()
        glob.clear()
        getProjectedMap(subs, poses_new, glob)
        mesh = glob.generateMesh()
        env.torch.cuda.synchronize()
        t_e2e.append((time.perf_counter() - t0) * 1e3)
        d2h = int(len(mesh[2]) * 28)
    # incremental re-projection when a tenth of the poses moved
    glob.clear()
    getProjectedMap(subs, poses_old, glob)
    moved = poses_old.copy()
    for k in rng.choice(n_sub, max(1, n_sub // 10), replace=False):
        moved[k] = poses_new[k]
    a, b = env.event(), env.event()
    a.record(env.stream)
    _, rst = reprojectSubmaps(subs, poses_old, moved, glob)
    b.record(env.stream)
    env.torch.cuda.synchronize()
    voxels = 4096.0 * blocks_in
    line = {
        "metric": "submap_merge_voxels_per_s", "value": voxels / (ms * 1e-3), "unit": "voxels/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"C5 global-map re-merge: {n_sub} submaps (8 robots x "
                               f"{args.c3_submaps}, 10 frames each, corridor scene, 5 cm voxels) "
                               "re-projected into an empty global TSDF after a pose-graph update "
                               "(sigma_t 5 cm, sigma_yaw 1 deg)",
                   "parallelism": "one GPU",
                   "l2": f"the submaps hold {blocks_in * B.BLOCK_BYTES / 1e6:.0f} MB, the global map "
                         f"{glob.num_blocks * B.BLOCK_BYTES / 1e6:.0f} MB (L2: 126 MB)"},
        "e2e": {"value": voxels / (min(t_e2e[1:]) * 1e-3), "unit": "voxels/s",
                "ms_per_step": min(t_e2e[1:]), "h2d_bytes_per_step": 28 * n_sub,
                "d2h_bytes_per_step": d2h,
                "what": "poses from the host, re-projection, marching cubes, mesh to the host"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": top, "library_kernel": prof[top][1] == 0,
                     "achieved": bytes_step / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_step / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_step,
                     "ms_per_launch": ms, "step_frac": bytes_step / (ms * 1e-3) / 1e9 / peak,
                     "note": "49152 (B_in + 2 B_out) bytes per call over the whole call (mark + "
                             "list + project kernels)"},
        "stages_ms_per_step": {k: v[0] for k, v in prof.items() if v[0] > 0},
        "per_step": {"blocks_in": int(st.blocks_in), "blocks_out": int(st.blocks_out),
                     "global_blocks": glob.num_blocks, "submaps": n_sub},
        "reproject_10pct_moved": {"ms": a.elapsed_time(b), "submaps_moved": int(rst.submaps_moved),
                                  "blocks_dirty": int(rst.blocks_dirty),
                                  "candidates": int(rst.candidates)},
        "setup_s": build_s, "clocks": clocks,
    }
    # the ESDF of the re-merged map (the client's updateEsdfBatch, map_server.h:141-145), on the
    # full-rebuild result
    glob.clear()
    getProjectedMap(subs, poses_new, glob)
    line["esdf"] = B.time_esdf(glob)
    if not args.no_cpu_baseline:
        # the reference merges single-threaded: time a bounded number of submaps with the oracle
        from oracle import oracle_py as orc
        og = orc.Layer(0.05)
        t0 = time.perf_counter()
        done = 0
        for L, T in zip(subs, poses_new):
            ol = orc.Layer(0.05)
            ol.upload(*L.download())
            t1 = time.perf_counter()
            og.merge_from(ol, T)
            done += 4096 * L.num_blocks
            if time.perf_counter() - t0 > args.cpu_seconds:
                break
            del t1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": done / dt, "unit": "voxels/s", "cores": 1, "kind": "port",
                                "sample": f"{done // 4096} source blocks merged single-threaded, as "
                                          "the reference does (oracle port; includes the upload)"}
    print(json.dumps(line), flush=True)
    for L in subs:
        L.close()
    glob.close()


# ----------------------------------------------------------------------------- C4
def run_c4(env):
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    args, torch = env.args, env.torch
    target = args.c4_blocks
    frames_per_step = 4
    layer = Layer(env.ctx, 0.02, max_blocks=int(target * 1.03) + 4096)
    integ = TsdfIntegrator(TsdfIntegratorConfig(**C4_CFG), layer)
    legs = {"device": [], "e2e": []}
    stats_hist = []
    step = 0
    t_setup = time.perf_counter()
    sampler = B.ClockSampler(env.local_rank)
    sampler.start()
    launches0 = env.ctx.kernel_launches
    last = None
    while layer.num_blocks < target and step < 20000:
        host_leg = (step // 2) % 2 == 1  # both robots (step % 2) appear in both legs
        fr = synth.corridor_frames(step * frames_per_step, frames_per_step, robot=step % 2,
                                   advance=1.0, cam=synth.CAM_1280x720, device=env.dev)
        e = make_entry(env, fr, pinned=host_leg)
        torch.cuda.synchronize()
        a, b = env.event(), env.event()
        a.record(env.stream)
        if not host_leg:    # inputs resident in HBM
            st = integ.integrateBatch(e["poses"], e["d_pts"], e["d_cols"], e["offs"])
        else:               # host buffers: the H2D copy is inside the timed region
            st = integ.integrateBatch(e["poses"], e["h_pts"], e["h_cols"], e["offs"])
        b.record(env.stream)
        torch.cuda.synchronize()
        legs["e2e" if host_leg else "device"].append((a.elapsed_time(b), e["n"], layer.num_blocks))
        stats_hist.append((int(st.rays), int(st.voxel_updates), int(st.general_updates),
                           int(st.blocks_touched), e["n"]))
        last = e
        step += 1
    clocks = sampler.stop()
    launches = env.ctx.kernel_launches - launches0
    setup_s = time.perf_counter() - t_setup
    K = max(args.steps, 10)

    def rate(rows):
        rows = rows[-K:]
        return sum(r[1] for r in rows) / (sum(r[0] for r in rows) * 1e-3), \
            float(np.mean([r[0] for r in rows]))

    v_dev, ms_dev = rate(legs["device"])
    v_e2e, ms_e2e = rate(legs["e2e"])
    v_first, _ = rate(legs["device"][:K])
    hs = layer.hash_stats()
    # a subset of the 98 GB layer comes back through the listed-blocks download
    probe_idx = layer_probe_indices(last, 0.02)
    t0 = time.perf_counter()
    vox, flags, found = layer.download_blocks(probe_idx)
    dl_ms = (time.perf_counter() - t0) * 1e3
    tail = stats_hist[-2 * K:]
    n_pts = float(np.mean([r[4] for r in tail]))
    touched = float(np.mean([r[3] for r in tail]))
    # contract bytes per step; B_touched here is per 4-frame job (a lower bound of the per-frame sum)
    bytes_step = 16 * n_pts + 2 * B.BLOCK_BYTES * touched
    env.ctx.reset_profile()
    env.ctx.set_profiling(True)
    integ.integrateBatch(last["poses"], last["d_pts"], last["d_cols"], last["offs"])
    env.ctx.set_profiling(False)
    prof = env.ctx.profile()
    per_step = dict(n_pts=n_pts, rays=float(np.mean([r[0] for r in tail])),
                    pairs=float(np.mean([r[1] for r in tail])),
                    general=float(np.mean([r[2] for r in tail])), blocks=touched)
    line = {
        "metric": "tsdf_points_integrated_per_s", "value": v_dev, "unit": "points/s", "n_gpus": 1,
        "steps": K, "warmup": max(0, len(legs["device"]) - K), "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"C4 fine resolution: 2 cm voxels, 6 cm truncation, max_ray 3 m, "
                               f"1280x720 frames, corridor trajectory until {target} blocks are "
                               f"allocated; per step one {frames_per_step}-frame job "
                               f"({frames_per_step * 921600} points) into the growing layer; value = "
                               f"the last {K} device-resident steps",
                   "parallelism": "one GPU",
                   "l2": "3.7 M fresh points per step; the layer is far larger than L2"},
        "e2e": {"value": v_e2e, "unit": "points/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 16 * n_pts, "d2h_bytes_per_step": 0,
                "what": "cg_integrate_batch with pinned host buffers on alternate steps (the layer "
                        "stays on the device: 98 GB); subset read-back below"},
        "gpu_launches": launches,
        "roofline": B.roofline_record(prof, 1, ms_dev, bytes_step, per_step, traffic_file=None),
        "stages_ms_per_step": {k: v[0] for k, v in prof.items() if v[0] > 0},
        "per_step": per_step,
        "layer": {"blocks": layer.num_blocks, "pool_bytes": layer.num_blocks * B.BLOCK_BYTES,
                  "hash_capacity": int(hs.hash_capacity), "hash_load": hs.load_factor,
                  "mean_probe_length": hs.mean_probe_length,
                  "max_probe_length": int(hs.max_probe_length), "frames": step * frames_per_step,
                  "value_first_steps": v_first, "fill_wall_s": setup_s,
                  "subset_download": {"blocks_asked": int(len(probe_idx)),
                                      "blocks_found": int(found.sum()), "ms": dl_ms,
                                      "observed_voxels": int((vox["weight"][found] > 0).sum())}},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        e = last if "h_pts" in last else make_entry(env, [(last["poses"][f],
                                                           last["d_pts"][int(last["offs"][f]):int(last["offs"][f + 1])],
                                                           last["d_cols"][int(last["offs"][f]):int(last["offs"][f + 1])])
                                                          for f in range(last["frames"])])
        line["cpu_baseline"] = cpu_baseline(e, C4_CFG, 0.02, args.cpu_seconds)
    print(json.dumps(line), flush=True)
    layer.close()


def layer_probe_indices(e, voxel):
    """Block indices around the last job's camera positions (some allocated, some not)."""
    bs = 16 * voxel
    idx = []
    for T in e["poses"]:
        c = np.floor(np.asarray(T[4:], np.float64) / bs).astype(np.int32)
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                idx.append((c[0] + dx, c[1] + dy, c[2]))
    return np.unique(np.asarray(idx, np.int32), axis=0)


def run(args, rank, world, local_rank):
    env = Env(args, rank, world, local_rank)
    try:
        if args.config == "C1":
            run_c1(env)
        elif args.config == "C3":
            run_c3(env)
        elif args.config == "C4":
            if rank == 0:
                run_c4(env)
        elif args.config == "C5":
            if rank == 0:
                run_c5(env)
    finally:
        env.close()
