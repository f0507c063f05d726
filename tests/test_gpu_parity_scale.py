"""CUDA path against the oracle AT THE SIZES THE BENCH QUOTES (BASELINE.json configs): one full C2
step (25 stride-1 640x480 frames fused, then merged into the global layer), the 40-submap
projection of dense submaps (cblox getProjectedMap, server_visualizer.cpp:123-126), and a C4-shaped
frame (2 cm voxels, 1280x720 camera).  Block sets bit-exact; distance / weight within 1e-4
relative or 1e-5 absolute; colours within 1 LSB — and the margins against those tolerances are
recorded, not only pass / fail (gpurun_out/parity_margins.json)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

C2 = dict(default_truncation_distance=0.16, max_ray_length_m=5.0, min_ray_length_m=0.1,
          use_const_weight=1, method=1)


def _batch(frames):
    poses = np.stack([T for (T, _, _) in frames]).astype(np.float32)
    pts = np.concatenate([p for (_, p, _) in frames])
    cols = np.concatenate([c for (_, _, c) in frames])
    offs = np.cumsum([0] + [len(p) for (_, p, _) in frames]).astype(np.uint64)
    return poses, pts, cols, offs


@pytest.mark.parametrize("robot", [0, 1])
def test_full_c2_step_against_oracle(gpu_ctx, robot):
    """25 full-density frames as one job (7.68 M points), then mergeLayerAintoLayerB into the
    global layer — exactly one bench step."""
    from coxgraph_b200 import Layer, TsdfIntegrator, mergeLayerAintoLayerB, synth
    from oracle import oracle_py as orc
    ocfg, gcfg = util.make_cfgs(**C2)
    frames = util.small_frames(25, stride=1, robot=robot, submap=3)
    sub, glob = Layer(gpu_ctx, 0.05, max_blocks=4096), Layer(gpu_ctx, 0.05, max_blocks=8192)
    o_sub, o_glob = orc.Layer(0.05), orc.Layer(0.05)
    poses, pts, cols, offs = _batch(frames)
    st = TsdfIntegrator(gcfg, sub).integrateBatch(poses, pts, cols, offs)
    assert st.points_in == 25 * 307200
    for (T, p, c) in frames:
        o_sub.integrate(ocfg, T, p, c)
    T_M_S = synth.robot_map_offset(1)
    mergeLayerAintoLayerB(sub, T_M_S, glob)
    o_glob.merge_from(o_sub, T_M_S)
    for name, g, o in (("submap", sub, o_sub), ("global", glob, o_glob)):
        got, ref = g.download(), o.download()
        util.record_margins(f"c2_step_robot{robot}_{name}", util.margins(got, ref))
        util.compare_layers(got, ref, f"full C2 step, robot {robot}, {name}")
    sub.close()
    glob.close()


def test_forty_dense_submaps_projected(gpu_ctx):
    """getProjectedMap over 40 dense submaps (2 robots x 20, 6 full frames each, every submap with
    its own pose) against the oracle's sequential mergeLayerAintoLayerB."""
    from coxgraph_b200 import Layer, TsdfIntegrator, getProjectedMap, synth
    from oracle import oracle_py as orc
    _, gcfg = util.make_cfgs(**C2)
    subs, poses = [], []
    o_glob = orc.Layer(0.05)
    rng = np.random.default_rng(11)
    for k in range(40):
        robot, sm = k % 2, k // 2
        frames = util.small_frames(6, stride=1, robot=robot, submap=sm)
        L = Layer(gpu_ctx, 0.05, max_blocks=1024)
        p, pts, cols, offs = _batch(frames)
        TsdfIntegrator(gcfg, L).integrateBatch(p, pts, cols, offs)
        subs.append(L)
        poses.append(synth.perturb_pose(synth.robot_map_offset(robot), rng, sigma_t=0.4,
                                        sigma_yaw_deg=10.0))
    poses = np.stack(poses)
    glob = Layer(gpu_ctx, 0.05, max_blocks=32768)
    getProjectedMap(subs, poses, glob)
    for L, T in zip(subs, poses):   # the device-fused submaps are the oracle's inputs
        ol = orc.Layer(0.05)
        ol.upload(*L.download())
        o_glob.merge_from(ol, T)
    got, ref = glob.download(), o_glob.download()
    util.record_margins("project_40_dense_submaps", util.margins(got, ref))
    util.compare_layers(got, ref, "40 dense submaps projected")
    for L in subs:
        L.close()
    glob.close()


def test_c4_shaped_frame_against_oracle(gpu_ctx):
    """2 cm voxels, 6 cm truncation, 3 m rays, two full 1280x720 frames (921,600 points each),
    corridor scene."""
    from coxgraph_b200 import Layer, TsdfIntegrator, synth
    from oracle import oracle_py as orc
    ocfg, gcfg = util.make_cfgs(default_truncation_distance=0.06, max_ray_length_m=3.0)
    frames = [(T, p.numpy(), c.numpy()) for (T, p, c) in
              synth.corridor_frames(40, 2, robot=1, advance=1.0, cam=synth.CAM_1280x720, stride=1)]
    gl, ol = Layer(gpu_ctx, 0.02, max_blocks=8192), orc.Layer(0.02)
    integ = TsdfIntegrator(gcfg, gl)
    for (T, p, c) in frames:
        integ.integratePointCloud(T, p, c)
        ol.integrate(ocfg, T, p, c)
    got, ref = gl.download(), ol.download()
    util.record_margins("c4_shaped_two_frames", util.margins(got, ref))
    util.compare_layers(got, ref, "C4-shaped frames")
    gl.close()


def test_c1_hundred_frames_into_one_submap(gpu_ctx):
    """configs[0] at full size: 100 stride-1 640x480 frames (30.72 M points) fused into ONE 5 cm
    submap as one job, against the sequential oracle frame by frame."""
    from coxgraph_b200 import Layer, TsdfIntegrator
    from oracle import oracle_py as orc
    ocfg, gcfg = util.make_cfgs(**C2)
    sub, o_sub = Layer(gpu_ctx, 0.05, max_blocks=8192), orc.Layer(0.05)
    frames = []
    import torch
    from coxgraph_b200 import synth
    for sm in range(4):  # four consecutive stretches of robot 0's trajectory: 100 distinct frames
        frames += [(T, p.cpu().numpy(), c.cpu().numpy()) for (T, p, c) in
                   synth.submap_frames(0, sm, 25, device=torch.device("cuda", 0))]
    poses, pts, cols, offs = _batch(frames)
    st = TsdfIntegrator(gcfg, sub).integrateBatch(poses, pts, cols, offs)
    assert st.points_in == 100 * 307200
    for (T, p, c) in frames:
        o_sub.integrate(ocfg, T, p, c)
    got, ref = sub.download(), o_sub.download()
    util.record_margins("c1_100_frames_one_submap", util.margins(got, ref))
    util.compare_layers(got, ref, "C1: 100 frames into one submap")
    sub.close()


def test_c3_shaped_corridor_submaps_projected(gpu_ctx):
    """configs[2] / configs[4] shape: submaps that follow one another ALONG a trajectory (corridor
    scene, 10 full frames each, 2 m per submap) fused on the device and projected into the global
    map under perturbed poses (the re-merge after a pose-graph update), against the oracle's
    fusion and its sequential mergeLayerAintoLayerB."""
    import torch
    from coxgraph_b200 import Layer, TsdfIntegrator, getProjectedMap, synth
    from oracle import oracle_py as orc
    ocfg, gcfg = util.make_cfgs(**C2)
    rng = np.random.default_rng(5)
    subs, o_subs, poses = [], [], []
    for sm in range(12):
        frames = [(T, p.cpu().numpy(), c.cpu().numpy()) for (T, p, c) in
                  synth.corridor_frames(sm * 10, 10, robot=0, advance=0.2, start_x=0.0,
                                        device=torch.device("cuda", 0))]
        L, ol = Layer(gpu_ctx, 0.05, max_blocks=768), orc.Layer(0.05)
        p, pts, cols, offs = _batch(frames)
        TsdfIntegrator(gcfg, L).integrateBatch(p, pts, cols, offs)
        for (T, q, c) in frames:
            ol.integrate(ocfg, T, q, c)
        if sm == 0:  # the fusion itself on this scene
            got, ref = L.download(), ol.download()
            util.record_margins("c3_corridor_submap_fused", util.margins(got, ref))
            util.compare_layers(got, ref, "corridor submap fused")
        subs.append(L)
        o_subs.append(ol)
        poses.append(synth.perturb_pose(synth.robot_map_offset(0), rng))  # sigma 5 cm / 1 deg (C5)
    poses = np.stack(poses)
    glob, o_glob = Layer(gpu_ctx, 0.05, max_blocks=16384), orc.Layer(0.05)
    getProjectedMap(subs, poses, glob)
    for L, T in zip(subs, poses):
        ol = orc.Layer(0.05)
        ol.upload(*L.download())
        o_glob.merge_from(ol, T)
    got, ref = glob.download(), o_glob.download()
    assert len(got[0]) > 3 * 135  # a map along a trajectory (one submap: 135 blocks), not one room
    util.record_margins("c3_12_corridor_submaps_projected", util.margins(got, ref))
    util.compare_layers(got, ref, "12 corridor submaps projected")
    for L in subs:
        L.close()
    glob.close()
