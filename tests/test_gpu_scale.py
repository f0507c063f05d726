"""Size-independent properties at BASELINE.json's full input sizes (the oracle is too slow to
compare voxel by voxel here): full 640x480 / 1280x720 frames, many submaps, re-merges.
  * batch == frame-by-frame (same layer, within the parity tolerance)
  * every stored distance lies in [-trunc, trunc], every weight in [0, max_weight]
  * block count is independent of how the job is cut into groups
  * merging a layer into an empty one with the identity pose reproduces it (weights/distances
    exactly, distances to 2 ulp) and merging it twice doubles the weights
  * translating a submap by a whole number of blocks permutes block indices exactly
  * projecting N submaps in one call == N single merges in order (bit-exact)
Reference shapes: configs[1] (C2), configs[3] (C4, 2 cm / 1280x720), configs[4] (C5, re-merge)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _device_frames(robot, submap, n, cam, dev):
    import torch
    from coxgraph_b200 import synth
    fr = synth.submap_frames(robot, submap, n, cam=cam, device=dev)
    poses = np.stack([T for (T, _, _) in fr]).astype(np.float32)
    pts = torch.cat([p for (_, p, _) in fr]).contiguous()
    cols = torch.cat([c for (_, _, c) in fr]).contiguous()
    offs = np.cumsum([0] + [len(p) for (_, p, _) in fr]).astype(np.uint64)
    return poses, pts, cols, offs


def _check_ranges(vox, trunc, max_weight):
    d, w = vox["distance"], vox["weight"]
    assert np.isfinite(d).all() and np.isfinite(w).all()
    assert (np.abs(d) <= trunc * (1 + 1e-6)).all(), "distance outside the truncation band"
    assert (w >= 0).all() and (w <= max_weight).all()
    assert (vox["rgba"][..., 3][w > 0] > 0).all()


@pytest.mark.parametrize("shape", ["C2_5cm_480p", "C4_2cm_720p"])
def test_full_size_batch_equals_sequential(gpu_ctx, shape):
    import torch
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    dev = torch.device("cuda", 0)
    if shape == "C2_5cm_480p":
        voxel, cfg, cam, n = 0.05, dict(default_truncation_distance=0.16, max_ray_length_m=5.0), \
            synth.CAM_640x480, 6
    else:
        voxel, cfg, cam, n = 0.02, dict(default_truncation_distance=0.06, max_ray_length_m=3.0), \
            synth.CAM_1280x720, 4
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, **cfg)
    poses, pts, cols, offs = _device_frames(0, 1, n, cam, dev)
    a, b = Layer(gpu_ctx, voxel, max_blocks=16384), Layer(gpu_ctx, voxel, max_blocks=16384)
    st = TsdfIntegrator(cfg, a).integrateBatch(poses, pts, cols, offs)
    assert st.points_in == int(offs[-1]) and st.rays > 0 and st.voxel_updates > st.rays
    ib = TsdfIntegrator(cfg, b)
    for f in range(n):
        lo, hi = int(offs[f]), int(offs[f + 1])
        ib.integratePointCloud(poses[f], pts[lo:hi], cols[lo:hi])
    la, lb = a.download(), b.download()
    util.compare_layers(la, lb, f"{shape}: batch vs per-frame")
    _check_ranges(la[1], cfg.default_truncation_distance, cfg.max_weight)
    assert a.num_blocks == st.blocks_allocated
    a.close()
    b.close()


def test_group_cut_does_not_change_the_block_set(gpu_ctx, monkeypatch):
    import torch
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    dev = torch.device("cuda", 0)
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, default_truncation_distance=0.16)
    poses, pts, cols, offs = _device_frames(1, 0, 6, synth.CAM_640x480, dev)
    a, b = Layer(gpu_ctx, 0.05, max_blocks=8192), Layer(gpu_ctx, 0.05, max_blocks=8192)
    TsdfIntegrator(cfg, a).integrateBatch(poses, pts, cols, offs)
    monkeypatch.setenv("CG_MAX_GROUP_POINTS", str(2 * 307200))
    TsdfIntegrator(cfg, b).integrateBatch(poses, pts, cols, offs)
    util.compare_layers(a.download(), b.download(), "one group vs three groups")
    a.close()
    b.close()


def test_merge_algebra_at_full_size(gpu_ctx):
    import torch
    from coxgraph_b200 import (Layer, TsdfIntegrator, TsdfIntegratorConfig, getProjectedMap,
                               mergeLayerAintoLayerB, synth)
    dev = torch.device("cuda", 0)
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, default_truncation_distance=0.16)
    subs, T_M_S = [], []
    for k in range(4):
        poses, pts, cols, offs = _device_frames(k % 2, k, 4, synth.CAM_640x480, dev)
        L = Layer(gpu_ctx, 0.05, max_blocks=4096)
        TsdfIntegrator(cfg, L).integrateBatch(poses, pts, cols, offs)
        subs.append(L)
        T_M_S.append(synth.robot_map_offset(k % 2))
    ident = np.array([1, 0, 0, 0, 0, 0, 0], np.float32)
    # identity merge into an empty layer: observed voxels come back exactly
    g = Layer(gpu_ctx, 0.05, max_blocks=16384)
    mergeLayerAintoLayerB(subs[0], ident, g)
    si, sv, _ = subs[0].download()
    gi, gv, _ = g.download()
    keep = np.array([(sv[b]["weight"] > 1e-6).any() for b in range(len(si))])
    assert np.array_equal(gi, si[keep]), "identity merge must keep exactly the blocks with data"
    obs = sv[keep]["weight"] > 1e-6
    assert np.array_equal(gv["weight"][obs], sv[keep]["weight"][obs])
    # (Da * Wa + 0) / Wa: one rounding of the product, one of the quotient
    assert np.allclose(gv["distance"][obs], sv[keep]["distance"][obs], rtol=3e-7, atol=1e-9)
    # merging it again doubles the weights and leaves the distances
    mergeLayerAintoLayerB(subs[0], ident, g)
    _, gv2, _ = g.download()
    assert np.array_equal(gv2["weight"][obs], 2 * sv[keep]["weight"][obs])
    assert np.allclose(gv2["distance"][obs], sv[keep]["distance"][obs], rtol=1e-6, atol=1e-7)
    # a translation by whole blocks permutes block indices exactly
    shift = np.array([1, 0, 0, 0, 0.8 * 3, -0.8 * 2, 0.8], np.float32)
    h = Layer(gpu_ctx, 0.05, max_blocks=16384)
    mergeLayerAintoLayerB(subs[0], shift, h)
    hi_, hv, _ = h.download()
    moved = gi + np.array([3, -2, 1], np.int32)
    order = np.lexsort((moved[:, 0], moved[:, 1], moved[:, 2]))
    assert np.array_equal(hi_, moved[order])
    # getProjectedMap(all) == the same merges one by one, bit for bit (C5: a full re-merge)
    p, q = Layer(gpu_ctx, 0.05, max_blocks=16384), Layer(gpu_ctx, 0.05, max_blocks=16384)
    st = getProjectedMap(subs, np.stack(T_M_S), p, want_stats=True)
    for L, T in zip(subs, T_M_S):
        mergeLayerAintoLayerB(L, T, q)
    util.compare_layers(p.download(), q.download(), "project vs merges", exact=True)
    assert st.blocks_in == sum(L.num_blocks for L in subs) and st.blocks_out >= p.num_blocks
    # re-merge after a pose update starts from an empty global layer (server_visualizer.cpp:123)
    rng = np.random.default_rng(5)
    p.removeAllBlocks()
    getProjectedMap(subs, np.stack([synth.perturb_pose(T, rng) for T in T_M_S]), p)
    assert 0.5 * q.num_blocks < p.num_blocks < 2 * q.num_blocks
    for L in subs + [g, h, p, q]:
        L.close()


def test_projection_in_several_batches_is_sequential_and_repeatable(gpu_ctx):
    """More submaps than one batch holds (64): destination blocks reached by submaps of different
    batches and by many submaps of one batch (the in-kernel turn order of k_project_batch).  The
    projection must equal the single merges in submap order bit for bit, every time it runs."""
    import torch
    from coxgraph_b200 import (Layer, TsdfIntegrator, TsdfIntegratorConfig, getProjectedMap,
                               mergeLayerAintoLayerB, synth)
    dev = torch.device("cuda", 0)
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, default_truncation_distance=0.16)
    rng = np.random.default_rng(11)
    base, T_M_S = [], []
    for k in range(6):  # six distinct submaps, reused under 150 different poses
        fr = synth.submap_frames(k % 2, k, 2, device=dev, stride=4)
        L = Layer(gpu_ctx, 0.05, max_blocks=2048)
        integ = TsdfIntegrator(cfg, L)
        for (T, pts, cols) in fr:
            integ.integratePointCloud(T, pts, cols)
        base.append(L)
    subs = [base[k % 6] for k in range(150)]
    for k in range(150):
        T_M_S.append(synth.perturb_pose(synth.robot_map_offset(k % 2), rng, sigma_t=0.3,
                                        sigma_yaw_deg=20.0))
    poses = np.stack(T_M_S)
    q = Layer(gpu_ctx, 0.05, max_blocks=16384)
    for L, T in zip(subs, T_M_S):
        mergeLayerAintoLayerB(L, T, q)
    want = q.download()
    p = Layer(gpu_ctx, 0.05, max_blocks=16384)
    for rep in range(3):
        p.removeAllBlocks()
        st = getProjectedMap(subs, poses, p, want_stats=True)
        util.compare_layers(p.download(), want, f"150-submap projection, run {rep}", exact=True)
        assert st.blocks_in == sum(L.num_blocks for L in subs)
        assert st.blocks_out >= p.num_blocks and st.blocks_candidate >= st.blocks_out
    for L in base + [p, q]:
        L.close()


def test_block_pool_beyond_4_gib(gpu_ctx):
    """C4 needs ~2 M blocks (98 GB): voxel offsets must be 64-bit.  Fill a layer with > 87 k far-away
    blocks (4.9 GB of pool) so that every block the frames touch lies beyond the 4 GiB mark, then
    fuse, merge and read back against the oracle."""
    from coxgraph_b200 import Layer, TsdfIntegrator, mergeLayerAintoLayerB, VOXEL_DTYPE
    from oracle import oracle_py as orc
    ocfg, gcfg = util.make_cfgs()
    gl = Layer(gpu_ctx, 0.05, max_blocks=100_000)
    chunk = 2048
    filler = np.zeros((chunk, 4096), VOXEL_DTYPE)
    filler["distance"], filler["weight"] = 0.03, 2.0
    filler["rgba"] = (9, 8, 7, 255)
    n_fill = 0
    for k in range(44):                     # 44 x 2048 = 90,112 blocks = 4.43 GB
        idx = np.stack([np.arange(chunk, dtype=np.int32) - 1024, np.full(chunk, 5000 + k, np.int32),
                        np.full(chunk, -3000, np.int32)], axis=1)
        gl.upload(idx, filler)
        n_fill += chunk
    assert gl.num_blocks == n_fill and n_fill * 49152 > 4 * 2**30
    frames = util.small_frames(2, stride=8)
    ol = orc.Layer(0.05)
    integ = TsdfIntegrator(gcfg, gl)
    for (T, p, c) in frames:
        ol.integrate(ocfg, T, p, c)
        integ.integratePointCloud(T, p, c)
    oi, ov, of = ol.download()
    assert gl.num_blocks == n_fill + len(oi)
    gi, gv, gf = gl.download()
    mine = gi[:, 1] < 4000                  # the filler blocks sit at y >= 5000
    util.compare_layers((gi[mine], gv[mine], gf[mine]), (oi, ov, of), "blocks beyond 4 GiB")
    far = gv[~mine]                           # the filler blocks are untouched
    assert (far["distance"] == np.float32(0.03)).all() and (far["weight"] == 2.0).all()
    assert (far["rgba"] == np.array([9, 8, 7, 255], np.uint8)).all()
    # merge out of the big layer (source blocks beyond 4 GiB) into a fresh one
    og, gg = orc.Layer(0.05), Layer(gpu_ctx, 0.05, max_blocks=4096)
    small = Layer(gpu_ctx, 0.05, max_blocks=2048)
    small.upload(gi[mine], gv[mine])
    T = np.array([np.cos(0.2), 0, 0, np.sin(0.2), 0.3, 0.1, 0.0], np.float32)
    og.merge_from(ol, T)
    mergeLayerAintoLayerB(small, T, gg)
    util.compare_layers(gg.download(), og.download(), "merge after the big layer")
    for L in (gl, gg, small):
        L.close()


def test_results_are_run_to_run_deterministic(gpu_ctx):
    """north_star: "updates stay deterministic".  The same job twice (full frames, 1/z^2 weights so
    that the weight sums are not integers) gives bit-identical layers, and so does the batched
    projection: the free-space sums are integer (fixed point), the ordered lists are sorted on
    their full key, the folds run in submap order."""
    import torch
    from coxgraph_b200 import (Layer, TsdfIntegrator, TsdfIntegratorConfig, getProjectedMap, synth)
    dev = torch.device("cuda", 0)
    cfg = TsdfIntegratorConfig(use_const_weight=0, method=1, default_truncation_distance=0.16)
    poses, pts, cols, offs = _device_frames(0, 3, 5, synth.CAM_640x480, dev)
    layers = []
    for rep in range(2):
        L = Layer(gpu_ctx, 0.05, max_blocks=8192)
        TsdfIntegrator(cfg, L).integrateBatch(poses, pts, cols, offs)
        layers.append(L)
    util.compare_layers(layers[0].download(), layers[1].download(), "same job twice", exact=True)
    T = np.stack([synth.robot_map_offset(1), synth.robot_map_offset(0)])
    globs = []
    for rep in range(2):
        G = Layer(gpu_ctx, 0.05, max_blocks=16384)
        getProjectedMap(layers, T, G)
        globs.append(G)
    util.compare_layers(globs[0].download(), globs[1].download(), "same projection twice", exact=True)
    for L in layers + globs:
        L.close()
