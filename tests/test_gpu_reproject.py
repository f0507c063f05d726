"""GPU: Layer::removeBlock and the incremental re-projection.

The reference rebuilds the whole global map after every pose-graph update
(coxgraph/include/coxgraph/server/coxgraph_server.h:275-283 ->
coxgraph/src/server/visualizer/server_visualizer.cpp:123-126).  cg_reproject_submaps rebuilds only
the destination blocks a moved submap reaches; the contract is that the layer ends up
bit-identical to that full rebuild (whose parity with the oracle is tests/test_gpu_merge.py's
subject), so the checks here are exact comparisons against cg_project_submaps into a fresh layer.
"""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _base_submaps(ctx, count, frames=2, stride=4):
    import torch
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    dev = torch.device("cuda", 0)
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, default_truncation_distance=0.16)
    out = []
    for k in range(count):
        L = Layer(ctx, 0.05, max_blocks=2048)
        integ = TsdfIntegrator(cfg, L)
        for (T, pts, cols) in synth.submap_frames(k % 2, k, frames, device=dev, stride=stride):
            integ.integratePointCloud(T, pts, cols)
        out.append(L)
    return out


def test_remove_blocks_keeps_the_rest_and_the_hash_usable(gpu_ctx):
    from coxgraph_b200 import Layer
    (L,) = _base_submaps(gpu_ctx, 1)
    idx, vox, flags = L.download()
    n = len(idx)
    assert n > 20
    rng = np.random.default_rng(3)
    drop = np.sort(rng.choice(n, n // 3, replace=False))
    absent = np.array([[900, 900, 900], [-77, 3, 5]], np.int32)
    assert L.removeBlocks(np.concatenate([idx[drop], absent])) == len(drop)
    keep = np.setdiff1d(np.arange(n), drop)
    i2, v2, f2 = L.download()
    assert L.num_blocks == len(keep)
    assert np.array_equal(i2, idx[keep]) and np.array_equal(f2, flags[keep])
    for name in ("distance", "weight", "rgba"):
        assert np.array_equal(v2[name], vox[keep][name])
    # freed slots are default-constructed again and the rebuilt hash finds / inserts correctly:
    # putting the removed blocks back restores the original layer
    L.upload(idx[drop], vox[drop], flags[drop])
    i3, v3, f3 = L.download()
    assert np.array_equal(i3, idx) and np.array_equal(f3, flags)
    for name in ("distance", "weight", "rgba"):
        assert np.array_equal(v3[name], vox[name])
    # removing everything one by one ends at an empty layer
    assert L.removeBlocks(idx) == n and L.num_blocks == 0
    assert len(L.download()[0]) == 0
    L.close()


def _fresh_projection(ctx, subs, poses):
    from coxgraph_b200 import Layer, getProjectedMap
    g = Layer(ctx, 0.05, max_blocks=16384)
    getProjectedMap(subs, poses, g)
    return g


@pytest.mark.parametrize("block_by_block", [True, False])
@pytest.mark.parametrize("nsub", [12, 80])
def test_reprojection_equals_full_rebuild(gpu_ctx, nsub, block_by_block, monkeypatch):
    """block_by_block: the dirty blocks are rebuilt one by one whatever their share of the map;
    otherwise the library takes the full rebuild once more than 30 % of the map is dirty (these
    room-sized maps always are) — the same layer either way."""
    from coxgraph_b200 import Layer, getProjectedMap, reprojectSubmaps, synth
    if block_by_block:
        monkeypatch.setenv("CG_REPROJECT_FULL_FRACTION", "2")
    base = _base_submaps(gpu_ctx, 6)
    subs = [base[k % 6] for k in range(nsub)]
    rng = np.random.default_rng(nsub)
    old = np.stack([synth.perturb_pose(synth.robot_map_offset(k % 2), rng, sigma_t=0.4,
                                       sigma_yaw_deg=25.0) for k in range(nsub)])
    g = Layer(gpu_ctx, 0.05, max_blocks=16384)
    getProjectedMap(subs, old, g)
    # nothing moved: nothing happens
    before = g.download()
    changed, st = reprojectSubmaps(subs, old, old, g)
    assert not changed.any() and st.submaps_moved == 0
    util.compare_layers(g.download(), before, "unchanged poses", exact=True)
    cur = old.copy()
    for step in range(3):
        new = cur.copy()
        moved = rng.choice(nsub, max(2, nsub // 6), replace=False)
        for k in moved:
            new[k] = synth.perturb_pose(cur[k], rng, sigma_t=0.05, sigma_yaw_deg=1.0)
        far = int(moved[0])
        new[far, 4:7] += np.array([40.0 + 7 * step, -25.0, 3.0], np.float32)  # leaves its old blocks
        still = int(np.setdiff1d(np.arange(nsub), moved)[0])
        new[still, 4] += 1e-4                                   # below the threshold: not moved
        changed, st = reprojectSubmaps(subs, cur, new, g, eps_translation=1e-3, eps_rotation=1e-4)
        assert np.array_equal(np.flatnonzero(changed), np.sort(moved))
        eff = np.where(changed[:, None], new, cur)
        want = _fresh_projection(gpu_ctx, subs, eff)
        util.compare_layers(g.download(), want.download(), f"re-projection step {step}", exact=True,
                            check_flags=True)
        assert st.submaps_moved == len(moved) and st.blocks_dirty > 0
        assert st.candidates >= st.blocks_folded > 0
        want.close()
        cur = eff
    # a submap that moves away from everything takes its blocks with it
    only = np.array([1, 0, 0, 0, 0, 0, 0], np.float32)[None]
    h = Layer(gpu_ctx, 0.05, max_blocks=4096)
    getProjectedMap(subs[:1], only, h)
    n0 = h.num_blocks
    away = only.copy()
    away[0, 4:7] = (80.0, 80.0, 0.0)
    changed, st = reprojectSubmaps(subs[:1], only, away, h)
    assert changed.all() and h.num_blocks > 0
    assert (st.blocks_removed == n0) if block_by_block else (st.full_rebuild == 1)
    want = _fresh_projection(gpu_ctx, subs[:1], away)
    util.compare_layers(h.download(), want.download(), "moved away", exact=True, check_flags=True)
    for L in base + [g, h, want]:
        L.close()
