"""GPU parity: cg_layer_mesh (marching cubes on the device-resident layer) against
the oracle's restatement of voxblox MeshIntegrator / MarchingCubes.  Everything on this path is
IEEE single precision in a fixed order, so the comparison is bit-exact: block order, per-block
vertex ranges, vertices, normals and colours."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import util

pytestmark = pytest.mark.gpu


def _same_mesh(layer, min_weight=1e-4, use_color=True, only_updated=False, what="mesh"):
    """Mesh `layer` on the GPU and its downloaded copy through the oracle; compare exactly."""
    idx, vox, flags = layer.download()
    o = orc.Layer(layer.voxel_size)
    o.upload(idx, vox, flags)
    gi, gb, gv, gn, gc = layer.generateMesh(min_weight, use_color, only_updated)
    ob, ov, on, oc = o.mesh(min_weight, use_color, only_updated)
    assert np.array_equal(gi, idx), f"{what}: block order"
    assert np.array_equal(gb, ob), f"{what}: per-block vertex ranges"
    assert np.array_equal(gv.view(np.uint32), ov.view(np.uint32)), f"{what}: vertices"
    assert np.array_equal(gn.view(np.uint32), on.view(np.uint32)), f"{what}: normals"
    assert np.array_equal(gc, oc), f"{what}: colours"
    return gi, gb, gv, gn, gc


def _fused_submap(ctx, robot=0, submap=0, frames=2, stride=4):
    import torch
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, default_truncation_distance=0.16)
    L = Layer(ctx, 0.05, max_blocks=4096)
    integ = TsdfIntegrator(cfg, L)
    for (T, pts, cols) in synth.submap_frames(robot, submap, frames, device=torch.device("cuda", 0),
                                              stride=stride):
        integ.integratePointCloud(T, pts, cols)
    return L, integ


def test_mesh_of_a_fused_submap_matches_oracle(gpu_ctx):
    L, integ = _fused_submap(gpu_ctx, frames=3, stride=2)
    gi, gb, gv, gn, gc = _same_mesh(L, what="fused submap")
    assert len(gv) > 10000 and len(gv) % 3 == 0
    # the surface lies inside the truncation band of observed voxels: every vertex has a colour
    # from an observed voxel (alpha 255) and unit normals where the triangle is not degenerate
    tri = gv.reshape(-1, 3, 3)
    area2 = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    nn = np.linalg.norm(gn.reshape(-1, 3, 3)[:, 0], axis=1)
    assert np.abs(nn[area2 > 1e-8] - 1).max() < 1e-5
    _same_mesh(L, min_weight=0.5, what="higher min_weight")
    _same_mesh(L, use_color=False, what="no colour")
    # only_updated: everything is flagged after integration; nothing after the flags are reset;
    # then only the blocks touched by one more frame
    _same_mesh(L, only_updated=True, what="all updated")
    from coxgraph_b200 import capi, synth
    capi.check(capi.load().cg_layer_reset_updated(L._h))
    assert len(L.generateMesh(only_updated=True)[2]) == 0
    import torch
    (T, pts, cols), = synth.submap_frames(0, 3, 1, device=torch.device("cuda", 0), stride=4)
    integ.integratePointCloud(T, pts, cols)
    part = _same_mesh(L, only_updated=True, what="blocks of the last frame")
    assert 0 < len(part[2]) < len(_same_mesh(L, what="whole layer again")[2])
    L.close()


def test_mesh_of_a_projected_map_matches_oracle(gpu_ctx):
    from coxgraph_b200 import Layer, getProjectedMap, synth
    subs = [_fused_submap(gpu_ctx, robot=k % 2, submap=k)[0] for k in range(4)]
    rng = np.random.default_rng(2)
    poses = np.stack([synth.perturb_pose(synth.robot_map_offset(k % 2), rng, sigma_t=0.2,
                                         sigma_yaw_deg=15.0) for k in range(4)])
    g = Layer(gpu_ctx, 0.05, max_blocks=16384)
    getProjectedMap(subs, poses, g)
    out = _same_mesh(g, what="projected map")
    assert len(out[2]) > 10000
    for L in subs + [g]:
        L.close()


def test_mesh_of_fine_voxels_matches_oracle(gpu_ctx):
    """configs[3] shape: 2 cm voxels, 1280x720 (sub-sampled), 6 cm truncation."""
    import torch
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, synth
    cfg = TsdfIntegratorConfig(use_const_weight=1, method=1, default_truncation_distance=0.06,
                               max_ray_length_m=3.0)
    L = Layer(gpu_ctx, 0.02, max_blocks=16384)
    integ = TsdfIntegrator(cfg, L)
    for (T, pts, cols) in synth.submap_frames(2, 0, 2, cam=synth.CAM_1280x720,
                                              device=torch.device("cuda", 0), stride=4):
        integ.integratePointCloud(T, pts, cols)
    out = _same_mesh(L, what="2 cm layer")
    assert len(out[2]) > 10000
    L.close()


def test_mesh_of_an_analytic_sphere_is_closed(gpu_ctx):
    from coxgraph_b200 import Layer
    blocks = np.array([(x, y, z) for z in (-1, 0) for y in (-1, 0) for x in (-1, 0)], np.int32)
    ctr = np.array([0.013, -0.021, 0.007])
    vox = np.zeros((8, 4096), orc.VOXEL_DTYPE)
    lin = np.arange(4096)
    loc = np.stack([lin & 15, (lin >> 4) & 15, lin >> 8], -1)
    for b, bi in enumerate(blocks):
        c = (bi[None, :] * 16 + loc + 0.5) * 0.05
        vox[b]["distance"] = np.linalg.norm(c - ctr, axis=-1) - 0.5
        vox[b]["weight"] = 1.0
        vox[b]["rgba"] = (40, 50, 60, 255)
    L = Layer(gpu_ctx, 0.05, max_blocks=64)
    L.upload(blocks, vox)
    gi, gb, gv, gn, gc = _same_mesh(L, what="sphere")
    assert np.abs(np.linalg.norm(gv - ctr, axis=1) - 0.5).max() < 2e-3
    key = np.round(gv / 1e-5).astype(np.int64)
    _, ids = np.unique(key, axis=0, return_inverse=True)
    t = ids.reshape(-1, 3)
    t = t[(t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])]
    e = set(map(tuple, np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])))
    assert all((b, a) in e for (a, b) in e), "the sphere mesh is not watertight"
    assert (gc == np.array([40, 50, 60, 255], np.uint8)).all()
    # an empty layer meshes to nothing
    L.removeAllBlocks()
    assert len(L.generateMesh()[2]) == 0
    L.close()


def test_connected_mesh_matches_the_sequential_hash_map_walk(gpu_ctx):
    """cg_mesh_connect (MeshLayer::getConnectedMesh / createConnectedMesh on the device: sorts and
    scans) against the oracle's sequential hash-map walk: same unique vertices in order of first
    occurrence, same renumbered indices, normals and colours of the first occurrences."""
    L, integ = _fused_submap(gpu_ctx, frames=3, stride=2)
    gi, gb, gv, gn, gc = L.generateMesh()
    cv, cn, cc, cidx = L.getConnectedMesh()
    oidx, first = orc.connect_mesh(gv)
    assert len(cv) == len(first) and 0 < len(first) < len(gv) / 2
    assert np.array_equal(cidx, oidx)
    assert np.array_equal(cv.view(np.uint32), gv[first].view(np.uint32))
    assert np.array_equal(cn.view(np.uint32), gn[first].view(np.uint32))
    assert np.array_equal(cc, gc[first])
    assert L.getConnectedMesh(fetch=False) == (len(first), len(gv))
    # an empty mesh
    from coxgraph_b200 import Layer
    E = Layer(gpu_ctx, 0.05, max_blocks=16)
    E.generateMesh()
    assert E.getConnectedMesh(fetch=False) == (0, 0)
