"""Mesh recovery: voxblox::MeshConverter + the TsdfRecover::processMesh loop.
This is the one part of the path whose source IS in the reference tree
(coxgraph/include/coxgraph/map_comm/mesh_converter.h, tsdf_recover.h:59-99), so the oracle's
restatement is checked here against hand-computed values of the reference's formulas (CPU), and
the CUDA path against the oracle, bit for bit (GPU)."""
import numpy as np
import pytest

from tests import util


def make_mesh(seed=0, blocks=6, tris_per_block=40, stamps=8, edge=0.8):
    """A synthetic voxblox_msgs/Mesh with observation history: small random triangles."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(-3, 4, (blocks, 3)).astype(np.int32)
    has = np.ones(blocks, np.uint8)
    has[blocks // 2] = 0                       # one block without history is skipped (:87)
    vb = (np.arange(blocks + 1) * 3 * tris_per_block).astype(np.uint32)
    T = blocks * tris_per_block
    base = rng.integers(0, 30000, (T, 1, 3))
    xyz = (base + rng.integers(0, 9000, (T, 3, 3))).reshape(-1, 3).astype(np.uint16)
    rgb = rng.integers(0, 256, (3 * T, 3)).astype(np.uint8)
    hist, hb = [], [0]
    for t in range(T):
        for _ in range(int(rng.integers(0, 3))):    # 0..2 (first, last) ranges per triangle
            a = int(rng.integers(0, stamps))
            hist += [a, min(stamps + 1, a + int(rng.integers(0, 3)))]
        hb.append(len(hist))
    hist[1] = hist[0] + 300                      # a range that wraps the uint8 key (:274)
    return dict(block_index=idx, block_has_history=has, vertex_begin=vb, x=xyz[:, 0], y=xyz[:, 1],
                z=xyz[:, 2], r=rgb[:, 0], g=rgb[:, 1], b=rgb[:, 2],
                hist_begin=np.array(hb, np.uint32), hist=np.array(hist, np.uint32),
                block_edge_length=edge)


def make_trajectory(n=10, seed=1):
    rng = np.random.default_rng(seed)
    poses = []
    for k in range(n):
        yaw = 0.1 * k
        poses.append([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2), 0.2 * k - 1.0, 0.1 * k, 0.5])
    stamps = 100.0 + 0.05 * np.arange(n) + rng.uniform(-0.01, 0.01, n)
    stamps[3] = stamps[2] + 0.004                # two poses that fall into the same bucket
    return np.array(poses, np.float32), stamps


def test_oracle_follows_mesh_converter_formulas():
    from oracle import oracle_py as orc
    # one block, one triangle, observed at stamp 0 only
    mesh = dict(block_index=np.array([[1, -2, 0]], np.int32), block_has_history=np.array([1], np.uint8),
                vertex_begin=np.array([0, 3], np.uint32),
                x=np.array([0, 6554, 0], np.uint16), y=np.array([0, 0, 9830], np.uint16),
                z=np.array([32768, 32768, 32768], np.uint16),
                r=np.array([255, 0, 0], np.uint8), g=np.array([0, 255, 0], np.uint8),
                b=np.array([0, 0, 255], np.uint8), hist_begin=np.array([0, 2], np.uint32),
                hist=np.array([0, 0], np.uint32), block_edge_length=0.8)
    poses = np.array([[1, 0, 0, 0, 0, 0, 0]], np.float32)
    offs, pts, cols = orc.mesh_to_frames(mesh, 0.05, poses, np.array([5.0]))
    f = np.float32(2.0) / np.float32(65535)
    e = np.float32(0.8)
    p0 = np.array([(np.float32(0) * f + np.float32(1)) * e, (np.float32(0) * f - np.float32(2)) * e,
                   (np.float32(32768) * f + np.float32(0)) * e], np.float32)
    p1 = p0.copy()
    p1[0] = (np.float32(6554) * f + np.float32(1)) * e          # 0.16 m along x
    assert np.array_equal(pts[0], p0) and np.array_equal(pts[1], p1)
    # edge p0-p1 (0.16 m) sampled every 5 cm strictly inside: 3 points (:224-230)
    n01 = int(np.sum(np.arange(1, 10) * np.float32(0.05) < np.float32(np.linalg.norm(p1 - p0))))
    assert n01 == 3 and np.allclose(pts[3:6, 0] - p0[0], [0.05, 0.10, 0.15], atol=1e-6)
    # then the centroid with colour blend(c2, 1/3, blend(c0, .5, c1, .5), 2/3) (:246-249)
    c = pts[3 + n01]
    assert np.allclose(c, (pts[0] + pts[1] + pts[2]) / 3, atol=1e-6)
    assert tuple(cols[3 + n01]) == (85, 85, 85, 255)
    # edge p0-p2 blends colors[0] with colors[1] (sic, :235-236): no blue on it
    n02 = int(np.sum(np.arange(1, 10) * np.float32(0.05) < np.float32(np.linalg.norm(pts[2] - p0))))
    e02 = cols[3 + n01 + 1: 3 + n01 + 1 + n02]
    assert n02 == 4 and (e02[:, 2] == 0).all() and (e02[:, 1] > 0).all()
    assert offs.tolist() == [0, len(pts)]
    # the cloud is handed over in the sensor frame: T_G_C^-1 * p (:203-204)
    T = np.array([[np.cos(0.3), 0, 0, np.sin(0.3), 1.0, -2.0, 0.25]], np.float32)
    _, pts_c, _ = orc.mesh_to_frames(mesh, 0.05, T, np.array([5.0]))
    from coxgraph_b200 import synth
    R = synth.matrix_from_quat(T[0, :4].astype(np.float64))
    assert np.allclose(pts_c, (pts.astype(np.float64) - T[0, 4:]) @ R, atol=1e-5)


def test_oracle_buckets_by_uint8_stamp_and_pose_time():
    from oracle import oracle_py as orc
    mesh = make_mesh()
    poses, stamps = make_trajectory()
    offs, pts, cols = orc.mesh_to_frames(mesh, 0.05, poses, stamps)
    sizes = np.diff(offs.astype(np.int64))
    assert sizes[2] == sizes[3] and sizes[2] > 0     # poses 2 and 3 share bucket 2
    assert sizes.sum() == len(pts) == len(cols) and (cols[:, 3] == 255).all()
    # the wrapped range (stamps a .. a+300) hits every bucket at least once
    assert (sizes > 0).all()


def _frozen():
    import hashlib
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "digests.json")
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    return json.load(open(here))["mesh_frames"], sha


def test_oracle_reproduces_frozen_mesh_frames():
    from oracle import oracle_py as orc
    want, sha = _frozen()
    poses, stamps = make_trajectory(12, 0)
    offs, pts, cols = orc.mesh_to_frames(make_mesh(seed=0, blocks=9, tris_per_block=60), 0.05, poses,
                                         stamps)
    assert len(pts) == want["num_points"]
    assert (sha(offs), sha(pts), sha(cols)) == (want["offsets_sha256"], want["points_sha256"],
                                               want["colors_sha256"])


@pytest.mark.gpu
def test_cuda_reproduces_frozen_mesh_frames(gpu_ctx):
    from coxgraph_b200 import meshToFrames
    want, sha = _frozen()
    poses, stamps = make_trajectory(12, 0)
    offs, pts, cols = meshToFrames(gpu_ctx, make_mesh(seed=0, blocks=9, tris_per_block=60), 0.05,
                                   poses, stamps)
    assert (sha(offs), sha(pts), sha(cols)) == (want["offsets_sha256"], want["points_sha256"],
                                               want["colors_sha256"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,vs", [(0, 0.05), (3, 0.02), (5, 0.2)])
def test_cuda_mesh_to_frames_is_bit_exact(gpu_ctx, seed, vs):
    from coxgraph_b200 import meshToFrames
    from oracle import oracle_py as orc
    mesh = make_mesh(seed=seed, blocks=9, tris_per_block=60)
    poses, stamps = make_trajectory(12, seed)
    o_offs, o_pts, o_cols = orc.mesh_to_frames(mesh, vs, poses, stamps)
    g_offs, g_pts, g_cols = meshToFrames(gpu_ctx, mesh, vs, poses, stamps)
    assert np.array_equal(g_offs, o_offs)
    assert np.array_equal(g_pts.view(np.uint32), o_pts.view(np.uint32)), "points differ"
    assert np.array_equal(g_cols, o_cols), "colours differ"


@pytest.mark.gpu
def test_cuda_recover_mesh_matches_process_mesh(gpu_ctx):
    """processMesh: clear, convert, integrate every non-empty pose cloud (tsdf_recover.h:59-99)."""
    from coxgraph_b200 import Layer, recoverMesh
    from oracle import oracle_py as orc
    mesh = make_mesh(seed=7, blocks=8, tris_per_block=80)
    poses, stamps = make_trajectory(10, 7)
    over = dict(max_ray_length_m=100.0, min_ray_length_m=0.0, default_truncation_distance=0.15)
    ocfg, gcfg = util.make_cfgs(**over)      # tsdf_recover.yaml: min / max ray 0 / 100
    offs, pts, cols = orc.mesh_to_frames(mesh, 0.05, poses, stamps)
    ol = orc.Layer(0.05)
    for i in range(len(poses)):
        a, b = int(offs[i]), int(offs[i + 1])
        if b > a:
            ol.integrate(ocfg, poses[i], pts[a:b], cols[a:b])
    gl = Layer(gpu_ctx, 0.05, max_blocks=8192)
    gl.upload(np.array([[50, 50, 50]], np.int32), np.zeros((1, 4096), orc.VOXEL_DTYPE))  # cleared first
    st = recoverMesh(gl, gcfg, mesh, 0.05, poses, stamps)
    assert st.points_in == int(offs[-1])
    util.compare_layers(gl.download(), ol.download(), "recovered layer")
    gl.close()
