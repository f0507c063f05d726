"""Known-answer tests that pin the CPU oracle (no GPU).  The reference tree has no tests or golden
vectors for this path, so the oracle is pinned by analytic and algebraic
facts about the algorithms it restates (upstream voxblox test ideas: test_tsdf_integrators,
test_merge_integration, test_tsdf_interpolator)."""
import numpy as np
import pytest

from coxgraph_b200 import synth
from oracle import oracle_py as orc
from tests import util

IDENT = np.array([1, 0, 0, 0, 0, 0, 0], np.float32)


def _plane_cloud(depth=2.0, w=160, h=120, f=150.0):
    u, v = np.meshgrid(np.arange(w), np.arange(h))
    d = np.stack([(u - 80.3) / f, (v - 60.2) / f, np.ones_like(u, dtype=float)], -1).reshape(-1, 3)
    pts = (d * depth).astype(np.float32)
    cols = np.zeros((len(pts), 4), np.uint8)
    cols[:, 0], cols[:, 1], cols[:, 2], cols[:, 3] = 200, 100, 50, 255
    return pts, cols


def test_transform_matches_float64_algebra():
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        t = rng.standard_normal(3) * 5
        T = np.concatenate([q, t]).astype(np.float32)
        p = (rng.standard_normal(3) * 4).astype(np.float32)
        R = synth.matrix_from_quat(T[:4].astype(np.float64))
        ref = R @ p.astype(np.float64) + T[4:].astype(np.float64)
        got = orc.transform_point(T, p)
        assert np.allclose(got, ref, rtol=0, atol=5e-6)
        Ti = orc.inverse_transform(T)
        back = orc.transform_point(Ti, got)
        assert np.allclose(back, p, atol=1e-5)


def test_raycaster_visits_l1_path():
    rng = np.random.default_rng(2)
    for _ in range(200):
        o = (rng.standard_normal(3) * 2).astype(np.float32)
        p = (o + rng.standard_normal(3) * 3).astype(np.float32)
        vox = orc.cast_ray(o, p, clearing=False, carving=True, max_ray=5.0, voxel_size_inv=20.0,
                           trunc=0.15)
        # steps + 1 indices, consecutive ones differ by one step along one axis
        d = np.abs(np.diff(vox, axis=0))
        assert (d.sum(axis=1) == 1).all()
        assert np.array_equal(vox[0], np.floor(o.astype(np.float32) * np.float32(20.0) + 1e-6))
        assert len(vox) == np.abs(vox[-1] - vox[0]).sum() + 1
        # the end voxel is the one holding p + unit * trunc (up to float rounding of that point)
        end = p + (p - o) / np.linalg.norm(p - o) * 0.15
        assert np.abs(vox[-1] - np.floor(end * 20.0 + 1e-6)).max() <= 1
    # no carving: the walk is confined to the truncation band
    vox = orc.cast_ray([0.01, 0.02, 0.03], [2.0, 0.3, 0.1], False, False, 5.0, 20.0, 0.15)
    assert 4 <= len(vox) <= 12
    # clearing ray: stops max_ray from the origin
    vox = orc.cast_ray([0.01, 0.02, 0.03], [9.0, 0.0, 0.0], True, True, 5.0, 20.0, 0.15)
    assert abs(vox[-1][0] - 100) <= 1
    # cast_from_origin=False walks the same voxels backwards
    a = orc.cast_ray([0.01, 0.02, 0.03], [1.3, 0.7, -0.4], False, True, 5.0, 20.0, 0.15, True)
    b = orc.cast_ray([0.01, 0.02, 0.03], [1.3, 0.7, -0.4], False, True, 5.0, 20.0, 0.15, False)
    assert np.array_equal(a[0], b[-1]) and np.array_equal(a[-1], b[0]) and len(a) == len(b)


@pytest.mark.parametrize("method", [0, 1, 2])
def test_plane_scene_gives_analytic_sdf(method):
    """Fronto-parallel plane at depth 2 m: along the optical axis the TSDF is depth - z."""
    pts, cols = _plane_cloud()
    T = np.array([1, 0, 0, 0, 0.013, 0.021, 0.007], np.float32)
    cfg = orc.default_config(default_truncation_distance=0.2, use_const_weight=1, method=method,
                             use_weight_dropoff=0)
    L = orc.Layer(0.05)
    L.integrate(cfg, T, pts, cols)
    idx, vox, _ = L.download()
    checked = 0
    for b, (bx, by, bz) in enumerate(idx):
        if bx != 0 or by != 0:
            continue
        for lz in range(16):
            v = vox[b, 0 + 16 * (0 + 16 * lz)]
            if v["weight"] <= 0:
                continue
            zc = (bz * 16 + lz + 0.5) * 0.05
            expect = np.clip(2.007 - zc, -0.2, 0.2)
            assert abs(v["distance"] - expect) < 1.5e-3
            if abs(expect) < 0.19:
                assert tuple(v["rgba"]) == (200, 100, 50, 255)
            checked += 1
    assert checked > 35
    # free space in front of the plane saturates at +truncation, nothing beyond -truncation
    assert np.isclose(vox["distance"].max(), 0.2) and vox["distance"].min() >= -0.2 - 1e-6


def test_weights_and_clamps():
    pts, cols = _plane_cloud()
    cfg = orc.default_config(default_truncation_distance=0.2, use_const_weight=1, method=1,
                             max_weight=3.5)
    L = orc.Layer(0.05)
    for _ in range(5):
        L.integrate(cfg, IDENT, pts, cols)
    _, vox, _ = L.download()
    assert vox["weight"].max() == np.float32(3.5)          # min(max_weight, W + w)
    assert np.abs(vox["distance"]).max() <= np.float32(0.2)


def test_validity_rules():
    cfg = orc.default_config(min_ray_length_m=0.5, max_ray_length_m=2.0, allow_clear=0,
                             use_const_weight=1)
    # (exactly axis-aligned rays are avoided: upstream's DDA divides by the zero ray component)
    pts = np.array([[0.01, 0.02, 0.3], [0.013, 0.021, 1.0], [0.02, 0.03, 3.0], [np.nan, 0, 1]],
                   np.float32)
    cols = np.full((4, 4), 255, np.uint8)
    L = orc.Layer(0.1)
    L.integrate(cfg, IDENT, pts, cols)
    _, vox, _ = L.download()
    zmax = 0.0
    idx, _, _ = L.download()
    for b, bi in enumerate(idx):
        w = vox[b]["weight"].reshape(16, 16, 16)
        zs = np.nonzero(w.sum(axis=(1, 2)))[0]
        if len(zs):
            zmax = max(zmax, (bi[2] * 16 + zs.max() + 1) * 0.1)
    assert 1.0 < zmax <= 1.0 + cfg.default_truncation_distance + 0.11   # only the 1 m point
    cfg.allow_clear = 1                                              # 3 m point clears to 2 m
    L2 = orc.Layer(0.1)
    L2.integrate(cfg, IDENT, pts, cols)
    assert L2.num_blocks >= L.num_blocks
    _, v2, _ = L2.download()
    assert v2["weight"].sum() > vox["weight"].sum()


def test_merged_bundles_use_reference_visit_order():
    """Two points in one voxel: the merged ray is their weighted mean (const weight)."""
    cfg = orc.default_config(use_const_weight=1, method=1, default_truncation_distance=0.1)
    pts = np.array([[0.011, 0.012, 1.013], [0.021, 0.022, 1.021]], np.float32)
    cols = np.array([[10, 20, 30, 255], [20, 40, 60, 255]], np.uint8)
    L = orc.Layer(0.05)
    L.integrate(cfg, IDENT, pts, cols)
    _, vox, _ = L.download()
    w = vox["weight"]
    assert w.max() == 2.0                   # one bundle of weight 2, not two rays of weight 1
    near = vox[(w > 0) & (np.abs(vox["distance"]) < 0.09)]
    assert len(near) and all(tuple(c) == (15, 30, 45, 255) for c in near["rgba"])


def _fused_submap(frames=2, stride=8, robot=0):
    ocfg, _ = util.make_cfgs()
    L = orc.Layer(0.05)
    for (T, p, c) in util.small_frames(frames, stride=stride, robot=robot):
        L.integrate(ocfg, T, p, c)
    return L


def test_interpolator_reproduces_voxel_centres_and_linear_fields():
    L = orc.Layer(0.1)
    idx = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1],
                    [1, 1, 1]], np.int32)
    vox = np.zeros((8, 4096), orc.VOXEL_DTYPE)
    lin = np.arange(4096)
    lx, ly, lz = lin % 16, (lin // 16) % 16, lin // 256
    for b, bi in enumerate(idx):
        x = (bi[0] * 16 + lx + 0.5) * 0.1
        y = (bi[1] * 16 + ly + 0.5) * 0.1
        z = (bi[2] * 16 + lz + 0.5) * 0.1
        vox[b]["distance"] = 0.3 * x - 0.2 * y + 0.1 * z + 0.05     # affine field
        vox[b]["weight"] = 2.0
        vox[b]["rgba"] = 100
    L.upload(idx, vox)
    rng = np.random.default_rng(3)
    for _ in range(200):
        p = rng.uniform(0.2, 3.0, 3).astype(np.float32)
        ok, d, w, rgba = L.interp(p, True)
        assert ok and abs(d - (0.3 * p[0] - 0.2 * p[1] + 0.1 * p[2] + 0.05)) < 2e-5
        assert abs(w - 2.0) < 1e-5 and 99 <= rgba[0] <= 100
    ok, d, w, _ = L.interp([0.25, 0.35, 0.45], True)               # exactly a voxel centre
    assert ok and abs(d - (0.075 - 0.07 + 0.045 + 0.05)) < 1e-6
    assert not L.interp([-0.3, 0.2, 0.2], True)[0]                # outside: no block
    ok, d, _, _ = L.interp([0.01, 0.01, 0.01], False)              # nearest at the corner voxel
    assert ok and abs(d - vox[0]["distance"][0]) < 1e-7
    assert not L.interp([0.01, 0.01, 0.01], True)[0]               # trilinear needs the -1 block


def test_merge_identity_doubles_weight_keeps_distance():
    A = _fused_submap()
    B = orc.Layer(0.05)
    ai, av, _ = A.download()
    B.upload(ai, av)
    B.merge_from(A, IDENT)
    bi, bv, bf = B.download()
    assert np.array_equal(bi, ai)
    obs = av["weight"] > 1e-6
    assert np.allclose(bv["weight"][obs], 2 * av["weight"][obs], rtol=1e-6)
    assert np.allclose(bv["distance"][obs], av["distance"][obs], atol=2e-6)
    assert (np.abs(bv["rgba"][obs].astype(int) - av["rgba"][obs].astype(int)) <= 1).all()
    assert (bf & 1).all()


def test_merge_whole_block_translation_shifts_blocks():
    A = _fused_submap()
    shift = np.array([2, -1, 1])
    T = np.array([1, 0, 0, 0, *(shift * 0.8)], np.float32)
    B = orc.Layer(0.05)
    B.merge_from(A, T)
    ai, av, _ = A.download()
    bi, bv, _ = B.download()
    has = (av["weight"] > 1e-6).any(axis=1)
    want = {tuple(i + shift) for i in ai[has]}
    got = {tuple(i) for i in bi}
    assert want == got
    amap = {tuple(i + shift): v for i, v in zip(ai, av)}
    for i, v in zip(bi, bv):
        src = amap[tuple(i)]
        obs = src["weight"] > 1e-6
        # the shifted voxel centres are off by float rounding (~1e-7 m), so the trilinear sample
        # sits a hair beside the source voxel
        assert np.allclose(v["distance"][obs], src["distance"][obs], atol=1e-5)
        assert np.allclose(v["weight"][obs], src["weight"][obs], rtol=2e-4, atol=1e-4)


def test_merge_rotation_by_90_degrees_permutes_voxels():
    A = _fused_submap()
    s = np.float32(np.sqrt(0.5))
    T = np.array([s, 0, 0, s, 0, 0, 0], np.float32)      # +90 deg about z: (x,y,z)->(-y,x,z)
    B = orc.Layer(0.05)
    B.merge_from(A, T)
    ai, av, _ = A.download()
    bi, bv, _ = B.download()
    bmap = {tuple(i): v for i, v in zip(bi, bv)}
    rng = np.random.default_rng(5)
    checked = 0
    for b in rng.choice(len(ai), size=min(20, len(ai)), replace=False):
        w = av[b]["weight"].reshape(16, 16, 16)          # [z, y, x]
        zs, ys, xs = np.nonzero(w > 1e-6)
        for z, y, x in list(zip(zs, ys, xs))[::97]:
            gx, gy, gz = ai[b][0] * 16 + x, ai[b][1] * 16 + y, ai[b][2] * 16 + z
            rx, ry = -gy - 1, gx                        # voxel centre (x+.5) -> (-(y+.5), x+.5)
            blk = (rx // 16, ry // 16, gz // 16)
            if blk not in bmap:
                continue
            v = bmap[blk][(rx % 16) + 16 * ((ry % 16) + 16 * (gz % 16))]
            src = av[b][x + 16 * (y + 16 * z)]
            if v["weight"] > 1e-6:
                assert abs(v["distance"] - src["distance"]) < 1e-4
                checked += 1
    assert checked > 50


def test_aligned_merge_and_thread_variants_agree():
    A = _fused_submap(robot=0)
    B = _fused_submap(robot=1)
    T = synth.robot_map_offset(1)
    G1, G2 = orc.Layer(0.05), orc.Layer(0.05)
    G1.merge_from(A, T)
    G2.merge_from(A, T, threads=4)
    a, b = G1.download(), G2.download()
    util.compare_layers(a, b, "threaded merge", exact=True, check_flags=True)
    # aligned merge == voxel-wise fold
    P1, P2, S = orc.Layer(0.05), orc.Layer(0.05), orc.Layer(0.05)
    P1.merge_from(A, IDENT)
    P2.merge_from(B, T)
    S.merge_aligned_from(P1)
    S.merge_aligned_from(P2)
    D = orc.Layer(0.05)
    D.merge_from(A, IDENT)
    D.merge_from(B, T)
    util.compare_layers(S.download(), D.download(), "aligned fold")


def test_threaded_integrator_allocates_the_same_blocks():
    ocfg, _ = util.make_cfgs()
    (T, p, c), = util.small_frames(1, stride=8)
    a, b = orc.Layer(0.05), orc.Layer(0.05)
    a.integrate(ocfg, T, p, c)
    b.integrate(ocfg, T, p, c, threads=4)
    assert np.array_equal(a.download()[0], b.download()[0])
    assert np.isclose(a.download()[1]["weight"].sum(), b.download()[1]["weight"].sum(), rtol=1e-5)


def test_fast_integrator_is_deterministic_subset():
    cfg_f = orc.default_config(**dict(util.CFG_FIELDS, method=2))
    cfg_s = orc.default_config(**dict(util.CFG_FIELDS, method=0))
    (T, p, c), = util.small_frames(1, stride=8)
    f1, f2, s = orc.Layer(0.05), orc.Layer(0.05), orc.Layer(0.05)
    f1.integrate(cfg_f, T, p, c)
    f2.integrate(cfg_f, T, p, c)
    s.integrate(cfg_s, T, p, c)
    util.compare_layers(f1.download(), f2.download(), "fast twice", exact=True)
    fast_blocks = {tuple(i) for i in f1.download()[0]}
    simple_blocks = {tuple(i) for i in s.download()[0]}
    assert fast_blocks <= simple_blocks and len(fast_blocks) > 0.5 * len(simple_blocks)


def test_default_alpha_switch(tmp_path):
    """The alpha byte of a default-constructed Color is a build switch (ORC_DEFAULT_ALPHA /
    CG_DEFAULT_ALPHA): with 0 instead of 255 only alpha bytes change."""
    import ctypes as C
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    so = str(tmp_path / "liborc_alpha0.so")
    subprocess.check_call(["g++", "-O1", "-DORC_DEFAULT_ALPHA=0", "-std=c++17", "-fPIC",
                           "-ffp-contract=off", "-fno-fast-math", "-pthread", "-shared", "-o", so,
                           os.path.join(here, "tsdf_oracle.cc")])
    alt = C.CDLL(so)
    alt.orc_layer_create.restype = C.c_void_p
    alt.orc_layer_create.argtypes = [C.c_float, C.c_int32]
    alt.orc_layer_num_blocks.restype = C.c_size_t
    alt.orc_layer_num_blocks.argtypes = [C.c_void_p]
    alt.orc_layer_download.argtypes = [C.c_void_p] * 4
    alt.orc_integrate_pointcloud.restype = C.c_int32
    alt.orc_integrate_pointcloud.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
    cfg = orc.default_config(default_truncation_distance=0.15, use_const_weight=1, method=1)
    (T, pts, cols), = util.small_frames(1, stride=16)
    ref = orc.Layer(0.05)
    ref.integrate(cfg, T, pts, cols)
    ri, rv, _ = ref.download()
    h = alt.orc_layer_create(0.05, 16)
    T = np.ascontiguousarray(T, np.float32)
    assert alt.orc_integrate_pointcloud(h, C.byref(cfg), T.ctypes.data, pts.ctypes.data,
                                        cols.ctypes.data, len(pts), 0, None) == 0
    n = alt.orc_layer_num_blocks(h)
    ai = np.zeros((n, 3), np.int32)
    av = np.zeros((n, 4096), orc.VOXEL_DTYPE)
    af = np.zeros(n, np.uint8)
    alt.orc_layer_download(h, ai.ctypes.data, av.ctypes.data, af.ctypes.data)
    assert np.array_equal(ai, ri)
    assert np.array_equal(av["distance"], rv["distance"]) and np.array_equal(av["weight"], rv["weight"])
    assert np.array_equal(av["rgba"][..., :3], rv["rgba"][..., :3])
    differs = av["rgba"][..., 3] != rv["rgba"][..., 3]
    # voxels that only ever saw free space keep the default colour; a voxel carved first and
    # coloured later blends the default alpha in: alpha bytes differ, nothing else does
    assert differs.any()
    assert (av["rgba"][..., 3][differs] < rv["rgba"][..., 3][differs]).all()
    untouched_colour = (rv["rgba"][..., :3] == 0).all(axis=-1) & (rv["weight"] > 0) & differs
    assert (av["rgba"][..., 3][untouched_colour] == 0).all()
