"""CPU known-answer tests pinning the oracle's ESDF section (EsdfIntegrator::
updateFromTsdfLayerBatch, [EXT] upstream voxblox; call site coxgraph/include/coxgraph/client/
map_server.h:141-145): exact answers behind a planar wall, the quasi-Euclidean bounds around a
sphere, and the order-independence the CUDA path relies on — the sequential bucket-queue run with
min_diff_m = 0 ends, bit for bit, in the fixed point a dense numpy Jacobi iteration computes."""
import numpy as np

from oracle import oracle_py as orc
from tests import util

VS = 0.05


def _layer_from_sdf(sdf, blocks, trunc=0.3, weight=1.0, vs=VS, unobserved=None):
    """TSDF layer over the listed blocks with distance = clip(sdf(centre), +-trunc)."""
    idx = np.array(sorted(blocks, key=lambda b: (b[2], b[1], b[0])), np.int32)
    vox = np.zeros((len(idx), 4096), orc.VOXEL_DTYPE)
    i = np.arange(4096)
    for b, (bx, by, bz) in enumerate(idx):
        x = (bx * 16 + (i & 15) + 0.5) * vs
        y = (by * 16 + ((i >> 4) & 15) + 0.5) * vs
        z = (bz * 16 + (i >> 8) + 0.5) * vs
        vox[b]["distance"] = np.clip(sdf(x, y, z), -trunc, trunc).astype(np.float32)
        w = np.full(4096, weight, np.float32)
        if unobserved is not None:
            w[unobserved(x, y, z)] = 0.0
        vox[b]["weight"] = w
    L = orc.Layer(vs)
    L.upload(idx, vox)
    return L, idx, vox


def test_planar_wall_gives_exact_multiples():
    L, idx, vox = _layer_from_sdf(lambda x, y, z: 1.6 - x, [(bx, by, 0) for bx in range(4) for by in range(2)])
    e = L.esdf_batch()
    assert np.array_equal(e["idx"], idx)
    d = e["distance"].reshape(len(idx), 16, 16, 16)  # [b, z, y, x]
    fl = e["flags"].reshape(len(idx), 16, 16, 16)
    assert (e["flags"] & 1).all() and not (e["flags"] & 6).any()  # observed, never left in the queue
    # the band |tsdf| < min_distance (0.2 m) is copied and fixed
    tsdf = vox["distance"].reshape(len(idx), 16, 16, 16)
    band = np.abs(tsdf) < np.float32(0.2)
    assert np.array_equal((fl & 8) != 0, band)
    assert np.array_equal(d[band], tsdf[band])
    # outside the band the wavefront walks along x: one voxel per hop from the last fixed voxel
    for b, (bx, by, bz) in enumerate(idx):
        x = (bx * 16 + np.arange(16) + 0.5) * VS
        expect = np.clip(1.6 - x, -2.0, 2.0)
        row = d[b, 3, 5, :]
        assert np.abs(row - expect).max() < 2e-6, (bx, row, expect)
    # every row is the same (a wall): no dependence on y, z
    assert np.array_equal(d[:, :, :, :], np.broadcast_to(d[:, :1, :1, :], d.shape))


def test_sphere_is_within_the_quasi_euclidean_bounds():
    c = np.array([1.6, 1.6, 1.6])
    blocks = [(x, y, z) for x in range(4) for y in range(4) for z in range(4)]
    L, idx, vox = _layer_from_sdf(
        lambda x, y, z: np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - 0.6, blocks)
    e = L.esdf_batch(orc.default_esdf_config(max_distance_m=1.5, default_distance_m=1.5))
    i = np.arange(4096)
    worst_over, worst_under = 0.0, 0.0
    for b, (bx, by, bz) in enumerate(idx):
        x = (bx * 16 + (i & 15) + 0.5) * VS
        y = (by * 16 + ((i >> 4) & 15) + 0.5) * VS
        z = (bz * 16 + (i >> 8) + 0.5) * VS
        true = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - 0.6
        d = e["distance"][b]
        sel = (np.abs(d) < 1.45)
        assert np.array_equal(np.sign(d[sel]), np.sign(true[sel]))
        # a chamfer path is never shorter than the straight line (minus the half-voxel the fixed
        # band's own sampling allows) and at most ~8 % longer
        worst_under = max(worst_under, float((np.abs(true[sel]) - np.abs(d[sel])).max()))
        worst_over = max(worst_over, float((np.abs(d[sel]) - 1.09 * np.abs(true[sel])).max()))
    assert worst_under < VS and worst_over < VS, (worst_under, worst_over)


def _scene():
    # an L-shaped wall with a hole of unobserved voxels and a missing block
    def sdf(x, y, z):
        return np.minimum(1.3 - x, 1.1 - y)
    blocks = [(x, y, z) for x in range(3) for y in range(3) for z in range(2) if (x, y, z) != (0, 0, 1)]
    return _layer_from_sdf(sdf, blocks,
                           unobserved=lambda x, y, z: ((x - 0.6) ** 2 + (y - 0.5) ** 2 + (z - 0.4) ** 2) < 0.04)


def test_sequential_queue_with_zero_threshold_ends_in_the_dense_fixed_point():
    L, idx, vox = _scene()
    for over in ({}, {"max_distance_m": 0.7, "default_distance_m": 0.7}, {"add_occupied_crust": 1},
                 {"num_buckets": 3}, {"multi_queue": 1}):
        cfg = orc.default_esdf_config(min_diff_m=0.0, **over)
        e = L.esdf_batch(cfg)
        dist, observed, fixed, crust = util.esdf_fixed_point_numpy(idx, vox, cfg, VS)
        assert np.array_equal((e["flags"] & 1) != 0, observed), over
        assert np.array_equal((e["flags"] & 8) != 0, fixed), over
        assert np.array_equal((e["flags"] & 2) != 0, crust), over
        assert np.array_equal(e["distance"].view(np.uint32), dist.view(np.uint32)), over
    assert (~observed).any() and fixed.any()


def test_default_threshold_stays_within_its_bound_of_the_fixed_point():
    L, idx, vox = _scene()
    exact = L.esdf_batch(orc.default_esdf_config(min_diff_m=0.0))
    dflt = L.esdf_batch()  # min_diff_m = 1e-3, 20 buckets
    assert np.array_equal(exact["flags"], dflt["flags"])
    gap = np.abs(dflt["distance"]) - np.abs(exact["distance"])
    assert gap.min() >= 0.0          # the early stop never undershoots
    assert gap.max() < 5e-3          # and stays within a few thresholds
    # parents of lowered voxels point at the neighbour the distance came from
    lowered = ((dflt["flags"] & 9) == 1) & (np.abs(dflt["distance"]) < 2.0)
    assert (np.abs(dflt["parent"][lowered]).sum(axis=-1) > 0).all()
    assert (dflt["parent"][~lowered] == 0).all()


def test_free_pointcloud_lists_observed_voxels_beyond_the_radius():
    L, idx, vox = _scene()
    e = L.esdf_batch(free_min_distance=0.5)
    sel = ((e["flags"] & 1) != 0) & (e["distance"] >= np.float32(0.5))
    pts = e["free_points"]
    assert len(pts) == int(sel.sum()) > 0
    assert np.array_equal(pts[:, 3], e["distance"][sel])
    b, lin = np.nonzero(sel)
    centre = (idx[b].astype(np.float64) * 16 + np.stack([lin & 15, (lin >> 4) & 15, lin >> 8], -1) + 0.5) * VS
    assert np.abs(pts[:, :3] - centre).max() < 1e-5


def test_order_independence_on_random_scenes():
    """Random smooth fields with unobserved pockets and missing blocks: whatever the bucket layout
    of the sequential queue, a zero threshold ends in the dense fixed point — the property that
    lets a block-parallel relaxation be compared with the sequential integrator bit for bit."""
    for seed in range(4):
        rng = np.random.default_rng(100 + seed)
        ctrs = rng.uniform(0.2, 1.4, size=(3, 3))
        rad = rng.uniform(0.15, 0.45, size=3)
        hole = rng.uniform(0.3, 1.3, size=3)

        def sdf(x, y, z):
            d = np.full(x.shape, 10.0)
            for c, r in zip(ctrs, rad):
                d = np.minimum(d, np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r)
            return d
        blocks = [(x, y, z) for x in range(2) for y in range(2) for z in range(2)]
        blocks.pop(int(rng.integers(len(blocks))))
        L, idx, vox = _layer_from_sdf(
            sdf, blocks, trunc=float(rng.uniform(0.12, 0.3)),
            unobserved=lambda x, y, z: ((x - hole[0]) ** 2 + (y - hole[1]) ** 2 + (z - hole[2]) ** 2) < 0.03)
        md = float(rng.uniform(0.04, 0.15))
        mx = float(rng.uniform(0.4, 1.5))
        for nb in (2, 20):
            cfg = orc.default_esdf_config(min_diff_m=0.0, min_distance_m=md, max_distance_m=mx,
                                          default_distance_m=mx, num_buckets=nb)
            e = L.esdf_batch(cfg)
            dist, observed, fixed, _ = util.esdf_fixed_point_numpy(idx, vox, cfg, VS)
            assert np.array_equal((e["flags"] & 1) != 0, observed), (seed, nb)
            assert np.array_equal((e["flags"] & 8) != 0, fixed), (seed, nb)
            assert np.array_equal(e["distance"].view(np.uint32), dist.view(np.uint32)), (seed, nb)
