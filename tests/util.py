"""Shared helpers of the parity tests: scenes, configs, and the layer comparison with the
tolerances BASELINE.json's north_star states (block set bit-exact; distance / weight within
1e-4 relative or 1e-5 absolute; colours within 1 LSB)."""
import numpy as np

from coxgraph_b200 import synth

RTOL = 1e-4
ATOL = 1e-5

CFG_FIELDS = dict(default_truncation_distance=0.15, max_ray_length_m=5.0, min_ray_length_m=0.1,
                  use_const_weight=1, method=1)


def make_cfgs(**over):
    """Same integrator config for the oracle and for the CUDA library."""
    from coxgraph_b200 import TsdfIntegratorConfig
    from oracle import oracle_py as orc
    f = dict(CFG_FIELDS)
    f.update(over)
    return orc.default_config(**f), TsdfIntegratorConfig(**f)


def small_frames(num_frames, stride=8, robot=0, submap=0, cam=synth.CAM_640x480):
    return [(T, p.numpy(), c.numpy())
            for (T, p, c) in synth.submap_frames(robot, submap, num_frames, cam=cam, stride=stride)]


def compare_layers(got, ref, what="layer", exact=False, check_flags=False):
    gi, gv, gf = got
    ri, rv, rf = ref
    assert gi.shape == ri.shape, f"{what}: {len(gi)} blocks vs oracle {len(ri)}"
    assert np.array_equal(gi, ri), f"{what}: allocated block index sets differ"
    if check_flags:
        assert np.array_equal(gf, rf), f"{what}: block flags differ"
    gd, rd = gv["distance"], rv["distance"]
    gw, rw = gv["weight"], rv["weight"]
    if exact:
        assert np.array_equal(gd.view(np.uint32), rd.view(np.uint32)), f"{what}: distance bits"
        assert np.array_equal(gw.view(np.uint32), rw.view(np.uint32)), f"{what}: weight bits"
        assert np.array_equal(gv["rgba"], rv["rgba"]), f"{what}: colours"
        return
    bad_d = np.abs(gd - rd) > np.maximum(ATOL, RTOL * np.abs(rd))
    bad_w = np.abs(gw - rw) > np.maximum(ATOL, RTOL * np.abs(rw))
    assert not bad_d.any(), (f"{what}: {bad_d.sum()} distances out of tolerance, max abs diff "
                             f"{np.abs(gd - rd).max()}")
    assert not bad_w.any(), (f"{what}: {bad_w.sum()} weights out of tolerance, max abs diff "
                             f"{np.abs(gw - rw).max()}")
    dc = np.abs(gv["rgba"].astype(np.int16) - rv["rgba"].astype(np.int16))
    assert dc.max(initial=0) <= 1, f"{what}: colour differs by {dc.max()} LSB"


def exact_fraction(got, ref):
    gv, rv = got[1], ref[1]
    same = (gv["distance"].view(np.uint32) == rv["distance"].view(np.uint32)) & \
           (gv["weight"].view(np.uint32) == rv["weight"].view(np.uint32)) & \
           (gv["rgba"] == rv["rgba"]).all(axis=-1)
    return float(same.mean()) if same.size else 1.0


def margins(got, ref):
    """How far inside the tolerance a CUDA layer is: max |d| / tol for distance and weight
    (tol = max(ATOL, RTOL |ref|)), the colour LSB histogram and the bit-exact fraction."""
    gi, gv, _ = got
    ri, rv, _ = ref
    out = {"blocks": int(len(gi)), "block_sets_equal": bool(gi.shape == ri.shape and np.array_equal(gi, ri))}
    if not out["block_sets_equal"]:
        return out
    for name in ("distance", "weight"):
        g, r = gv[name].astype(np.float64), rv[name].astype(np.float64)
        tol = np.maximum(ATOL, RTOL * np.abs(r))
        out[f"max_{name}_err_over_tol"] = float((np.abs(g - r) / tol).max(initial=0.0))
    dc = np.abs(gv["rgba"].astype(np.int16) - rv["rgba"].astype(np.int16)).max(axis=-1)
    out["colour_lsb_hist"] = [int(x) for x in np.bincount(dc.ravel(), minlength=3)[:8]]
    out["exact_fraction"] = exact_fraction(got, ref)
    return out


def record_margins(name, m):
    """Append the margins of a parity case to gpurun_out/parity_margins.json (brought back from
    the GPU box; summarised under profiles/)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "gpurun_out", "parity_margins.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = m
        with open(path, "w") as f:
            json.dump(data, f, indent=1)
    except OSError:
        pass
    print(f"[parity margins] {name}: {m}")


def esdf_fixed_point_numpy(idx, vox, cfg, voxel_size):
    """The fixed point of EsdfIntegrator's relaxation (min_diff_m = 0) by dense Jacobi iteration
    in numpy float32 — an independent statement of what both the sequential oracle (any queue
    order) and the block-parallel CUDA kernels must end in, bit for bit.  idx [B,3], vox [B,4096]
    VOXEL_DTYPE -> distance f32 [B,4096] (0 where unobserved), observed bool, fixed bool."""
    idx = np.asarray(idx, np.int64)
    lo = idx.min(axis=0)
    ext = (idx.max(axis=0) - lo + 1) * 16
    D = np.full((ext[2] + 2, ext[1] + 2, ext[0] + 2), np.nan, np.float32)  # z, y, x + 1-voxel rim
    fixed = np.zeros(D.shape, bool)
    crust = np.zeros(D.shape, bool)
    sl = []
    dflt = np.float32(cfg.default_distance_m)
    for b, bi in enumerate(idx):
        o = (bi - lo) * 16 + 1
        s = (slice(o[2], o[2] + 16), slice(o[1], o[1] + 16), slice(o[0], o[0] + 16))
        sl.append(s)
        d = vox[b]["distance"].reshape(16, 16, 16)
        w = vox[b]["weight"].reshape(16, 16, 16)
        obs = ~(w < np.float32(cfg.min_weight))
        fx = obs & (np.abs(d) < np.float32(cfg.min_distance_m))
        init = np.where(fx, d, np.sign(d).astype(np.float32) * dflt).astype(np.float32)
        if cfg.add_occupied_crust:
            init = np.where(obs, init, -dflt)
            crust[s] = ~obs
        else:
            init = np.where(obs, init, np.float32(np.nan))
        D[s] = init
        fixed[s] = fx
    vs = np.float32(voxel_size)
    w_cls = {1: np.float32(1.0) * vs, 2: np.float32(np.sqrt(np.float32(2.0))) * vs,
             3: np.float32(np.sqrt(np.float32(3.0))) * vs}
    mx = np.float32(cfg.max_distance_m)
    inner = (slice(1, -1),) * 3
    while True:
        sgn = np.where(D < 0, np.float32(-1), np.float32(1))
        mag = np.abs(D)
        best = mag[inner].copy()
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    order = abs(dx) + abs(dy) + abs(dz)
                    if order == 0:
                        continue
                    nb = D[1 + dz:D.shape[0] - 1 + dz, 1 + dy:D.shape[1] - 1 + dy,
                           1 + dx:D.shape[2] - 1 + dx]
                    a = sgn[inner] * nb
                    with np.errstate(invalid="ignore"):
                        ok = (a > 0) & (a < mx)
                        cand = np.where(ok, (a + w_cls[order]).astype(np.float32), np.float32(np.inf))
                        best = np.minimum(best, cand)
        with np.errstate(invalid="ignore"):
            lower = (best < mag[inner]) & ~fixed[inner] & ~np.isnan(D[inner]) & (D[inner] != 0)
        if not lower.any():
            break
        new = D[inner].copy()
        new[lower] = (sgn[inner] * best)[lower]
        D[inner] = new
    dist = np.stack([D[s].reshape(4096) for s in sl])
    observed = ~np.isnan(dist)
    return (np.where(observed, dist, np.float32(0)).astype(np.float32), observed,
            np.stack([fixed[s].reshape(4096) for s in sl]),
            np.stack([crust[s].reshape(4096) for s in sl]))
