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
