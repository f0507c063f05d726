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
