"""CPU: groundwork for a walk without the serial DDA.  The closed-form
evaluation of k sequential float32 additions and the direct RayCaster state at block entries must
agree with the sequential restatement of voxblox::RayCaster bit for bit."""
import os
import sys

import numpy as np

SCRIPTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts")
sys.path.insert(0, SCRIPTS)


def test_closed_form_accumulation_matches_sequential_additions():
    import raycaster_closed_form as cf
    cf.main(1500)   # asserts inside: float64 and integer-bit versions, incl. t0 <= 0


def test_direct_block_entry_state_matches_the_walk():
    import raycaster_segments_prototype as proto
    rng = np.random.default_rng(11)
    assert sum(proto.check(rng) for _ in range(25)) > 50


def test_prototype_walk_visits_the_oracles_voxels():
    """The prototype's sequential walk is the oracle's RayCaster (orc_cast_ray, R3)."""
    import raycaster_segments_prototype as proto
    from oracle import oracle_py as orc
    rng = np.random.default_rng(5)
    compared = 0
    for _ in range(20):
        origin = rng.uniform(-2, 2, 3).astype(np.float32)
        point = (origin + rng.standard_normal(3) * 1.5).astype(np.float32)
        # carving on, not clearing: start = origin, end = point + unit * trunc (in voxel units)
        unit = (point - origin) / np.linalg.norm(point - origin)
        want = orc.cast_ray(origin, point, False, True, 50.0, 20.0, 0.1)
        if len(want) < 2:
            continue
        # rebuild the same scaled end points the oracle uses, through its own arithmetic
        d = (point - origin).astype(np.float32)
        n = np.float32(np.sqrt(np.float32(d[0] * d[0]) + np.float32(np.float32(d[1] * d[1]) + np.float32(d[2] * d[2]))))
        u = (d / n).astype(np.float32)
        end = (point + u * np.float32(0.1)).astype(np.float32)
        curr, sign, t0, ts, steps = proto.setup((origin * np.float32(20.0)).astype(np.float32),
                                                (end * np.float32(20.0)).astype(np.float32))
        if (sign == 0).any():
            continue
        got = np.array([c for c, _ in proto.walk(curr, sign, t0, ts, steps)])
        assert np.array_equal(got, np.asarray(want).reshape(-1, 3)[:len(got)]) and len(got) == len(want)
        compared += 1
    assert compared >= 10
