"""bench.py's accounting helpers (no GPU): the per-stage byte model, the roofline record over all
stages (library calls included and flagged), and the config object both arms print."""
import json

import bench


def test_stage_bytes_follow_the_design_table():
    kw = dict(n_pts=7_680_000, rays=290_000, pairs=38_000_000, general=1_900_000, blocks=85)
    assert bench.stage_alg_bytes("bundle_sort", **kw) == 16 * 7_680_000
    assert bench.stage_alg_bytes("point_keys", **kw) == 20 * 7_680_000
    assert bench.stage_alg_bytes("gather_sorted", **kw) == 40 * 7_680_000
    assert bench.stage_alg_bytes("finalize", **kw) == 98304 * 85
    assert bench.stage_alg_bytes("merge_resample", b_in=80, b_out=70, **kw) == 49152 * (80 + 140)
    assert bench.stage_alg_bytes("no_such_stage", **kw) == 0.0


def test_roofline_record_picks_the_dominant_stage_over_all_stages():
    steps = 10
    prof = {"bundle_sort": (3.5, 0), "walk_segments": (2.0, 20), "point_keys": (0.9, 10),
            "transfer": (9.0, 0), "fold_wide": (0.0, 10)}
    per_step = dict(n_pts=7_680_000, rays=290_000, pairs=38_000_000, general=1_900_000, blocks=85)
    r = bench.roofline_record(prof, steps, 1.5, 308e6, per_step, traffic_file=None)
    assert r["kernel"] == "bundle_sort" and r["library_kernel"] is True   # not hidden as in round 1
    assert abs(r["ms_per_launch"] - 0.35) < 1e-12
    assert abs(r["achieved"] - 16 * 7_680_000 / 0.35e-3 / 1e9) < 1e-6      # the stage's OWN bytes
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["step_frac"] - 308e6 / 1.5e-3 / 1e9 / r["peak"]) < 1e-12  # the headline fraction
    assert abs(r["library_stages_ms_per_step"] - 0.35) < 1e-12
    assert r["traffic"] is None and r["bound"] == "hbm" and r["unit"] == "GB/s"
    json.dumps(r)


def test_both_arms_print_the_same_config_object():
    assert set(bench.CONFIG) == {"workload", "l2"} and "C2" in bench.CONFIG["workload"]
    src = open(bench.__file__).read()
    assert src.count('"config": dict(CONFIG)') == 2


def test_layer_diff_reports_margins():
    import numpy as np
    dt = np.dtype([("distance", "<f4"), ("weight", "<f4"), ("rgba", "u1", (4,))])
    idx = np.array([[0, 0, 0]], np.int32)
    a = np.zeros((1, 4096), dt)
    a["distance"], a["weight"] = 0.1, 2.0
    b = a.copy()
    b["distance"][0, 5] += 0.5e-5          # half the absolute tolerance... of 1e-5
    b["rgba"][0, 7, 1] = 1                 # one LSB
    d = bench.layer_diff((idx, b, None), (idx, a, None))
    assert d["block_sets_equal"] and d["within_tolerance"]
    assert 0.4 < d["max_distance_err_over_tol"] < 0.6 and d["colour_lsb_hist"][1] == 1
    b["weight"][0, 9] = 2.1
    assert not bench.layer_diff((idx, b, None), (idx, a, None))["within_tolerance"]
    assert not bench.layer_diff((np.array([[1, 0, 0]], np.int32), a, None), (idx, a, None))["block_sets_equal"]
