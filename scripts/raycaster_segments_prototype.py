#!/usr/bin/env python
"""CPU prototype for DESIGN.md §10(1): the RayCaster state at the entry of every (ray, block)
segment computed directly — no serial voxel-by-voxel walk — and checked against the sequential
walk (the restatement of voxblox::RayCaster the oracle uses, R3) on random rays.

A ray advances on the axis with the smallest t_to_next_boundary (first minimum on ties) and then
adds t_step_size to it.  The value an axis holds before its (i+1)-th step is T_a(i) = t0_a (+) i
additions of ts_a, which scripts/raycaster_closed_form.py evaluates in O(binades).  A step of axis a
with pre-step value T_a(j) comes after exactly those steps of axis b whose pre-step value is smaller
(or equal, when b < a), so the number of b-steps taken before it is a binary search over i.  Hence
the state right after the j-th step of axis a — in particular after every step that crosses a block
face — follows without walking, and all segments of all rays are independent work items.

Generic rays only (all three direction components non-zero, finite); axis-parallel rays keep the
sequential walk.    python scripts/raycaster_segments_prototype.py [rays]
"""
import sys

import numpy as np

from raycaster_closed_form import jump

f32 = np.float32
VPS = 16


def setup(start, end):
    """voxblox::RayCaster::setupRayCaster on scaled coordinates (float32 arithmetic)."""
    start, end = start.astype(f32), end.astype(f32)
    curr = np.floor(start + f32(1e-6)).astype(np.int64)
    last = np.floor(end + f32(1e-6)).astype(np.int64)
    steps = int(np.abs(last - curr).sum())
    ray = (end - start).astype(f32)
    sign = np.sign(ray).astype(np.int64)
    shifted = (start - curr.astype(f32)).astype(f32)
    corrected = np.maximum(0, sign).astype(f32)
    t0 = ((corrected - shifted).astype(f32) / ray).astype(f32)
    ts = (sign.astype(f32) / ray).astype(f32)
    return curr, sign, t0, ts, steps


def walk(curr, sign, t0, ts, steps):
    """Sequential reference: list of (voxel index, t_to_next_boundary) before every step."""
    c, t = curr.copy(), t0.copy()
    out = []
    for _ in range(steps + 1):
        out.append((c.copy(), t.copy()))
        m = 0
        if t[1] < t[m]:
            m = 1
        if t[2] < t[m]:
            m = 2
        c[m] += sign[m]
        t[m] = f32(t[m] + ts[m])
    return out


def T(t0, ts, a, i):
    return jump(t0[a], ts[a], i)[0] if i > 0 else f32(t0[a])


def steps_before(t0, ts, b, a, t_a, limit):
    """Number of steps axis b has taken when axis a is about to step with pre-step value t_a."""
    def before(i):  # does b's (i+1)-th step come first?
        v = T(t0, ts, b, i)
        return v < t_a or (v == t_a and b < a)
    lo, hi = 0, limit + 1          # smallest i in [0, limit + 1] with not before(i)
    while lo < hi:
        mid = (lo + hi) // 2
        if before(mid):
            lo = mid + 1
        else:
            hi = mid
    return lo


def state_after(curr, sign, t0, ts, steps, a, j):
    """State after the j-th step of axis a (j >= 1), or None if the walk ends before it."""
    t_a = T(t0, ts, a, j - 1)
    k = [0, 0, 0]
    k[a] = j
    for b in range(3):
        if b != a:
            k[b] = steps_before(t0, ts, b, a, t_a, steps)
    n = sum(k)
    if n > steps:
        return None
    c = curr + sign * np.array(k)
    t = np.array([T(t0, ts, b, k[b]) for b in range(3)], f32)
    return n, c, t


def check(rng):
    origin = rng.uniform(-3, 3, 3)
    direction = rng.standard_normal(3)
    direction /= np.linalg.norm(direction)
    start = origin * 20.0
    end = (origin + direction * rng.uniform(0.3, 5.0)) * 20.0   # up to 100 voxels
    curr, sign, t0, ts, steps = setup(start, end)
    if (sign == 0).any() or steps == 0:
        return 0
    ref = walk(curr, sign, t0, ts, steps)
    events = 0
    for a in range(3):
        for j in range(1, steps + 1):
            before, after = curr[a] + sign[a] * (j - 1), curr[a] + sign[a] * j
            if before // VPS == after // VPS:
                continue                      # this step stays inside the block on axis a
            st = state_after(curr, sign, t0, ts, steps, a, j)
            if st is None:
                continue
            n, c, t = st
            rc, rt = ref[n]
            assert np.array_equal(c, rc), (a, j, n, c, rc)
            assert np.array_equal(t.view(np.uint32), rt.view(np.uint32)), (a, j, n, t, rt)
            # and the step before it was indeed axis a's j-th
            assert ref[n - 1][0][a] == before
            events += 1
    return events


if __name__ == "__main__":
    rays = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(3)
    total = sum(check(rng) for _ in range(rays))
    print(f"{rays} rays, {total} block entries: direct state == sequential walk (index and float32 "
          f"t_to_next_boundary, bit for bit)")
