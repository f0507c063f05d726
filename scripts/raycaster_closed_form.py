#!/usr/bin/env python
"""Groundwork for DESIGN.md §10(1): voxblox::RayCaster advances t_to_next_boundary by repeated
float32 additions of the constant t_step_size.  Inside one binade of t every such addition moves t
by the same whole number of ulps (round-to-nearest of a constant fraction), so the value after k
additions has a closed form per binade; only the additions that cross into the next binade (and
the first of a run of exact half-ulp ties) have to be carried out for real.  This script checks that claim bit for bit
against the sequential loop on random (t0, step) pairs — CPU only, no GPU needed.

    python scripts/raycaster_closed_form.py [cases]
"""
import sys

import numpy as np

f32 = np.float32


def sequential(t0, ts, k):
    t = f32(t0)
    for _ in range(k):
        t = f32(t + ts)
    return t


def jump(t0, ts, k):
    """Value of k sequential float32 additions t += ts (t0 > 0, ts > 0), and how many of them had
    to be executed as real additions."""
    t, left, real_adds = f32(t0), k, 0
    while left > 0:
        m, e = np.frexp(t)                     # t = m * 2^e, m in [0.5, 1)
        ulp = np.ldexp(1.0, int(e) - 24)       # spacing of float32 numbers in t's binade
        top = np.ldexp(1.0, int(e))            # first value of the next binade
        q = float(ts) / ulp                    # the step in ulps of t (exact in float64)
        whole, frac = divmod(q, 1.0)
        if q < 0.5 or t == 0:                  # steps below half an ulp (t stalls) and t = 0
            t = f32(t + ts)
            left -= 1
            real_adds += 1
            continue
        if frac == 0.5:
            # exact ties round to even: the first addition depends on the parity of t, after it t is
            # even and every further addition in this binade adds `whole` ulps if that is even,
            # `whole + 1` otherwise
            t_next = f32(t + ts)
            left -= 1
            real_adds += 1
            if left == 0 or float(t_next) >= top:
                t = t_next
                continue
            t = t_next
            inc = (whole if whole % 2 == 0 else whole + 1.0) * ulp
        else:
            inc = (whole + (1.0 if frac > 0.5 else 0.0)) * ulp   # what every addition in this binade adds
        # additions that stay below the top of the binade (the sum itself must be < top)
        room = int(np.floor((top - float(t) - float(ts)) / inc)) + 1 if float(t) + float(ts) < top else 0
        n = max(0, min(left, room))
        if n > 0:
            t = f32(float(t) + n * inc)        # exact: a multiple of ulp below `top`
            left -= n
        if left > 0:                           # the addition that crosses into the next binade
            t = f32(t + ts)
            left -= 1
            real_adds += 1
    return t, real_adds


def main(cases):
    rng = np.random.default_rng(7)
    worst = 0
    for c in range(cases):
        steps = int(rng.integers(2, 400))
        ts = f32(1.0 / (steps * rng.uniform(0.3, 3.0)))          # like sign / (end - start)
        t0 = f32(ts * rng.uniform(0.0, 1.0) + 1e-12)             # first boundary inside one step
        k = int(rng.integers(1, steps + 1))
        want = sequential(t0, ts, k)
        got, real_adds = jump(t0, ts, k)
        assert got.view(np.uint32) == want.view(np.uint32), (c, float(t0), float(ts), k, float(got), float(want))
        worst = max(worst, real_adds)
    print(f"{cases} cases: closed form == sequential float32 accumulation bit for bit; "
          f"at most {worst} real additions per jump")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20000)
