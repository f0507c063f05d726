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


def jump_bits(t0, ts, k):
    """The same in integer arithmetic on the float32 bit patterns (what a CUDA kernel would do: no
    float64, one real addition per binade crossing).  Handles t0 <= 0 (the 1e-6 of the grid index
    can make the first boundary slightly negative) by real additions until t is positive."""
    t, left = f32(t0), k
    sb = int(f32(ts).view(np.uint32))
    s_exp, s_man = (sb >> 23) & 0xFF, (sb & 0x7FFFFF) | 0x800000      # ts = s_man * 2^(s_exp - 150)
    while left > 0:
        tb = int(t.view(np.uint32))
        t_exp = (tb >> 23) & 0xFF
        if t <= 0 or t_exp == 0 or s_exp == 0:      # non-positive or denormal: plain additions
            t = f32(t + ts)
            left -= 1
            continue
        t_man = (tb & 0x7FFFFF) | 0x800000
        shift = t_exp - s_exp                       # ulp(t) = 2^shift * ulp(ts)
        if shift < 0 or shift == 0:                 # the step is at least as coarse as t: crossing soon
            t = f32(t + ts)
            left -= 1
            continue
        if shift > 25:                              # the step is below half an ulp of t: t stalls
            return t
        whole, rem, half = s_man >> shift, s_man & ((1 << shift) - 1), 1 << (shift - 1)
        if rem == half:                             # ties: the first one for real, then constant
            t_next = f32(t + ts)
            left -= 1
            if left == 0 or ((int(t_next.view(np.uint32)) >> 23) & 0xFF) != t_exp:
                t = t_next
                continue
            t = t_next
            t_man = (int(t.view(np.uint32)) & 0x7FFFFF) | 0x800000
            inc = whole if whole % 2 == 0 else whole + 1
        else:
            inc = whole + (1 if rem > half else 0)
        if inc == 0:
            return t
        n = min(left, (0xFFFFFF - t_man) // inc)    # additions that keep the mantissa below 2^24
        if n > 0:
            t = np.uint32((t_exp << 23) | ((t_man + n * inc) & 0x7FFFFF)).view(f32)
            left -= n
        if left > 0:                                # the addition that reaches the next binade
            t = f32(t + ts)
            left -= 1
    return t


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
        if c % 3 == 0:                       # a first boundary at or slightly below zero
            t0 = f32(-rng.uniform(0, 2e-5) * float(ts)) if c % 6 == 0 else f32(0.0)
            want = sequential(t0, ts, k)
        got = jump_bits(t0, ts, k)
        assert np.uint32(got.view(np.uint32)) == want.view(np.uint32), ("bits", c, float(t0), float(ts), k)
    print(f"{cases} cases: closed form (float64 and integer-bit versions) == sequential float32 "
          f"accumulation bit for bit; at most {worst} real additions per jump")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20000)
