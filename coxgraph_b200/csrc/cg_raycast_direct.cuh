// cg_raycast_direct.cuh — RayCaster state without walking (groundwork).
//
// voxblox::RayCaster advances t_to_next_boundary by repeated float additions of t_step_size.
// Inside one binade of t every such addition moves t by the same whole number of ulps, so the value
// after k additions has a closed form per binade (accumulate_steps: one real addition per binade
// crossing, one at the start of a run of half-ulp ties).  A step of axis a with pre-step value T
// comes after exactly those steps of axis b whose pre-step value is smaller (or equal, when b < a:
// Eigen minCoeff takes the first minimum), so the number of b-steps taken before it is a binary
// search (steps_before), and the state right after the j-th step of axis a — in particular after
// every step that crosses a block face — follows directly (raycast_state_after).  Checked against
// the sequential RayCaster by cg_debug_selftest(which = 2) and, on the CPU, by
// scripts/raycaster_closed_form.py / raycaster_segments_prototype.py.  Not yet used by the walk.
#pragma once
#include "cg_math.cuh"

namespace cg {

// t after `left` sequential additions t = t + ts (ts > 0 finite), bit for bit
__device__ __forceinline__ float accumulate_steps(float t, float ts, unsigned left) {
  const uint32_t sb = __float_as_uint(ts);
  const int s_exp = static_cast<int>((sb >> 23) & 0xFFu);
  const uint32_t s_man = (sb & 0x7FFFFFu) | 0x800000u;  // ts = s_man * 2^(s_exp - 150)
  while (left > 0) {
    const uint32_t tb = __float_as_uint(t);
    const int t_exp = static_cast<int>((tb >> 23) & 0xFFu);
    const int shift = t_exp - s_exp;  // ulp(t) = 2^shift * ulp(ts)
    // non-positive, denormal or non-finite values, and steps at least as coarse as t: plain additions
    if (!(t > 0.0f) || t_exp == 0 || t_exp == 255 || s_exp == 0 || shift <= 0) {
      t = t + ts;
      --left;
      continue;
    }
    if (shift > 25) return t;  // the step is below half an ulp of t: t stalls
    uint32_t t_man = (tb & 0x7FFFFFu) | 0x800000u;
    const uint32_t whole = s_man >> shift, rem = s_man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    uint32_t inc;
    if (rem == half) {
      // exact ties round to even: the first addition depends on the parity of t; after it t is
      // even and every further addition in this binade adds `whole` ulps if that is even, else + 1
      const float tn = t + ts;
      --left;
      const bool same_binade = static_cast<int>((__float_as_uint(tn) >> 23) & 0xFFu) == t_exp;
      t = tn;
      if (left == 0 || !same_binade) continue;
      t_man = (__float_as_uint(t) & 0x7FFFFFu) | 0x800000u;
      inc = (whole & 1u) ? whole + 1u : whole;
    } else {
      inc = whole + (rem > half ? 1u : 0u);
    }
    if (inc == 0) return t;
    const uint32_t n = min(left, (0xFFFFFFu - t_man) / inc);  // additions that stay in the binade
    if (n > 0) {
      t = __uint_as_float((static_cast<uint32_t>(t_exp) << 23) | ((t_man + n * inc) & 0x7FFFFFu));
      left -= n;
    }
    if (left > 0) {  // the addition that reaches the next binade
      t = t + ts;
      --left;
    }
  }
  return t;
}

// number of steps axis b (first boundary t0b, step tsb) has taken when axis a is about to step
// with pre-step value ta; b_first = (b < a): equal values go to the lower axis
__device__ __forceinline__ unsigned steps_before(float t0b, float tsb, bool b_first, float ta,
                                                unsigned limit) {
  unsigned lo = 0, hi = limit + 1;  // smallest i with: b's (i+1)-th step does NOT come first
  while (lo < hi) {
    const unsigned mid = (lo + hi) >> 1;
    const float v = accumulate_steps(t0b, tsb, mid);
    if (v < ta || (v == ta && b_first)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct DirectState {
  unsigned n;   // global step index (number of advances made)
  int c[3];     // voxel index
  float t[3];   // t_to_next_boundary
  bool valid;   // false: the walk ends before this step
};
// state right after the j-th step (j >= 1) of axis a; c0 / t0 / ts / sign = the RayCaster's initial
// state, steps = ray_length_in_steps.  All three direction components must be non-zero and finite.
__device__ __forceinline__ DirectState raycast_state_after(const int c0[3], const int sign[3],
                                                           const float t0[3], const float ts[3],
                                                           unsigned steps, int a, unsigned j) {
  DirectState s;
  const float ta = accumulate_steps(t0[a], ts[a], j - 1);
  unsigned k[3];
#pragma unroll
  for (int b = 0; b < 3; ++b) k[b] = (b == a) ? j : steps_before(t0[b], ts[b], b < a, ta, steps);
  s.n = k[0] + k[1] + k[2];
  s.valid = s.n <= steps;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    s.c[b] = c0[b] + sign[b] * static_cast<int>(k[b]);
    s.t[b] = accumulate_steps(t0[b], ts[b], k[b]);
  }
  return s;
}

}  // namespace cg
