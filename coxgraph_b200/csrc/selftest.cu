// selftest.cu — device-side self checks of arithmetic shortcuts (used by tests/test_gpu_math.py).
#include "cg_internal.cuh"
#include "cg_raycast_direct.cuh"

namespace cg {

__device__ __forceinline__ uint32_t lcg(uint64_t& s) {
  s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  return static_cast<uint32_t>(s >> 32);
}

// div_with_rcp(num, den, RN(1/den)) must equal the IEEE quotient num / den bit for bit over the
// operand ranges of the bundle fold: den = running weight (small integers, 1/z^2 sums), num =
// weighted coordinate sums.
__global__ void k_check_div(uint64_t seed, uint64_t per_thread, unsigned long long* mismatches) {
  uint64_t s = seed + 0x9E3779B97F4A7C15ULL * (blockIdx.x * static_cast<uint64_t>(blockDim.x) +
                                               threadIdx.x + 1);
  unsigned long long bad = 0;
  for (uint64_t i = 0; i < per_thread; ++i) {
    const uint32_t a = lcg(s), b = lcg(s), c = lcg(s);
    float den;
    switch (c & 3) {
      case 0: den = static_cast<float>((b % 1000000u) + 1u); break;               // point counts
      case 1: den = static_cast<float>((b % 4096u) + 1u) + (b >> 20) * (1.0f / 4096.0f); break;
      case 2: den = __uint_as_float(0x3A000000u + (b % 0x0F000000u)); break;       // ~5e-4 .. 3e4
      default: den = __uint_as_float(0x3F800000u | (b & 0x007FFFFFu)); break;      // [1, 2) mantissas
    }
    float num = __uint_as_float(0x30000000u + (a % 0x1C000000u));                  // ~5e-10 .. 7e12
    if (c & 4) num = -num;
    if ((c & 0xF0) == 0) num = static_cast<float>(a % 100000u) * den;              // exact quotients
    const float r = 1.0f / den;
    const float q = div_with_rcp(num, den, r);
    const float ref = num / den;
    if (__float_as_uint(q) != __float_as_uint(ref)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// round_half_away_pos(v) == roundf(v) on [0, 256]
__global__ void k_check_round(uint64_t seed, uint64_t per_thread, unsigned long long* mismatches) {
  uint64_t s = seed + 0xD1B54A32D192ED03ULL * (blockIdx.x * static_cast<uint64_t>(blockDim.x) +
                                               threadIdx.x + 1);
  unsigned long long bad = 0;
  for (uint64_t i = 0; i < per_thread; ++i) {
    const uint32_t a = lcg(s);
    float v = (a & 1) ? __uint_as_float(a % 0x43800001u)                  // any float in [0, 256]
                      : static_cast<float>(a % 513u) * 0.5f + ((a >> 12) % 3 - 1) * 1e-5f * ((a >> 9) & 1);
    if (!(v >= 0.0f)) v = 0.0f;
    if (round_half_away_pos(v) != roundf(v)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// raycast_state_after (cg_raycast_direct.cuh) against the sequential RayCaster: random rays of the
// integrator's shapes (sensor origin, point up to ~6 m away, 5 cm or 2 cm voxels, carving); at every
// step that crosses a block face the directly computed state must equal the walked one — step
// index, voxel index and t_to_next_boundary bit for bit.  counters[0] = mismatches, [1] = block
// entries compared.
__global__ void k_check_direct_raycast(uint64_t seed, uint64_t per_thread,
                                       unsigned long long* counters) {
  uint64_t s = seed + 0xA24BAED4963EE407ULL * (blockIdx.x * static_cast<uint64_t>(blockDim.x) +
                                               threadIdx.x + 1);
  unsigned long long bad = 0, seen = 0;
  for (uint64_t i = 0; i < per_thread; ++i) {
    auto uni = [&](float lo, float hi) { return lo + (hi - lo) * (lcg(s) >> 8) * (1.0f / 16777216.0f); };
    const V3 origin = V3{uni(-6.0f, 6.0f), uni(-4.0f, 4.0f), uni(0.0f, 3.0f)};
    const V3 dir = V3{uni(-1.0f, 1.0f), uni(-1.0f, 1.0f), uni(-1.0f, 1.0f)};
    const float len = uni(0.2f, 6.0f);
    const V3 point = origin + normalized3(dir) * len;
    const bool fine = (lcg(s) & 1u) != 0;
    RayCaster rc;
    rc.init(origin, point, (lcg(s) & 7u) == 0, true, 5.0f, fine ? 50.0f : 20.0f, fine ? 0.06f : 0.16f);
    if (!rc.valid || rc.steps == 0 || rc.sx == 0 || rc.sy == 0 || rc.sz == 0) continue;
    const int c0[3] = {rc.cx, rc.cy, rc.cz}, sign[3] = {rc.sx, rc.sy, rc.sz};
    const float t0[3] = {rc.tnx, rc.tny, rc.tnz}, ts[3] = {rc.tsx, rc.tsy, rc.tsz};
    const unsigned steps = rc.steps;
    unsigned k[3] = {0, 0, 0};
    for (unsigned n = 1; n <= steps; ++n) {
      const int px = rc.cx, py = rc.cy, pz = rc.cz;
      rc.step();
      const int a = rc.cx != px ? 0 : (rc.cy != py ? 1 : 2);
      ++k[a];
      const int before = a == 0 ? px : (a == 1 ? py : pz);
      const int after = a == 0 ? rc.cx : (a == 1 ? rc.cy : rc.cz);
      if ((before >> 4) == (after >> 4)) continue;  // stays inside the block
      const DirectState d = raycast_state_after(c0, sign, t0, ts, steps, a, k[a]);
      ++seen;
      const bool ok = d.valid && d.n == n && d.c[0] == rc.cx && d.c[1] == rc.cy && d.c[2] == rc.cz &&
                      __float_as_uint(d.t[0]) == __float_as_uint(rc.tnx) &&
                      __float_as_uint(d.t[1]) == __float_as_uint(rc.tny) &&
                      __float_as_uint(d.t[2]) == __float_as_uint(rc.tnz);
      if (!ok) ++bad;
    }
  }
  if (bad) atomicAdd(&counters[0], bad);
  if (seen) atomicAdd(&counters[1], seen);
}

}  // namespace cg

using namespace cg;

extern "C" int32_t cg_debug_selftest(cg_context* ctx, int32_t which, uint64_t samples,
                                     uint64_t* mismatches) {
  if (!ctx || !mismatches) return CG_ERR_INVALID_ARG;
  CG_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* d = nullptr;
  CG_CUDA(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
  CG_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
  const unsigned blocks = ctx->num_sms * 8, threads = 256;
  const uint64_t per_thread = samples / (static_cast<uint64_t>(blocks) * threads) + 1;
  if (which == 0)
    k_check_div<<<blocks, threads, 0, ctx->stream>>>(0x1234567ULL, per_thread, d);
  else if (which == 1)
    k_check_round<<<blocks, threads, 0, ctx->stream>>>(0x7654321ULL, per_thread, d);
  else if (which == 2 || which == 3)  // samples = rays; 2 -> mismatches, 3 -> block entries compared
    k_check_direct_raycast<<<blocks, threads, 0, ctx->stream>>>(0x2468ACEULL, per_thread, d);
  else {
    cudaFree(d);
    return CG_ERR_INVALID_ARG;
  }
  unsigned long long h = 0;
  CG_CUDA(cudaMemcpyAsync(&h, d + (which == 3 ? 1 : 0), sizeof(h), cudaMemcpyDeviceToHost,
                          ctx->stream));
  CG_CUDA(cudaStreamSynchronize(ctx->stream));
  CG_CUDA(cudaGetLastError());
  cudaFree(d);
  *mismatches = h;
  return CG_OK;
}
