// selftest.cu — device-side self checks of arithmetic shortcuts (used by tests/test_gpu_math.py).
#include "cg_internal.cuh"

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

}  // namespace cg

using namespace cg;

extern "C" int32_t cg_debug_selftest(cg_context* ctx, int32_t which, uint64_t samples,
                                     uint64_t* mismatches) {
  if (!ctx || !mismatches) return CG_ERR_INVALID_ARG;
  CG_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* d = nullptr;
  CG_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
  CG_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
  const unsigned blocks = ctx->num_sms * 8, threads = 256;
  const uint64_t per_thread = samples / (static_cast<uint64_t>(blocks) * threads) + 1;
  if (which == 0)
    k_check_div<<<blocks, threads, 0, ctx->stream>>>(0x1234567ULL, per_thread, d);
  else if (which == 1)
    k_check_round<<<blocks, threads, 0, ctx->stream>>>(0x7654321ULL, per_thread, d);
  else {
    cudaFree(d);
    return CG_ERR_INVALID_ARG;
  }
  unsigned long long h = 0;
  CG_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CG_CUDA(cudaStreamSynchronize(ctx->stream));
  CG_CUDA(cudaGetLastError());
  cudaFree(d);
  *mismatches = h;
  return CG_OK;
}
