// integrate.cu — TSDF integration of point clouds into the block-hashed layer.
//
// Replaces voxblox::TsdfIntegratorBase::integratePointCloud (Simple / Merged semantics, R1-R6 of
// SURVEY.md §8a); reference call site coxgraph/include/coxgraph/map_comm/tsdf_recover.h:75.
//
// Pipeline per submitted job (one frame, or a batch of frames into the same layer):
//   per frame   k_point_keys      validity, T_G_C * p, bundle key (voxel of p_G), in index-getter order
//               radix sort        stable sort-by-key  -> bundles, points inside a bundle in visit order
//               k_mark_heads/scan bundle ids
//               k_fold_bundles    sequential weighted mean / colour blend per bundle -> one ray
//   per job     scan              pair offsets from the closed-form ray length (steps + 1)
//               k_ray_walk        3-D DDA per ray; warp-independent hash insert of every block the
//                                 ray visits; emits (hash entry, voxel) keys in ray order
//               radix sort        stable sort-by-key -> per-voxel update lists in canonical order
//               k_voxel_update    one owner per voxel replays its updates sequentially
// The only value-dependent ordering is per voxel: (frame, non-clearing before clearing, bundle
// key ascending) — exactly the oracle's canonical order, so results do not depend on scheduling.
#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "cg_internal.cuh"

namespace cg {

struct Ray {          // 24 B
  float px, py, pz;   // merged point, global frame
  float weight;       // merged weight
  uint32_t color;     // merged colour
  uint32_t frame_clr; // frame index in job | clearing << 31
};

__device__ __forceinline__ int order_index(int k, int n, int mode) {
  // voxblox MixedThreadSafeIndex: groups of 1024 visited round-robin
  if (mode != 0) return k;
  const int groups = n / 1024;
  if (groups * 1024 <= k) return k;
  return (k % groups) * 1024 + (k / groups);
}

__device__ __forceinline__ V3 load_point(const float* pts, int i) {
  return V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
}

// ------------------------------------------------------------------ MERGED front half
__global__ void k_point_keys(IntegratorParams P, Xform T, const float* __restrict__ pts, int n,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                             int32_t* err) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = order_index(k, n, P.order_mode);
  const V3 pc = load_point(pts, i);
  bool clearing = false;
  uint64_t key = kInvalidPointKey;
  if (point_valid(P, pc, &clearing)) {
    const V3 pg = apply(T, pc);
    const float lim = 524000.0f * P.voxel_size;
    if (fabsf(pg.x) < lim && fabsf(pg.y) < lim && fabsf(pg.z) < lim) {
      const int vx = grid_index(pg.x, P.voxel_size_inv);
      const int vy = grid_index(pg.y, P.voxel_size_inv);
      const int vz = grid_index(pg.z, P.voxel_size_inv);
      key = pack_voxel_key(vx, vy, vz) | (static_cast<uint64_t>(clearing) << kClearingBit);
    } else {
      atomicOr(err, kErrOutOfRange);
    }
  }
  keys[k] = key;
  vals[k] = static_cast<uint32_t>(i);
}

__global__ void k_mark_heads(const uint64_t* __restrict__ keys, int n, uint32_t* flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t key = keys[k];
  flags[k] = (key != kInvalidPointKey && (k == 0 || keys[k - 1] != key)) ? 1u : 0u;
}

// one thread per bundle head: MergedTsdfIntegrator::integrateVoxel, first half
__global__ void k_fold_bundles(IntegratorParams P, Xform T, int frame,
                               const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                               const uint32_t* __restrict__ flags, const uint32_t* __restrict__ scan,
                               int n, const float* __restrict__ pts,
                               const uint32_t* __restrict__ cols, uint32_t* frame_base, Ray* rays,
                               uint32_t* ray_count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t base = frame_base[frame];
  if (k == n - 1) frame_base[frame + 1] = base + scan[k] + flags[k];
  if (!flags[k]) return;
  const uint64_t key = keys[k];
  const bool clearing = (key >> kClearingBit) & 1;
  uint32_t merged_color = kDefaultColor;
  V3 merged = V3{0.0f, 0.0f, 0.0f};
  float merged_weight = 0.0f;
  for (int j = k; j < n && keys[j] == key; ++j) {
    const int i = static_cast<int>(vals[j]);
    const V3 pc = load_point(pts, i);
    const float w = voxel_weight(P, pc.z);
    if (w < kEps) continue;
    merged = (merged * merged_weight + pc * w) / (merged_weight + w);
    merged_color = blend_colors(merged_color, merged_weight, cols[i], w);
    merged_weight += w;
    if (clearing) break;  // only the first point of a clearing bundle is used
  }
  const V3 pg = apply(T, merged);
  RayCaster rc;
  rc.init(T.t, pg, clearing, P.carving != 0, P.max_ray, P.voxel_size_inv, P.trunc);
  const uint32_t r = base + scan[k];
  Ray ray;
  ray.px = pg.x;
  ray.py = pg.y;
  ray.pz = pg.z;
  ray.weight = merged_weight;
  ray.color = merged_color;
  ray.frame_clr = static_cast<uint32_t>(frame) | (clearing ? 0x80000000u : 0u);
  rays[r] = ray;
  ray_count[r] = rc.valid ? rc.steps + 1u : 0u;
}

// ------------------------------------------------------------------ SIMPLE front half
__global__ void k_simple_flags(IntegratorParams P, const float* __restrict__ pts, int n,
                               uint32_t* flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  bool clearing;
  flags[k] = point_valid(P, load_point(pts, order_index(k, n, P.order_mode)), &clearing) ? 1u : 0u;
}

__global__ void k_simple_rays(IntegratorParams P, Xform T, int frame,
                              const uint32_t* __restrict__ flags, const uint32_t* __restrict__ scan,
                              int n, const float* __restrict__ pts,
                              const uint32_t* __restrict__ cols, uint32_t* frame_base, Ray* rays,
                              uint32_t* ray_count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t base = frame_base[frame];
  if (k == n - 1) frame_base[frame + 1] = base + scan[k] + flags[k];
  if (!flags[k]) return;
  const int i = order_index(k, n, P.order_mode);
  const V3 pc = load_point(pts, i);
  bool clearing = false;
  point_valid(P, pc, &clearing);
  const V3 pg = apply(T, pc);
  RayCaster rc;
  rc.init(T.t, pg, clearing, P.carving != 0, P.max_ray, P.voxel_size_inv, P.trunc);
  const uint32_t r = base + scan[k];
  Ray ray;
  ray.px = pg.x;
  ray.py = pg.y;
  ray.pz = pg.z;
  ray.weight = voxel_weight(P, pc.z);
  ray.color = cols[i];
  ray.frame_clr = static_cast<uint32_t>(frame) | (clearing ? 0x80000000u : 0u);
  rays[r] = ray;
  ray_count[r] = rc.valid ? rc.steps + 1u : 0u;
}

__global__ void k_copy_base(uint32_t* frame_base, int frame) {
  frame_base[frame + 1] = frame_base[frame];
}

__global__ void k_totals(const uint32_t* frame_base, int frames, const uint32_t* ray_count,
                         const uint32_t* ray_offset, size_t upper, CallCounters* c) {
  c->rays = frame_base[frames];
  c->pairs = upper ? static_cast<unsigned long long>(ray_offset[upper - 1]) + ray_count[upper - 1]
                   : 0ull;
  c->touched = 0;
}

// ------------------------------------------------------------------ back half
// Walk every ray (Amanatides-Woo DDA exactly as voxblox::RayCaster), allocate every block it
// visits (R4: allocation on first visit), emit one (hash entry << 12 | voxel) key per visit.
template <class K>
__global__ void k_ray_walk(IntegratorParams P, const float* __restrict__ poses,
                           const Ray* __restrict__ rays, const uint32_t* __restrict__ ray_offset,
                           uint32_t num_rays, LayerView L, K* __restrict__ pkeys,
                           uint32_t* __restrict__ pvals) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= num_rays) return;
  const Ray ray = rays[r];
  const uint32_t frame = ray.frame_clr & 0x7FFFFFFFu;
  const bool clearing = (ray.frame_clr >> 31) != 0;
  const float* T = poses + 7 * frame;
  const V3 origin = V3{T[4], T[5], T[6]};
  RayCaster rc;
  rc.init(origin, V3{ray.px, ray.py, ray.pz}, clearing, P.carving != 0, P.max_ray,
          P.voxel_size_inv, P.trunc);
  if (!rc.valid) {
    if (!rc.in_range) atomicOr(L.err, kErrOutOfRange);
    return;
  }
  size_t out = ray_offset[r];
  int lbx = 0x7FFFFFFF, lby = 0, lbz = 0;
  K entry_bits = 0;
  for (unsigned s = 0; s <= rc.steps; ++s) {
    const int bx = rc.cx >> 4, by = rc.cy >> 4, bz = rc.cz >> 4;
    if (bx != lbx || by != lby || bz != lbz) {
      lbx = bx;
      lby = by;
      lbz = bz;
      entry_bits = static_cast<K>(L.insert_entry(pack_block_key(bx, by, bz))) << 12;
    }
    const int lin = (rc.cx & 15) + 16 * ((rc.cy & 15) + 16 * (rc.cz & 15));
    pkeys[out] = entry_bits | static_cast<K>(lin);
    pvals[out] = r;
    ++out;
    rc.step();
  }
}

// One owner thread per voxel: replays the voxel's update list in order (R5).
template <class K>
__global__ void k_voxel_update(IntegratorParams P, const float* __restrict__ poses,
                               const Ray* __restrict__ rays, const K* __restrict__ pkeys,
                               const uint32_t* __restrict__ pvals, size_t num_pairs, LayerView L,
                               CallCounters* counters) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= num_pairs) return;
  const K key = pkeys[i];
  const K prev = i ? pkeys[i - 1] : ~key;
  if (i && prev == key) return;  // not the head of its voxel segment
  const uint32_t entry = static_cast<uint32_t>(key >> 12);
  const int lin = static_cast<int>(key & 4095);
  const int slot = L.hash_vals[entry];
  if (slot < 0) return;  // pool exhausted; error already flagged
  if (i == 0 || static_cast<uint32_t>(prev >> 12) != entry) {
    atomicAdd(&counters->touched, 1ull);
    L.updated[slot] = 1;
  }
  int bx, by, bz;
  unpack_block_key(L.hash_keys[entry], bx, by, bz);
  const int gx = bx * 16 + (lin & 15), gy = by * 16 + ((lin >> 4) & 15), gz = bz * 16 + (lin >> 8);
  const V3 center = V3{center_coord(gx, P.voxel_size), center_coord(gy, P.voxel_size),
                       center_coord(gz, P.voxel_size)};
  float* dp = L.dist_plane(slot) + lin;
  float* wp = L.weight_plane(slot) + lin;
  uint32_t* cp = L.color_plane(slot) + lin;
  VoxelState v{*dp, *wp, *cp};
  for (size_t j = i; j < num_pairs && pkeys[j] == key; ++j) {
    const Ray ray = rays[pvals[j]];
    const float* T = poses + 7 * (ray.frame_clr & 0x7FFFFFFFu);
    update_tsdf_voxel(P, V3{T[4], T[5], T[6]}, V3{ray.px, ray.py, ray.pz}, center, ray.color,
                      ray.weight, v);
  }
  *dp = v.d;
  *wp = v.w;
  *cp = v.c;
}

// ------------------------------------------------------------------ host orchestration
static IntegratorParams make_params(const cg_layer* L, const cg_integrator_config* c, int freespace) {
  IntegratorParams P;
  P.trunc = c->default_truncation_distance;
  P.max_weight = c->max_weight;
  P.min_ray = c->min_ray_length_m;
  P.max_ray = c->max_ray_length_m;
  P.voxel_size = L->v.voxel_size;
  P.voxel_size_inv = L->v.voxel_size_inv;
  P.sparsity_factor = c->sparsity_compensation_factor;
  P.carving = c->voxel_carving_enabled;
  P.const_weight = c->use_const_weight;
  P.allow_clear = c->allow_clear;
  P.weight_dropoff = c->use_weight_dropoff;
  P.use_sparsity = c->use_sparsity_compensation_factor;
  P.order_mode = c->integration_order_mode;
  P.freespace = freespace;
  return P;
}

static size_t env_size(const char* name, size_t dflt) {
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  return static_cast<size_t>(strtoull(s, nullptr, 10));
}

template <class K>
static int32_t run_back_half(cg_context* ctx, cg_layer* L, const IntegratorParams& P,
                             uint32_t num_rays, size_t num_pairs, int key_bits) {
  cudaStream_t s = ctx->stream;
  CG_CUDA(ctx->pkey_a.reserve(num_pairs * sizeof(K)));
  CG_CUDA(ctx->pkey_b.reserve(num_pairs * sizeof(K)));
  CG_CUDA(ctx->pval_a.reserve(num_pairs * sizeof(uint32_t)));
  CG_CUDA(ctx->pval_b.reserve(num_pairs * sizeof(uint32_t)));
  k_ray_walk<K><<<grid_for(num_rays, 128), 128, 0, s>>>(
      P, ctx->poses.as<float>(), ctx->rays.as<Ray>(), ctx->ray_offset.as<uint32_t>(), num_rays,
      L->v, ctx->pkey_a.as<K>(), ctx->pval_a.as<uint32_t>());
  cub::DoubleBuffer<K> dk(ctx->pkey_a.as<K>(), ctx->pkey_b.as<K>());
  cub::DoubleBuffer<uint32_t> dv(ctx->pval_a.as<uint32_t>(), ctx->pval_b.as<uint32_t>());
  size_t tmp = 0;
  CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, num_pairs, 0, key_bits, s));
  CG_CUDA(ctx->cub_tmp.reserve(tmp));
  CG_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, dk, dv, num_pairs, 0, key_bits, s));
  k_voxel_update<K><<<grid_for(num_pairs, 128), 128, 0, s>>>(
      P, ctx->poses.as<float>(), ctx->rays.as<Ray>(), dk.Current(), dv.Current(), num_pairs, L->v,
      ctx->d_counters);
  CG_CUDA(cudaGetLastError());
  return CG_OK;
}

// frames [f0, f1) of the job as one group; splits itself when the pair list would not fit
static int32_t integrate_group(cg_layer* L, const cg_integrator_config* cfg,
                               const IntegratorParams& P, const float* h_poses,
                               const float* d_points, const uint8_t* d_colors,
                               const uint64_t* offs, size_t f0, size_t f1,
                               cg_integrate_stats* stats) {
  cg_context* ctx = L->ctx;
  cudaStream_t s = ctx->stream;
  const size_t F = f1 - f0;
  const size_t upper = offs[f1] - offs[f0];  // upper bound on rays
  if (upper == 0) return CG_OK;
  size_t max_n = 0;
  for (size_t f = f0; f < f1; ++f) max_n = std::max<size_t>(max_n, offs[f + 1] - offs[f]);
  if (max_n > 0x7FFFFFFFull || upper > 0xFFFFFFF0ull) {
    set_error("too many points in one frame / group");
    return CG_ERR_INVALID_ARG;
  }
  CG_CUDA(ctx->rays.reserve(upper * sizeof(Ray)));
  CG_CUDA(ctx->ray_count.reserve(upper * sizeof(uint32_t)));
  CG_CUDA(ctx->ray_offset.reserve(upper * sizeof(uint32_t)));
  CG_CUDA(ctx->poses.reserve(F * 7 * sizeof(float)));
  CG_CUDA(ctx->frame_base.reserve((F + 1) * sizeof(uint32_t)));
  CG_CUDA(ctx->key_a.reserve(max_n * sizeof(uint64_t)));
  CG_CUDA(ctx->key_b.reserve(max_n * sizeof(uint64_t)));
  CG_CUDA(ctx->val_a.reserve(max_n * sizeof(uint32_t)));
  CG_CUDA(ctx->val_b.reserve(max_n * sizeof(uint32_t)));
  CG_CUDA(ctx->flags.reserve(max_n * sizeof(uint32_t)));
  CG_CUDA(ctx->scan.reserve(max_n * sizeof(uint32_t)));
  size_t tmp_sort = 0, tmp_scan = 0, tmp_scan2 = 0;
  {
    cub::DoubleBuffer<uint64_t> dk(ctx->key_a.as<uint64_t>(), ctx->key_b.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> dv(ctx->val_a.as<uint32_t>(), ctx->val_b.as<uint32_t>());
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, static_cast<int>(max_n), 0, 61, s);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, ctx->flags.as<uint32_t>(),
                                  ctx->scan.as<uint32_t>(), static_cast<int>(max_n), s);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan2, ctx->ray_count.as<uint32_t>(),
                                  ctx->ray_offset.as<uint32_t>(), upper, s);
  }
  CG_CUDA(ctx->cub_tmp.reserve(std::max(tmp_sort, std::max(tmp_scan, tmp_scan2))));
  CG_CUDA(cudaMemcpyAsync(ctx->poses.p, h_poses + 7 * f0, F * 7 * sizeof(float),
                          cudaMemcpyHostToDevice, s));
  CG_CUDA(cudaMemsetAsync(ctx->frame_base.p, 0, sizeof(uint32_t), s));
  CG_CUDA(cudaMemsetAsync(ctx->ray_count.p, 0, upper * sizeof(uint32_t), s));

  for (size_t f = f0; f < f1; ++f) {
    const int n = static_cast<int>(offs[f + 1] - offs[f]);
    const int fi = static_cast<int>(f - f0);
    if (n == 0) {
      k_copy_base<<<1, 1, 0, s>>>(ctx->frame_base.as<uint32_t>(), fi);
      continue;
    }
    const float* pts = d_points + 3 * offs[f];
    const uint32_t* cols = reinterpret_cast<const uint32_t*>(d_colors) + offs[f];
    const Xform T = make_xform(h_poses + 7 * f);
    const unsigned g = grid_for(n, 256);
    if (cfg->method == CG_METHOD_MERGED) {
      k_point_keys<<<g, 256, 0, s>>>(P, T, pts, n, ctx->key_a.as<uint64_t>(),
                                     ctx->val_a.as<uint32_t>(), L->v.err);
      cub::DoubleBuffer<uint64_t> dk(ctx->key_a.as<uint64_t>(), ctx->key_b.as<uint64_t>());
      cub::DoubleBuffer<uint32_t> dv(ctx->val_a.as<uint32_t>(), ctx->val_b.as<uint32_t>());
      CG_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_sort, dk, dv, n, 0, 61, s));
      k_mark_heads<<<g, 256, 0, s>>>(dk.Current(), n, ctx->flags.as<uint32_t>());
      CG_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_scan, ctx->flags.as<uint32_t>(),
                                            ctx->scan.as<uint32_t>(), n, s));
      k_fold_bundles<<<g, 256, 0, s>>>(P, T, fi, dk.Current(), dv.Current(),
                                       ctx->flags.as<uint32_t>(), ctx->scan.as<uint32_t>(), n, pts,
                                       cols, ctx->frame_base.as<uint32_t>(), ctx->rays.as<Ray>(),
                                       ctx->ray_count.as<uint32_t>());
    } else {
      k_simple_flags<<<g, 256, 0, s>>>(P, pts, n, ctx->flags.as<uint32_t>());
      CG_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_scan, ctx->flags.as<uint32_t>(),
                                            ctx->scan.as<uint32_t>(), n, s));
      k_simple_rays<<<g, 256, 0, s>>>(P, T, fi, ctx->flags.as<uint32_t>(), ctx->scan.as<uint32_t>(),
                                      n, pts, cols, ctx->frame_base.as<uint32_t>(),
                                      ctx->rays.as<Ray>(), ctx->ray_count.as<uint32_t>());
    }
  }
  CG_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_scan2, ctx->ray_count.as<uint32_t>(),
                                        ctx->ray_offset.as<uint32_t>(), upper, s));
  k_totals<<<1, 1, 0, s>>>(ctx->frame_base.as<uint32_t>(), static_cast<int>(F),
                           ctx->ray_count.as<uint32_t>(), ctx->ray_offset.as<uint32_t>(), upper,
                           ctx->d_counters);
  CG_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(CallCounters),
                          cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  CG_CUDA(cudaGetLastError());
  const uint32_t num_rays = static_cast<uint32_t>(ctx->h_counters->rays);
  const size_t num_pairs = ctx->h_counters->pairs;
  const size_t max_pairs = env_size("CG_MAX_PAIRS", size_t(768) << 20);
  if (num_pairs > max_pairs || num_pairs >= 0xFFFFFFF0ull) {
    if (F > 1) {  // the front half never touches the layer: safe to redo in two halves
      const size_t mid = f0 + F / 2;
      int32_t rc = integrate_group(L, cfg, P, h_poses, d_points, d_colors, offs, f0, mid, stats);
      if (rc) return rc;
      return integrate_group(L, cfg, P, h_poses, d_points, d_colors, offs, mid, f1, stats);
    }
    if (num_pairs >= 0xFFFFFFF0ull) {
      set_error("a single frame produces %zu voxel updates; split the point cloud", num_pairs);
      return CG_ERR_INVALID_ARG;
    }
  }
  if (num_rays == 0 || num_pairs == 0) return CG_OK;
  int hash_bits = 0;
  while ((size_t(1) << hash_bits) < L->hash_cap) ++hash_bits;
  const int key_bits = hash_bits + 12;
  int32_t rc = key_bits <= 32 ? run_back_half<uint32_t>(ctx, L, P, num_rays, num_pairs, key_bits)
                              : run_back_half<uint64_t>(ctx, L, P, num_rays, num_pairs, key_bits);
  if (rc) return rc;
  if (stats) {
    stats->rays += num_rays;
    stats->voxel_updates += num_pairs;
    // blocks_touched is accumulated on the device (counters->touched) per group
    CG_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(CallCounters),
                            cudaMemcpyDeviceToHost, s));
    CG_CUDA(cudaStreamSynchronize(s));
    stats->blocks_touched += ctx->h_counters->touched;
  }
  return CG_OK;
}

static int32_t integrate_job(cg_layer* L, const cg_integrator_config* cfg, size_t F,
                             const float* h_poses, const float* d_points, const uint8_t* d_colors,
                             const uint64_t* offs, int freespace, cg_integrate_stats* stats) {
  if (cfg->method == CG_METHOD_FAST) {
    set_error("method FAST is order- and wall-clock-dependent in the reference and has no "
              "deterministic device form; use MERGED or SIMPLE");
    return CG_ERR_UNSUPPORTED;
  }
  if (cfg->method != CG_METHOD_MERGED && cfg->method != CG_METHOD_SIMPLE) return CG_ERR_INVALID_ARG;
  if (cfg->enable_anti_grazing) {
    set_error("enable_anti_grazing is not supported yet");
    return CG_ERR_UNSUPPORTED;
  }
  if (!(cfg->default_truncation_distance > 0.0f) || !(cfg->max_ray_length_m > 0.0f)) {
    set_error("invalid integrator config");
    return CG_ERR_INVALID_ARG;
  }
  CG_CUDA(cudaSetDevice(L->ctx->device));
  const IntegratorParams P = make_params(L, cfg, freespace);
  const int64_t blocks_before = L->num_blocks;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->points_in = offs[F] - offs[0];
  }
  const size_t max_group_points = env_size("CG_MAX_GROUP_POINTS", size_t(48) << 20);
  int32_t rc = CG_OK;
  size_t f0 = 0;
  while (f0 < F && rc == CG_OK) {
    size_t f1 = f0 + 1;
    while (f1 < F && offs[f1 + 1] - offs[f0] <= max_group_points) ++f1;
    rc = integrate_group(L, cfg, P, h_poses, d_points, d_colors, offs, f0, f1, stats);
    f0 = f1;
  }
  const int32_t rc2 = finish_call(L, nullptr);
  if (stats) stats->blocks_allocated = L->num_blocks - blocks_before;
  return rc ? rc : rc2;
}

static int32_t stage_inputs(cg_context* ctx, const float* pts, const uint8_t* cols, size_t n) {
  CG_CUDA(cudaSetDevice(ctx->device));
  CG_CUDA(ctx->points.reserve(n * 3 * sizeof(float)));
  CG_CUDA(ctx->colors.reserve(n * 4));
  CG_CUDA(cudaMemcpyAsync(ctx->points.p, pts, n * 3 * sizeof(float), cudaMemcpyHostToDevice,
                          ctx->stream));
  CG_CUDA(cudaMemcpyAsync(ctx->colors.p, cols, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  return CG_OK;
}

}  // namespace cg

using namespace cg;

extern "C" {

void cg_integrator_config_default(cg_integrator_config* c) {
  if (!c) return;
  c->default_truncation_distance = 0.1f;
  c->max_weight = 10000.0f;
  c->voxel_carving_enabled = 1;
  c->min_ray_length_m = 0.1f;
  c->max_ray_length_m = 5.0f;
  c->use_const_weight = 0;
  c->allow_clear = 1;
  c->use_weight_dropoff = 1;
  c->use_sparsity_compensation_factor = 0;
  c->sparsity_compensation_factor = 1.0f;
  c->enable_anti_grazing = 0;
  c->method = CG_METHOD_MERGED;
  c->integration_order_mode = CG_ORDER_MIXED;
  c->start_voxel_subsampling_factor = 2.0f;
  c->max_consecutive_ray_collisions = 2;
}

int32_t cg_integrate_batch_device(cg_layer* L, const cg_integrator_config* cfg, size_t F,
                                  const float* poses, const float* d_pts, const uint8_t* d_cols,
                                  const uint64_t* offs, int32_t freespace,
                                  cg_integrate_stats* stats) {
  if (!L || !cfg || !poses || !offs || (offs[F] > offs[0] && (!d_pts || !d_cols))) {
    set_error("cg_integrate: null argument");
    return CG_ERR_INVALID_ARG;
  }
  for (size_t f = 0; f < F; ++f)
    if (offs[f + 1] < offs[f]) {
      set_error("frame_offsets must be non-decreasing");
      return CG_ERR_INVALID_ARG;
    }
  return integrate_job(L, cfg, F, poses, d_pts, d_cols, offs, freespace, stats);
}

int32_t cg_integrate_batch(cg_layer* L, const cg_integrator_config* cfg, size_t F,
                           const float* poses, const float* pts, const uint8_t* cols,
                           const uint64_t* offs, int32_t freespace, cg_integrate_stats* stats) {
  if (!L || !cfg || !poses || !offs) {
    set_error("cg_integrate: null argument");
    return CG_ERR_INVALID_ARG;
  }
  const size_t first = offs[0], total = offs[F] - offs[0];
  if (total && (!pts || !cols)) {
    set_error("cg_integrate: null argument");
    return CG_ERR_INVALID_ARG;
  }
  int32_t rc = stage_inputs(L->ctx, pts + 3 * first, cols + 4 * first, total);
  if (rc) return rc;
  std::vector<uint64_t> rel(F + 1);
  for (size_t f = 0; f <= F; ++f) rel[f] = offs[f] - first;
  return cg_integrate_batch_device(L, cfg, F, poses, L->ctx->points.as<float>(),
                                   L->ctx->colors.as<uint8_t>(), rel.data(), freespace, stats);
}

int32_t cg_integrate_pointcloud_device(cg_layer* L, const cg_integrator_config* cfg,
                                       const float T[7], const float* d_pts, const uint8_t* d_cols,
                                       size_t n, int32_t freespace, cg_integrate_stats* stats) {
  const uint64_t offs[2] = {0, n};
  return cg_integrate_batch_device(L, cfg, 1, T, d_pts, d_cols, offs, freespace, stats);
}

int32_t cg_integrate_pointcloud(cg_layer* L, const cg_integrator_config* cfg, const float T[7],
                                const float* pts, const uint8_t* cols, size_t n, int32_t freespace,
                                cg_integrate_stats* stats) {
  const uint64_t offs[2] = {0, n};
  return cg_integrate_batch(L, cfg, 1, T, pts, cols, offs, freespace, stats);
}

}  // extern "C"
