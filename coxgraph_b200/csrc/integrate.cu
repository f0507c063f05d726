// integrate.cu — TSDF integration of point clouds into the block-hashed layer.
//
// Replaces voxblox::TsdfIntegratorBase::integratePointCloud (Simple / Merged semantics, R1-R6 of
// SURVEY.md §8a); reference call site coxgraph/include/coxgraph/map_comm/tsdf_recover.h:75.
//
// One job = one frame, or a batch of frames fused into the same layer (the loop of
// tsdf_recover.h:71-86).  All frames of a job go through every stage together:
//   k_point_keys    validity, T_G_C * p, bundle key = (frame | clearing | voxel of p_G relative to
//                   the sensor voxel), written in index-getter ("mixed") order
//   radix sort      stable sort-by-key -> bundles in canonical order, points of a bundle in the
//                   order the reference visits them
//   select          bundle heads
//   k_fold_bundles  one warp per bundle: cooperative gather, then the reference's *sequential*
//                   weighted mean / colour blend (bit-exact: the merged point decides which voxels
//                   and blocks the ray visits) -> one ray per bundle
//   scan            pair offsets from the closed-form ray length (steps + 1)
//   k_ray_walk      3-D DDA per ray (voxblox::RayCaster); inserts every visited block into the
//                   GPU hash (allocation on first visit, R4); emits (hash entry, voxel) keys
//   radix sort      stable sort-by-key -> per-voxel update lists in canonical order
//   select          voxel segment heads
//   k_voxel_update  one warp per voxel segment; 32 updates at a time: weights by prefix sum, the
//                   clamped weighted average as an ordered composition of clamped affine maps
//                   x -> clamp(a x + b) (associative, so it reduces in log steps), colours
//                   replayed sequentially for the few updates inside the truncation band.
// Per-voxel order is (frame, non-clearing before clearing, bundle key ascending) — the oracle's
// canonical order — independent of scheduling.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "cg_internal.cuh"

namespace cg {

struct Ray {           // 24 B
  float px, py, pz;    // merged point, global frame
  float weight;        // merged weight
  uint32_t color;      // merged colour
  uint32_t frame_clr;  // frame index in job | clearing << 31 ; 0xFFFFFFFF = no ray
};
constexpr uint32_t kNoRay = 0xFFFFFFFFu;

// bundle key layout: [frame | clearing(1) | z(13) y(13) x(13)], voxel index relative to the
// voxel holding the sensor origin
constexpr int kRelBits = 13;
constexpr int kRelOffset = 1 << (kRelBits - 1);
constexpr int kBundleClearBit = 3 * kRelBits;
constexpr int kBundleFrameShift = kBundleClearBit + 1;

__device__ __forceinline__ int order_index(int k, int n, int mode) {
  // voxblox MixedThreadSafeIndex: groups of 1024 visited round-robin
  if (mode != 0) return k;
  const int groups = n / 1024;
  if (groups * 1024 <= k) return k;
  return (k % groups) * 1024 + (k / groups);
}

__device__ __forceinline__ V3 load_point(const float* pts, size_t i) {
  return V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
}

// frame of global slot g (offs has F+1 entries)
__device__ __forceinline__ int frame_of(const uint64_t* __restrict__ offs, int F, uint64_t g) {
  int lo = 0, hi = F;  // offs[lo] <= g < offs[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offs[mid] <= g) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------ front half
// slot g enumerates (frame, visit rank k); vals = global point index
__global__ void k_point_keys(IntegratorParams P, const float* __restrict__ poses,
                             const uint64_t* __restrict__ offs, int F,
                             const float* __restrict__ pts, uint64_t total,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                             int32_t* err) {
  const uint64_t g = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (g >= total) return;
  const int f = frame_of(offs, F, g);
  const uint64_t base = offs[f];
  const int n = static_cast<int>(offs[f + 1] - base);
  const int i = order_index(static_cast<int>(g - base), n, P.order_mode);
  const V3 pc = load_point(pts, base + i);
  bool clearing = false;
  uint64_t key = kInvalidPointKey;
  if (point_valid(P, pc, &clearing)) {
    const Xform T = make_xform(poses + 7 * f);
    const V3 pg = apply(T, pc);
    const float lim = 524000.0f * P.voxel_size;
    if (fabsf(pg.x) < lim && fabsf(pg.y) < lim && fabsf(pg.z) < lim && fabsf(T.t.x) < lim &&
        fabsf(T.t.y) < lim && fabsf(T.t.z) < lim) {
      const int rx = grid_index(pg.x, P.voxel_size_inv) - grid_index(T.t.x, P.voxel_size_inv);
      const int ry = grid_index(pg.y, P.voxel_size_inv) - grid_index(T.t.y, P.voxel_size_inv);
      const int rz = grid_index(pg.z, P.voxel_size_inv) - grid_index(T.t.z, P.voxel_size_inv);
      if (abs(rx) < kRelOffset && abs(ry) < kRelOffset && abs(rz) < kRelOffset) {
        key = (static_cast<uint64_t>(f) << kBundleFrameShift) |
              (static_cast<uint64_t>(clearing) << kBundleClearBit) |
              (static_cast<uint64_t>(rz + kRelOffset) << (2 * kRelBits)) |
              (static_cast<uint64_t>(ry + kRelOffset) << kRelBits) |
              static_cast<uint64_t>(rx + kRelOffset);
      } else {
        atomicOr(err, kErrOutOfRange);
      }
    } else {
      atomicOr(err, kErrOutOfRange);
    }
  }
  keys[g] = key;
  vals[g] = static_cast<uint32_t>(base + i);
}

struct BundleHead {
  const uint64_t* keys;
  __device__ __forceinline__ bool operator()(uint32_t i) const {
    // the first invalid key is a head too: it terminates the last real bundle
    return i == 0 || keys[i] != keys[i - 1];
  }
};

// Gather the sorted points next to each other: (x, y, z, rgba) per sorted slot, so that the
// sequential fold streams contiguous memory.
__global__ void k_gather_sorted(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                uint32_t total, const float* __restrict__ pts,
                                const uint32_t* __restrict__ cols, float4* __restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= total) return;
  if (keys[j] == kInvalidPointKey) return;
  const uint32_t i = vals[j];
  out[j] = make_float4(pts[3 * size_t(i)], pts[3 * size_t(i) + 1], pts[3 * size_t(i) + 2],
                       __uint_as_float(cols[i]));
}

// MergedTsdfIntegrator::integrateVoxel, first half: the reference's *sequential* weighted mean and
// colour blend over the points of a bundle (bit-exact: the merged point decides which voxels and
// blocks the ray visits).  Persistent groups of 4 lanes: a group owns one bundle at a time
// (fetched from a global counter), loads 32 points with 8 coalesced 64-byte reads, then runs the
// recurrence over them from shuffled operands; a warp advances 8 independent recurrences.
constexpr int kFoldGroup = 4;
constexpr int kFoldChunk = 32;  // points per group per round (8 per lane)

__global__ void __launch_bounds__(128)
k_fold_bundles(IntegratorParams P, const uint64_t* __restrict__ keys, uint32_t total,
               const uint32_t* __restrict__ heads, const uint32_t* __restrict__ num_heads,
               const float4* __restrict__ sorted, uint32_t* work_counter, Ray* __restrict__ folded) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (kFoldGroup - 1);
  const int gbase = lane & ~(kFoldGroup - 1);
  const uint32_t nb = *num_heads;
  uint32_t cur = 0, end = 0, my_b = 0, frame_clr = 0;
  bool finished = false, clearing = false;
  FoldState st;
  fold_reset(st);
  for (;;) {
    const bool need = !finished && cur >= end;  // uniform within a group
    const unsigned m = __ballot_sync(full, need && sub == 0);
    if (m) {
      const int leader = __ffs(m) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(work_counter, static_cast<uint32_t>(__popc(m)));
      base = __shfl_sync(full, base, leader);
      if (need) {
        const uint32_t b = base + __popc(m & ((1u << gbase) - 1u));
        if (b < nb) {
          const uint32_t start = heads[b];
          const uint64_t key = keys[start];
          if (key == kInvalidPointKey) {  // sentinel bundle of dropped points
            if (sub == 0) folded[b].frame_clr = kNoRay;
            cur = end = 0;                // fetch again on the next round
          } else {
            cur = start;
            end = (b + 1 < nb) ? heads[b + 1] : total;
            my_b = b;
            clearing = (key >> kBundleClearBit) & 1;
            frame_clr = static_cast<uint32_t>(key >> kBundleFrameShift) |
                        (clearing ? 0x80000000u : 0u);
            fold_reset(st);
          }
        } else {
          finished = true;
        }
      }
    }
    if (__all_sync(full, finished)) break;
    const bool work = !finished && cur < end;
    // a clearing bundle only uses its first point: do not stream the rest
    const uint32_t n = work ? min(static_cast<uint32_t>(clearing ? kFoldGroup : kFoldChunk),
                                  end - cur)
                            : 0u;
    float4 q[kFoldChunk / kFoldGroup];
#pragma unroll
    for (int u = 0; u < kFoldChunk / kFoldGroup; ++u) {
      const uint32_t o = u * kFoldGroup + sub;
      q[u] = (o < n) ? sorted[cur + o] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    const uint32_t n_max = __reduce_max_sync(full, n);
    bool done = false;
#pragma unroll
    for (int t = 0; t < kFoldChunk; ++t) {
      if (t < static_cast<int>(n_max)) {  // warp-uniform: shuffles are executed by all lanes
        const int src = gbase + (t & (kFoldGroup - 1));
        const float4 v = q[t / kFoldGroup];
        const float px = __shfl_sync(full, v.x, src), py = __shfl_sync(full, v.y, src);
        const float pz = __shfl_sync(full, v.z, src), pw = __shfl_sync(full, v.w, src);
        if (t < static_cast<int>(n) && !done) {
          const float w = voxel_weight(P, pz);
          if (!(w < kEps)) {
            fold_step(st, px, py, pz, __float_as_uint(pw), w);
            done = clearing;  // only the first point of a clearing bundle is used
          }
        }
      }
    }
    if (work) {
      cur = done ? end : cur + n;
      if (cur >= end && sub == 0) {
        Ray r;
        r.px = st.m.x;  // camera frame; k_bundle_rays moves it to the global frame
        r.py = st.m.y;
        r.pz = st.m.z;
        r.weight = st.W;
        r.color = fold_color(st);
        r.frame_clr = frame_clr;
        folded[my_b] = r;
      }
    }
  }
}

// one thread per bundle: T_G_C * merged point, ray set-up, pair count
__global__ void k_bundle_rays(IntegratorParams P, const float* __restrict__ poses,
                              const uint32_t* __restrict__ num_heads, Ray* __restrict__ rays,
                              uint32_t* __restrict__ ray_count) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= *num_heads) return;
  Ray ray = rays[b];
  if (ray.frame_clr == kNoRay) {
    ray_count[b] = 0;
    return;
  }
  const Xform T = make_xform(poses + 7 * (ray.frame_clr & 0x7FFFFFFFu));
  const V3 pg = apply(T, V3{ray.px, ray.py, ray.pz});
  RayCaster rc;
  rc.init(T.t, pg, (ray.frame_clr >> 31) != 0, P.carving != 0, P.max_ray, P.voxel_size_inv,
          P.trunc);
  rays[b].px = pg.x;
  rays[b].py = pg.y;
  rays[b].pz = pg.z;
  ray_count[b] = rc.valid ? rc.steps + 1u : 0u;
}

// SIMPLE: one ray per valid point, in (frame, visit rank) order
struct ValidSlot {
  IntegratorParams P;
  const uint64_t* offs;
  int F;
  const float* pts;
  __device__ __forceinline__ bool operator()(uint32_t g) const {
    const int f = frame_of(offs, F, g);
    const uint64_t base = offs[f];
    const int n = static_cast<int>(offs[f + 1] - base);
    bool clearing;
    return point_valid(P, load_point(pts, base + order_index(static_cast<int>(g - base), n,
                                                             P.order_mode)), &clearing);
  }
};

__global__ void k_simple_rays(IntegratorParams P, const float* __restrict__ poses,
                              const uint64_t* __restrict__ offs, int F,
                              const uint32_t* __restrict__ slots,
                              const uint32_t* __restrict__ num_slots, const float* __restrict__ pts,
                              const uint32_t* __restrict__ cols, Ray* __restrict__ rays,
                              uint32_t* __restrict__ ray_count) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= *num_slots) return;
  const uint32_t g = slots[r];
  const int f = frame_of(offs, F, g);
  const uint64_t base = offs[f];
  const int n = static_cast<int>(offs[f + 1] - base);
  const size_t i = base + order_index(static_cast<int>(g - base), n, P.order_mode);
  const V3 pc = load_point(pts, i);
  bool clearing = false;
  point_valid(P, pc, &clearing);
  const Xform T = make_xform(poses + 7 * f);
  const V3 pg = apply(T, pc);
  RayCaster rc;
  rc.init(T.t, pg, clearing, P.carving != 0, P.max_ray, P.voxel_size_inv, P.trunc);
  Ray ray;
  ray.px = pg.x;
  ray.py = pg.y;
  ray.pz = pg.z;
  ray.weight = voxel_weight(P, pc.z);
  ray.color = cols[i];
  ray.frame_clr = static_cast<uint32_t>(f) | (clearing ? 0x80000000u : 0u);
  rays[r] = ray;
  ray_count[r] = rc.valid ? rc.steps + 1u : 0u;
}

__global__ void k_totals(const uint32_t* num_rays, const uint32_t* ray_count,
                         const uint32_t* ray_offset, size_t upper, CallCounters* c) {
  c->rays = *num_rays;
  c->pairs = upper ? static_cast<unsigned long long>(ray_offset[upper - 1]) + ray_count[upper - 1]
                   : 0ull;
  c->touched = 0;
}

// ------------------------------------------------------------------ back half
// Walk every ray (Amanatides-Woo DDA exactly as voxblox::RayCaster), allocate every block it
// visits (R4: allocation on first visit), emit one (hash entry << 12 | voxel) key per visit.
// One lane per ray; every kWalkRound steps the warp flushes the keys staged in shared memory so
// that each ray's run goes out as contiguous 128-byte rows instead of 32 scattered words.
constexpr int kWalkRound = 32;
constexpr int kWalkWarps = 4;
template <class K>
__global__ void __launch_bounds__(kWalkWarps * 32)
k_ray_walk(IntegratorParams P, const float* __restrict__ poses, const Ray* __restrict__ rays,
           const uint32_t* __restrict__ ray_offset, uint32_t num_rays, LayerView L,
           K* __restrict__ pkeys, uint32_t* __restrict__ pvals) {
  __shared__ K stage[kWalkWarps][32][kWalkRound + 1];
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  RayCaster rc;
  rc.valid = false;
  rc.steps = 0;
  uint32_t out = 0;
  if (r < num_rays) {
    const Ray ray = rays[r];
    if (ray.frame_clr != kNoRay) {
      const uint32_t frame = ray.frame_clr & 0x7FFFFFFFu;
      const float* T = poses + 7 * frame;
      rc.init(V3{T[4], T[5], T[6]}, V3{ray.px, ray.py, ray.pz}, (ray.frame_clr >> 31) != 0,
              P.carving != 0, P.max_ray, P.voxel_size_inv, P.trunc);
      if (!rc.valid && !rc.in_range) atomicOr(L.err, kErrOutOfRange);
      out = ray_offset[r];
    }
  }
  uint32_t remaining = rc.valid ? rc.steps + 1u : 0u;
  int lbx = 0x7FFFFFFF, lby = 0, lbz = 0;
  K entry_bits = 0;
  while (__any_sync(full, remaining > 0)) {
    const uint32_t n = min(remaining, static_cast<uint32_t>(kWalkRound));
    for (uint32_t s = 0; s < n; ++s) {
      const int bx = rc.cx >> 4, by = rc.cy >> 4, bz = rc.cz >> 4;
      if (bx != lbx || by != lby || bz != lbz) {
        lbx = bx;
        lby = by;
        lbz = bz;
        entry_bits = static_cast<K>(L.insert_entry(pack_block_key(bx, by, bz))) << 12;
      }
      const int lin = (rc.cx & 15) + 16 * ((rc.cy & 15) + 16 * (rc.cz & 15));
      stage[wib][lane][s] = entry_bits | static_cast<K>(lin);
      rc.step();
    }
    __syncwarp();
    // flush: row l = the n_l keys of lane l's ray, contiguous at out_l
    for (int l = 0; l < 32; ++l) {
      const uint32_t n_l = __shfl_sync(full, n, l);
      if (n_l == 0) continue;
      const uint32_t out_l = __shfl_sync(full, out, l);
      const uint32_t r_l = __shfl_sync(full, r, l);
      if (static_cast<uint32_t>(lane) < n_l) {
        pkeys[out_l + lane] = stage[wib][l][lane];
        pvals[out_l + lane] = r_l;
      }
    }
    __syncwarp();
    out += n;
    remaining -= n;
  }
}

template <class K>
struct SegmentHead {
  const K* keys;
  __device__ __forceinline__ bool operator()(uint32_t i) const {
    return i == 0 || keys[i] != keys[i - 1];
  }
};

// x -> clamp(a x + b, lo, hi), a >= 0.  Closed under composition.
struct ClampedAffine {
  float a, b, lo, hi;
};
__device__ __forceinline__ ClampedAffine compose(const ClampedAffine& f, const ClampedAffine& g) {
  // g after f
  ClampedAffine r;
  r.a = g.a * f.a;
  r.b = g.a * f.b + g.b;
  r.lo = fminf(fmaxf(g.a * f.lo + g.b, g.lo), g.hi);
  r.hi = fminf(fmaxf(g.a * f.hi + g.b, g.lo), g.hi);
  return r;
}
constexpr float kBig = 1.0e30f;

// state-independent part of updateTsdfVoxel for one (ray, voxel) visit
struct Visit {
  float sdf, w;
  uint32_t col;
};
__device__ __forceinline__ Visit make_visit(const IntegratorParams& P, const float* __restrict__ poses,
                                            const Ray& ray, V3 center) {
  const float* T = poses + 7 * (ray.frame_clr & 0x7FFFFFFFu);
  const V3 origin = V3{T[4], T[5], T[6]};
  const V3 pg = V3{ray.px, ray.py, ray.pz};
  // computeDistance + weight drop-off / sparsity compensation
  const V3 v_voxel_origin = center - origin;
  const V3 v_point_origin = pg - origin;
  const float dist_G = norm3(v_point_origin);
  const float dist_G_V = dot3(v_voxel_origin, v_point_origin) / dist_G;
  Visit v;
  v.sdf = dist_G - dist_G_V;
  v.w = ray.weight;
  if (P.weight_dropoff && v.sdf < -P.voxel_size) {
    v.w = v.w * (P.trunc + v.sdf) / (P.trunc - P.voxel_size);
    v.w = fmaxf(v.w, 0.0f);
  }
  if (P.use_sparsity && fabsf(v.sdf) < P.trunc) v.w *= P.sparsity_factor;
  v.col = ray.color;
  return v;
}

// voxel addressed by a pair key
struct VoxelRef {
  int slot;
  uint32_t entry;
  V3 center;
  float* dp;
  float* wp;
  uint32_t* cp;
};
template <class K>
__device__ __forceinline__ VoxelRef voxel_ref(const IntegratorParams& P, const LayerView& L, K key) {
  VoxelRef r;
  r.entry = static_cast<uint32_t>(key >> 12);
  const int lin = static_cast<int>(key & 4095);
  r.slot = L.hash_vals[r.entry];
  int bx, by, bz;
  unpack_block_key(L.hash_keys[r.entry], bx, by, bz);
  r.center = V3{center_coord(bx * 16 + (lin & 15), P.voxel_size),
                center_coord(by * 16 + ((lin >> 4) & 15), P.voxel_size),
                center_coord(bz * 16 + (lin >> 8), P.voxel_size)};
  const int s = r.slot < 0 ? 0 : r.slot;
  r.dp = L.dist_plane(s) + lin;
  r.wp = L.weight_plane(s) + lin;
  r.cp = L.color_plane(s) + lin;
  return r;
}

// Replays updates [start, end) of one voxel on (D, W, C), 32 at a time: weights by prefix sum,
// the clamped weighted average as an ordered tree reduction of clamped affine maps, colours
// sequentially for the updates inside the truncation band.  The ray records of chunk c+1 and the
// ray ids of chunk c+2 are in flight while chunk c is reduced.
__device__ __forceinline__ void replay_segment(const IntegratorParams& P,
                                               const float* __restrict__ poses,
                                               const Ray* __restrict__ rays,
                                               const uint32_t* __restrict__ pvals, uint32_t start,
                                               uint32_t end, V3 center, int lane, float& D, float& W,
                                               uint32_t& C) {
  const unsigned full = 0xFFFFFFFFu;
  uint32_t id_next = (start + lane < end) ? pvals[start + lane] : 0u;
  Ray ray_next = rays[id_next];
  id_next = (start + 32 + lane < end) ? pvals[start + 32 + lane] : 0u;
  for (uint32_t c0 = start; c0 < end; c0 += 32) {
    const Ray ray = ray_next;
    if (c0 + 32 < end) {
      ray_next = rays[id_next];
      id_next = (c0 + 64 + lane < end) ? pvals[c0 + 64 + lane] : 0u;
    }
    const bool act = c0 + lane < end;
    Visit v = make_visit(P, poses, ray, center);
    if (!act) v.w = 0.0f;
    const float sdf = v.sdf, w = v.w;
    // weights: W_k = min(max_weight, W_{k-1} + w_k)  ==  min(max_weight, W_0 + sum w)
    float pre = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_up_sync(full, pre, d);
      if (lane >= d) pre += o;
    }
    const float w_prev = fminf(P.max_weight, W + (pre - w));
    const float w_new = w_prev + w;
    const bool skip = !act || w_new < kEps;  // "new_weight < kFloatEpsilon -> return"
    const float inv = __fdividef(1.0f, w_new);
    ClampedAffine f;
    f.a = skip ? 1.0f : w_prev * inv;
    f.b = skip ? 0.0f : (sdf * w) * inv;
    f.lo = skip ? -kBig : -P.trunc;
    f.hi = skip ? kBig : P.trunc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {  // ordered tree reduction into lane 0
      ClampedAffine g;
      g.a = __shfl_down_sync(full, f.a, d);
      g.b = __shfl_down_sync(full, f.b, d);
      g.lo = __shfl_down_sync(full, f.lo, d);
      g.hi = __shfl_down_sync(full, f.hi, d);
      if ((lane & (2 * d - 1)) == 0) f = compose(f, g);
    }
    const float fa = __shfl_sync(full, f.a, 0), fb = __shfl_sync(full, f.b, 0);
    const float flo = __shfl_sync(full, f.lo, 0), fhi = __shfl_sync(full, f.hi, 0);
    D = fminf(fmaxf(fa * D + fb, flo), fhi);
    unsigned band = __ballot_sync(full, !skip && fabsf(sdf) < P.trunc);
    while (band) {
      const int t = __ffs(band) - 1;
      band &= band - 1;
      C = blend_colors(C, __shfl_sync(full, w_prev, t), __shfl_sync(full, v.col, t),
                       __shfl_sync(full, w, t));
    }
    const float w_after = skip ? w_prev : fminf(P.max_weight, w_new);
    W = __shfl_sync(full, w_after, 31);
  }
}

// Long segments (>= kLongSegment updates: the voxels next to the sensor, crossed by every ray)
// are split into sub-blocks that many warps reduce in parallel, see k_long_partials.
constexpr uint32_t kLongSegment = 2048;
constexpr uint32_t kLongSub = 1024;
struct LongSeg {
  uint32_t start, end;    // pair range
  uint32_t item_base;     // first sub-block index
  uint32_t pad;
};
struct LongPartial {
  float sum_w;            // sum of the effective weights of the sub-block
  uint32_t not_free;      // any update with sdf < trunc (not a pure free-space observation)
};

// One warp per voxel segment (R5 updateTsdfVoxel replayed over the voxel's update list);
// segments are handed out dynamically, kSegBatch at a time.
constexpr uint32_t kSegBatch = 8;
template <class K>
__global__ void __launch_bounds__(256)
k_voxel_update(IntegratorParams P, const float* __restrict__ poses, const Ray* __restrict__ rays,
               const K* __restrict__ pkeys, const uint32_t* __restrict__ pvals, uint32_t num_pairs,
               const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ num_segs,
               uint32_t* work_counter, unsigned long long* long_counter, LongSeg* long_list,
               uint32_t long_cap, LayerView L, CallCounters* counters) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t ns = *num_segs;
  for (;;) {
    uint32_t s0 = 0;
    if (lane == 0) s0 = atomicAdd(work_counter, kSegBatch);
    s0 = __shfl_sync(full, s0, 0);
    if (s0 >= ns) break;
    const uint32_t s1 = min(ns, s0 + kSegBatch);
    for (uint32_t s = s0; s < s1; ++s) {
      const uint32_t start = seg_start[s];
      const uint32_t end = (s + 1 < ns) ? seg_start[s + 1] : num_pairs;
      const K key = pkeys[start];
      const VoxelRef vr = voxel_ref(P, L, key);
      if (vr.slot < 0) continue;  // pool exhausted; error already flagged
      if (lane == 0 && (start == 0 || static_cast<uint32_t>(pkeys[start - 1] >> 12) != vr.entry)) {
        atomicAdd(&counters->touched, 1ull);
        L.updated[vr.slot] = 1;
      }
      if (end - start >= kLongSegment) {
        if (lane == 0) {
          const uint32_t nsub = (end - start + kLongSub - 1) / kLongSub;
          // one 64-bit atomic hands out the list index (high word) and the sub-block range (low)
          const unsigned long long old =
              atomicAdd(long_counter, (1ull << 32) | static_cast<unsigned long long>(nsub));
          const uint32_t idx = static_cast<uint32_t>(old >> 32);
          if (idx < long_cap) long_list[idx] = LongSeg{start, end, static_cast<uint32_t>(old), 0u};
        }
        continue;
      }
      float D = *vr.dp, W = *vr.wp;
      uint32_t C = *vr.cp;
      replay_segment(P, poses, rays, pvals, start, end, vr.center, lane, D, W, C);
      if (lane == 0) {
        *vr.dp = D;
        *vr.wp = W;
        *vr.cp = C;
      }
    }
  }
}

// sub-block t of the long segments: sum of weights + "all free space" flag (one warp each)
template <class K>
__global__ void __launch_bounds__(256)
k_long_partials(IntegratorParams P, const float* __restrict__ poses, const Ray* __restrict__ rays,
                const K* __restrict__ pkeys, const uint32_t* __restrict__ pvals,
                const unsigned long long* __restrict__ long_counter,
                const LongSeg* __restrict__ long_list, uint32_t long_cap, LayerView L,
                LongPartial* __restrict__ partials) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t num_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned long long cnt = *long_counter;
  const uint32_t nlong = min(static_cast<uint32_t>(cnt >> 32), long_cap);
  const uint32_t nitems = static_cast<uint32_t>(cnt);
  for (uint32_t t = warp; t < nitems; t += num_warps) {
    // long_list is ordered by item_base (both come from the same atomic): binary search
    uint32_t lo = 0, hi = nlong;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (long_list[mid].item_base <= t) lo = mid; else hi = mid;
    }
    const LongSeg seg = long_list[lo];
    if (t < seg.item_base) continue;  // list overflowed: handled by the slow path elsewhere
    const uint32_t a = seg.start + (t - seg.item_base) * kLongSub;
    const uint32_t b = min(seg.end, a + kLongSub);
    const VoxelRef vr = voxel_ref(P, L, pkeys[seg.start]);
    float sum = 0.0f;
    bool not_free = false;
    for (uint32_t c0 = a; c0 < b; c0 += 32) {
      const uint32_t j = c0 + lane;
      float w = 0.0f;
      if (j < b) {
        const Visit v = make_visit(P, poses, rays[pvals[j]], vr.center);
        w = v.w;
        not_free = not_free || !(v.sdf >= P.trunc);
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) w += __shfl_xor_sync(full, w, d);
      sum += w;
    }
    const bool any_not_free = __any_sync(full, not_free);
    if (lane == 0) partials[t] = LongPartial{sum, any_not_free ? 1u : 0u};
  }
}

// one warp per long segment: closed form when every update is a free-space observation of a
// voxel that is fresh or already at +truncation (then every step of the reference clamps to
// +truncation again), otherwise the general replay.
template <class K>
__global__ void __launch_bounds__(256)
k_long_finish(IntegratorParams P, const float* __restrict__ poses, const Ray* __restrict__ rays,
              const K* __restrict__ pkeys, const uint32_t* __restrict__ pvals,
              const unsigned long long* __restrict__ long_counter,
              const LongSeg* __restrict__ long_list, uint32_t long_cap, LayerView L,
              const LongPartial* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t num_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned long long cnt = *long_counter;
  const uint32_t nlong = min(static_cast<uint32_t>(cnt >> 32), long_cap);
  for (uint32_t i = warp; i < nlong; i += num_warps) {
    const LongSeg seg = long_list[i];
    const VoxelRef vr = voxel_ref(P, L, pkeys[seg.start]);
    float D = *vr.dp, W = *vr.wp;
    uint32_t C = *vr.cp;
    const uint32_t nsub = (seg.end - seg.start + kLongSub - 1) / kLongSub;
    float sum = 0.0f;
    bool not_free = false;
    for (uint32_t t = 0; t < nsub; ++t) {  // fixed order: deterministic
      const LongPartial p = partials[seg.item_base + t];
      sum += p.sum_w;
      not_free = not_free || p.not_free != 0;
    }
    const bool fresh_or_saturated = (W == 0.0f) || (D == P.trunc);
    if (!not_free && fresh_or_saturated && P.trunc > 0.0f) {
      const float w_total = W + sum;
      if (!(w_total < kEps)) {
        D = P.trunc;
        W = fminf(P.max_weight, w_total);
      }
    } else {
      replay_segment(P, poses, rays, pvals, seg.start, seg.end, vr.center, lane, D, W, C);
    }
    if (lane == 0) {
      *vr.dp = D;
      *vr.wp = W;
      *vr.cp = C;
    }
  }
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
  CG_CUDA(ctx->seg_start.reserve(num_pairs * sizeof(uint32_t)));
  uint32_t* d_num = ctx->d_select_count;
  {
    StageScope sc(ctx, kStageRayWalk, 1);
    k_ray_walk<K><<<grid_for(num_rays, 128), 128, 0, s>>>(
        P, ctx->poses.as<float>(), ctx->rays.as<Ray>(), ctx->ray_offset.as<uint32_t>(), num_rays,
        L->v, ctx->pkey_a.as<K>(), ctx->pval_a.as<uint32_t>());
  }
  cub::DoubleBuffer<K> dk(ctx->pkey_a.as<K>(), ctx->pkey_b.as<K>());
  cub::DoubleBuffer<uint32_t> dv(ctx->pval_a.as<uint32_t>(), ctx->pval_b.as<uint32_t>());
  size_t tmp = 0, tmp2 = 0;
  CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, num_pairs, 0, key_bits, s));
  thrust::counting_iterator<uint32_t> iota(0);
  CG_CUDA(cub::DeviceSelect::If(nullptr, tmp2, iota, ctx->seg_start.as<uint32_t>(), d_num,
                                static_cast<int>(num_pairs), SegmentHead<K>{nullptr}, s));
  CG_CUDA(ctx->cub_tmp.reserve(std::max(tmp, tmp2)));
  {
    StageScope sc(ctx, kStagePairSort, 0);
    CG_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, dk, dv, num_pairs, 0, key_bits, s));
  }
  {
    StageScope sc(ctx, kStageSegments, 0);
    CG_CUDA(cub::DeviceSelect::If(ctx->cub_tmp.p, tmp2, iota, ctx->seg_start.as<uint32_t>(), d_num,
                                  static_cast<int>(num_pairs), SegmentHead<K>{dk.Current()}, s));
  }
  {
    StageScope sc(ctx, kStageVoxelUpdate, 3);
    const uint32_t long_cap = static_cast<uint32_t>(num_pairs / kLongSegment + 1);
    const size_t max_items = num_pairs / kLongSub + long_cap + 1;
    CG_CUDA(ctx->long_list.reserve(long_cap * sizeof(LongSeg)));
    CG_CUDA(ctx->long_partials.reserve(max_items * sizeof(LongPartial)));
    CG_CUDA(cudaMemsetAsync(ctx->d_work_counter, 0, sizeof(uint32_t), s));
    CG_CUDA(cudaMemsetAsync(ctx->d_long_counter, 0, sizeof(unsigned long long), s));
    k_voxel_update<K><<<ctx->num_sms * 8, 256, 0, s>>>(
        P, ctx->poses.as<float>(), ctx->rays.as<Ray>(), dk.Current(), dv.Current(),
        static_cast<uint32_t>(num_pairs), ctx->seg_start.as<uint32_t>(), d_num,
        ctx->d_work_counter, ctx->d_long_counter, ctx->long_list.as<LongSeg>(), long_cap, L->v,
        ctx->d_counters);
    k_long_partials<K><<<ctx->num_sms * 8, 256, 0, s>>>(
        P, ctx->poses.as<float>(), ctx->rays.as<Ray>(), dk.Current(), dv.Current(),
        ctx->d_long_counter, ctx->long_list.as<LongSeg>(), long_cap, L->v,
        ctx->long_partials.as<LongPartial>());
    k_long_finish<K><<<ctx->num_sms, 256, 0, s>>>(
        P, ctx->poses.as<float>(), ctx->rays.as<Ray>(), dk.Current(), dv.Current(),
        ctx->d_long_counter, ctx->long_list.as<LongSeg>(), long_cap, L->v,
        ctx->long_partials.as<LongPartial>());
  }
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
  const size_t total = offs[f1] - offs[f0];  // points of the group = upper bound on rays
  if (total == 0) return CG_OK;
  if (total > 0x7FFFFFF0ull || F > (1u << 20)) {
    set_error("too many points / frames in one group");
    return CG_ERR_INVALID_ARG;
  }
  const size_t upper = total + 1;  // + the sentinel bundle of dropped points
  CG_CUDA(ctx->rays.reserve(upper * sizeof(Ray)));
  CG_CUDA(ctx->ray_count.reserve(upper * sizeof(uint32_t)));
  CG_CUDA(ctx->ray_offset.reserve(upper * sizeof(uint32_t)));
  CG_CUDA(ctx->poses.reserve(F * 7 * sizeof(float)));
  CG_CUDA(ctx->frame_base.reserve((F + 1) * sizeof(uint64_t)));
  CG_CUDA(ctx->scan.reserve(upper * sizeof(uint32_t)));  // bundle heads / valid slots
  const bool merged = cfg->method == CG_METHOD_MERGED;
  if (merged) {
    CG_CUDA(ctx->key_a.reserve(total * sizeof(uint64_t)));
    CG_CUDA(ctx->key_b.reserve(total * sizeof(uint64_t)));
    CG_CUDA(ctx->val_a.reserve(total * sizeof(uint32_t)));
    CG_CUDA(ctx->val_b.reserve(total * sizeof(uint32_t)));
    CG_CUDA(ctx->sorted_pts.reserve(total * sizeof(float4)));
  }
  int frame_bits = 0;
  while ((size_t(1) << frame_bits) < F) ++frame_bits;
  const int bundle_bits = kBundleFrameShift + frame_bits;
  cub::DoubleBuffer<uint64_t> dk(ctx->key_a.as<uint64_t>(), ctx->key_b.as<uint64_t>());
  cub::DoubleBuffer<uint32_t> dv(ctx->val_a.as<uint32_t>(), ctx->val_b.as<uint32_t>());
  thrust::counting_iterator<uint32_t> iota(0);
  uint32_t* d_num = ctx->d_select_count;
  // group-relative frame offsets on the device
  std::vector<uint64_t> rel(F + 1);
  for (size_t f = 0; f <= F; ++f) rel[f] = offs[f0 + f] - offs[f0];
  const float* pts = d_points + 3 * offs[f0];
  const uint32_t* cols = reinterpret_cast<const uint32_t*>(d_colors) + offs[f0];
  const uint64_t* d_offs = ctx->frame_base.as<uint64_t>();
  ValidSlot valid{P, d_offs, static_cast<int>(F), pts};
  size_t tmp_sort = 0, tmp_sel = 0, tmp_scan = 0;
  if (merged) {
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, total, 0, bundle_bits, s);
    cub::DeviceSelect::If(nullptr, tmp_sel, iota, ctx->scan.as<uint32_t>(), d_num,
                          static_cast<int>(total), BundleHead{nullptr}, s);
  } else {
    cub::DeviceSelect::If(nullptr, tmp_sel, iota, ctx->scan.as<uint32_t>(), d_num,
                          static_cast<int>(total), valid, s);
  }
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, ctx->ray_count.as<uint32_t>(),
                                ctx->ray_offset.as<uint32_t>(), static_cast<int>(upper), s);
  CG_CUDA(ctx->cub_tmp.reserve(std::max(tmp_sort, std::max(tmp_sel, tmp_scan))));
  {
    StageScope sc(ctx, kStageTransfer, 0);
    CG_CUDA(cudaMemcpyAsync(ctx->poses.p, h_poses + 7 * f0, F * 7 * sizeof(float),
                            cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaMemcpyAsync(ctx->frame_base.p, rel.data(), (F + 1) * sizeof(uint64_t),
                            cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaMemsetAsync(ctx->ray_count.p, 0, upper * sizeof(uint32_t), s));
  }
  if (merged) {
    {
      StageScope sc(ctx, kStagePointKeys, 1);
      k_point_keys<<<grid_for(total, 256), 256, 0, s>>>(
          P, ctx->poses.as<float>(), d_offs, static_cast<int>(F), pts, total,
          ctx->key_a.as<uint64_t>(), ctx->val_a.as<uint32_t>(), L->v.err);
    }
    {
      StageScope sc(ctx, kStageBundleSort, 0);
      CG_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_sort, dk, dv, total, 0,
                                              bundle_bits, s));
    }
    {
      StageScope sc(ctx, kStageBundleScan, 0);
      CG_CUDA(cub::DeviceSelect::If(ctx->cub_tmp.p, tmp_sel, iota, ctx->scan.as<uint32_t>(), d_num,
                                    static_cast<int>(total), BundleHead{dk.Current()}, s));
    }
    {
      StageScope sc(ctx, kStageFold, 3);
      CG_CUDA(cudaMemsetAsync(ctx->d_work_counter, 0, sizeof(uint32_t), s));
      k_gather_sorted<<<grid_for(total, 256), 256, 0, s>>>(
          dk.Current(), dv.Current(), static_cast<uint32_t>(total), pts, cols,
          ctx->sorted_pts.as<float4>());
      k_fold_bundles<<<ctx->num_sms * 16, 128, 0, s>>>(
          P, dk.Current(), static_cast<uint32_t>(total), ctx->scan.as<uint32_t>(), d_num,
          ctx->sorted_pts.as<float4>(), ctx->d_work_counter, ctx->rays.as<Ray>());
      k_bundle_rays<<<grid_for(upper, 256), 256, 0, s>>>(P, ctx->poses.as<float>(), d_num,
                                                         ctx->rays.as<Ray>(),
                                                         ctx->ray_count.as<uint32_t>());
    }
  } else {
    {
      StageScope sc(ctx, kStageBundleScan, 0);
      CG_CUDA(cub::DeviceSelect::If(ctx->cub_tmp.p, tmp_sel, iota, ctx->scan.as<uint32_t>(), d_num,
                                    static_cast<int>(total), valid, s));
    }
    StageScope sc(ctx, kStageFold, 1);
    k_simple_rays<<<grid_for(total, 256), 256, 0, s>>>(
        P, ctx->poses.as<float>(), d_offs, static_cast<int>(F), ctx->scan.as<uint32_t>(), d_num, pts,
        cols, ctx->rays.as<Ray>(), ctx->ray_count.as<uint32_t>());
  }
  {
    StageScope sc(ctx, kStageRayScan, 1);
    CG_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_scan, ctx->ray_count.as<uint32_t>(),
                                          ctx->ray_offset.as<uint32_t>(), static_cast<int>(upper),
                                          s));
    k_totals<<<1, 1, 0, s>>>(d_num, ctx->ray_count.as<uint32_t>(), ctx->ray_offset.as<uint32_t>(),
                             upper, ctx->d_counters);
  }
  // the staged host buffer (rel) must outlive its async copy: the sync below covers it
  CG_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(CallCounters),
                          cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  CG_CUDA(cudaGetLastError());
  const uint32_t num_rays = static_cast<uint32_t>(ctx->h_counters->rays);
  const size_t num_pairs = ctx->h_counters->pairs;
  const size_t max_pairs = env_size("CG_MAX_PAIRS", size_t(768) << 20);
  if (num_pairs > max_pairs || num_pairs >= 0x7FFFFFF0ull) {
    if (F > 1) {  // the front half never touches the layer: safe to redo in two halves
      const size_t mid = f0 + F / 2;
      int32_t rc = integrate_group(L, cfg, P, h_poses, d_points, d_colors, offs, f0, mid, stats);
      if (rc) return rc;
      return integrate_group(L, cfg, P, h_poses, d_points, d_colors, offs, mid, f1, stats);
    }
    if (num_pairs >= 0x7FFFFFF0ull) {
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
  StageScope sc(ctx, kStageTransfer, 0);
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
