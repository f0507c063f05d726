// mesh_connect.cu — MeshLayer::getConnectedMesh on the device (the PLY step of
// voxgraph::SubmapVisuals::saveAndPubCombinedMesh with a file path,
// coxgraph/src/server/visualizer/server_visualizer.cpp:118-126: io::outputMeshLayerAsPly ->
// MeshLayer::getConnectedMesh -> voxblox::createConnectedMesh, [EXT] mesh/mesh_utils.h).
//
// Upstream walks the block meshes in order and keeps a hash map from the vertex position,
// discretised as round(double(v) * 1e10) per axis (approximate_vertex_proximity_threshold 1e-10),
// to the index of the first vertex that fell into that cell; every triangle corner is renumbered
// to that first vertex; positions, normals and colours of the first occurrences make up the
// connected mesh.  Here, on the mesh cg_layer_mesh left in the context ((z, y, x) block order):
//   k_weld_keys      the three 64-bit cell indices per vertex
//   3 x stable radix sort of the vertex ids by (z, y, x) cell -> equal cells adjacent, ids ascending
//   k_weld_heads + max-scan   every vertex learns the smallest id of its cell = the first occurrence
//   scan of "is first" -> new index; k_weld_write compacts and renumbers
// Output is identical to the sequential hash-map walk: vertex order = order of first occurrence.
#include <cub/cub.cuh>

#include <algorithm>

#include "cg_internal.cuh"

namespace cg {

__device__ __forceinline__ unsigned long long weld_cell(float v) {
  // LongIndex(std::round(scaled)) of mesh_utils.h, biased so that unsigned order = signed order
  // threshold_inv = 1 / double(FloatingPoint(1e-10)): the threshold parameter is a float upstream
  const double inv = 1.0 / static_cast<double>(1e-10f);
  const long long k = static_cast<long long>(round(static_cast<double>(v) * inv));
  return static_cast<unsigned long long>(k) ^ 0x8000000000000000ull;
}

struct MaxU32 {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__global__ void k_weld_keys(const float* __restrict__ vertices, uint32_t n,
                            unsigned long long* __restrict__ kx, unsigned long long* __restrict__ ky,
                            unsigned long long* __restrict__ kz, uint32_t* __restrict__ ids) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  kx[i] = weld_cell(vertices[3 * i]);
  ky[i] = weld_cell(vertices[3 * i + 1]);
  kz[i] = weld_cell(vertices[3 * i + 2]);
  ids[i] = i;
}

__global__ void k_weld_gather(const unsigned long long* __restrict__ comp,
                              const uint32_t* __restrict__ ids, uint32_t n,
                              unsigned long long* __restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = comp[ids[j]];
}

// head position of every sorted entry's cell (0 for non-heads: a max-scan propagates the heads)
__global__ void k_weld_heads(const unsigned long long* __restrict__ kx,
                             const unsigned long long* __restrict__ ky,
                             const unsigned long long* __restrict__ kz,
                             const uint32_t* __restrict__ ids, uint32_t n,
                             uint32_t* __restrict__ head_pos) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  bool head = j == 0;
  if (!head) {
    const uint32_t a = ids[j], b = ids[j - 1];
    head = kx[a] != kx[b] || ky[a] != ky[b] || kz[a] != kz[b];
  }
  head_pos[j] = head ? j : 0u;
}

// rep[i] = first occurrence of vertex i's cell; first[i] = 1 where i is one
__global__ void k_weld_representatives(const uint32_t* __restrict__ ids,
                                       const uint32_t* __restrict__ head_pos, uint32_t n,
                                       uint32_t* __restrict__ rep, uint32_t* __restrict__ first) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t i = ids[j], r = ids[head_pos[j]];  // stable sorts: the head holds the smallest id
  rep[i] = r;
  first[i] = r == i ? 1u : 0u;
}

__global__ void k_weld_write(const float* __restrict__ vertices, const float* __restrict__ normals,
                             const uint32_t* __restrict__ colors, const uint32_t* __restrict__ rep,
                             const uint32_t* __restrict__ first,
                             const uint32_t* __restrict__ new_index, uint32_t n,
                             float* __restrict__ out_v, float* __restrict__ out_n,
                             uint32_t* __restrict__ out_c, uint32_t* __restrict__ out_idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_idx[i] = new_index[rep[i]];
  if (first[i]) {
    const uint32_t o = new_index[i];
    for (int k = 0; k < 3; ++k) {
      out_v[3 * o + k] = vertices[3 * i + k];
      out_n[3 * o + k] = normals[3 * i + k];
    }
    out_c[o] = colors[i];
  }
}

}  // namespace cg

using namespace cg;

extern "C" {

int32_t cg_mesh_connect(cg_context* ctx, size_t capacity_vertices, size_t capacity_indices,
                        float* vertices, float* normals, uint8_t* colors, uint32_t* indices,
                        size_t* num_vertices_out, size_t* num_indices_out) {
  if (!ctx) return CG_ERR_INVALID_ARG;
  CG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t n = ctx->mc_total;
  if (num_vertices_out) *num_vertices_out = 0;
  if (num_indices_out) *num_indices_out = n;
  if (n == 0) return CG_OK;
  if (n >= 0x7FFFFFF0ull) {
    set_error("cg_mesh_connect: %zu vertices", n);
    return CG_ERR_INVALID_ARG;
  }
  const uint32_t N = static_cast<uint32_t>(n);
  // scratch: 3 cell components + 2 sort key buffers (u64), ids x2, head/rep/first/new (u32)
  CG_CUDA(ctx->weld_keys.reserve(5 * n * sizeof(unsigned long long)));
  CG_CUDA(ctx->weld_words.reserve(6 * n * sizeof(uint32_t)));
  CG_CUDA(ctx->weld_out.reserve(n * (6 * sizeof(float) + 2 * sizeof(uint32_t))));
  unsigned long long* kx = ctx->weld_keys.as<unsigned long long>();
  unsigned long long *ky = kx + n, *kz = ky + n, *ka = kz + n, *kb = ka + n;
  uint32_t* ida = ctx->weld_words.as<uint32_t>();
  uint32_t *idb = ida + n, *head = idb + n, *rep = head + n, *first = rep + n, *newi = first + n;
  const unsigned grid = grid_for(n, 256);
  ctx->own_launches += 6;
  k_weld_keys<<<grid, 256, 0, s>>>(ctx->mc_vertices.as<float>(), N, kx, ky, kz, ida);
  cub::DoubleBuffer<unsigned long long> dk(ka, kb);
  cub::DoubleBuffer<uint32_t> dv(ida, idb);
  size_t tmp = 0, tmp2 = 0, tmp3 = 0;
  CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, static_cast<int>(n), 0, 64, s));
  CG_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp2, head, head, MaxU32(), static_cast<int>(n), s));
  CG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp3, first, newi, static_cast<int>(n), s));
  CG_CUDA(ctx->cub_tmp.reserve(std::max(tmp, std::max(tmp2, tmp3))));
  const unsigned long long* comps[3] = {kz, ky, kx};  // least significant first
  for (int c = 0; c < 3; ++c) {
    k_weld_gather<<<grid, 256, 0, s>>>(comps[c], dv.Current(), N, dk.Current());
    CG_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, dk, dv, static_cast<int>(n), 0, 64, s));
  }
  const uint32_t* ids = dv.Current();
  k_weld_heads<<<grid, 256, 0, s>>>(kx, ky, kz, ids, N, head);
  CG_CUDA(cub::DeviceScan::InclusiveScan(ctx->cub_tmp.p, tmp2, head, head, MaxU32(), static_cast<int>(n), s));
  k_weld_representatives<<<grid, 256, 0, s>>>(ids, head, N, rep, first);
  CG_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp3, first, newi, static_cast<int>(n), s));
  float* out_v = ctx->weld_out.as<float>();
  float* out_n = out_v + 3 * n;
  uint32_t* out_c = reinterpret_cast<uint32_t*>(out_n + 3 * n);
  uint32_t* out_idx = out_c + n;
  k_weld_write<<<grid, 256, 0, s>>>(ctx->mc_vertices.as<float>(), ctx->mc_normals.as<float>(),
                                    ctx->mc_colors.as<uint32_t>(), rep, first, newi, N, out_v, out_n,
                                    out_c, out_idx);
  uint32_t last_new = 0, last_first = 0;
  CG_CUDA(cudaMemcpyAsync(&last_new, newi + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaMemcpyAsync(&last_first, first + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  CG_CUDA(cudaGetLastError());
  const size_t unique = static_cast<size_t>(last_new) + last_first;
  if (num_vertices_out) *num_vertices_out = unique;
  if (!vertices && !normals && !colors && !indices) return CG_OK;
  if ((vertices || normals || colors) && capacity_vertices < unique) {
    set_error("cg_mesh_connect: capacity %zu < %zu vertices", capacity_vertices, unique);
    return CG_ERR_INVALID_ARG;
  }
  if (indices && capacity_indices < n) {
    set_error("cg_mesh_connect: capacity %zu < %zu indices", capacity_indices, n);
    return CG_ERR_INVALID_ARG;
  }
  if (vertices) CG_CUDA(cudaMemcpyAsync(vertices, out_v, unique * 12, cudaMemcpyDeviceToHost, s));
  if (normals) CG_CUDA(cudaMemcpyAsync(normals, out_n, unique * 12, cudaMemcpyDeviceToHost, s));
  if (colors) CG_CUDA(cudaMemcpyAsync(colors, out_c, unique * 4, cudaMemcpyDeviceToHost, s));
  if (indices) CG_CUDA(cudaMemcpyAsync(indices, out_idx, n * 4, cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  return CG_OK;
}

}  // extern "C"
