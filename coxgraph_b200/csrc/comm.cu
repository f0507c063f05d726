// comm.cu — the server's global merge across the GPUs of one box, behind the C ABI (SURVEY.md
// §8b cg_comm_init / cg_gather_global, §8e).
//
// Every rank projects its submaps into a *partial* global layer; a global block is owned by rank
// cg_block_owner(index).  The exchange is an owner-pull over NVLink peer memory: each rank lists
// the slots of its partial blocks per owner, and the owner's fold kernel reads those blocks
// straight out of the peers' block pools (peer pointers opened through CUDA IPC) while it folds
// them into its own layer with voxblox::mergeLayerAintoLayerB's aligned form (Block::mergeBlock /
// mergeVoxelAIntoVoxelB, R10; reference call site of that overload:
// coxgraph/src/server/submap_collection.cpp:31-33) — gather and fold are ONE kernel, there is no
// packing pass, no staging copy and no host synchronisation between listing, exchange and fold.
// NCCL (loaded at run time from libnccl.so.2, the one torch has loaded if there is one) is the
// bootstrap and the barrier only: unique id -> communicator, one all-gather of the IPC handles
// when a partial layer is bound, and a one-word all-reduce on the stream before and after the fold
// kernels (all partial layers complete / all peers done reading).  Sources are folded in ascending
// rank order, one launch per source, so the result is deterministic.
#include <dlfcn.h>
#include <string.h>

#include <vector>

#include "cg_internal.cuh"

namespace cg {

// ---- the few NCCL entry points, resolved with dlsym (the library stays loadable without NCCL)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2 };
enum { ncclSum = 0 };
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;

static int32_t load_nccl() {
  if (g_nccl.lib) return CG_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("multi-GPU merge needs NCCL: %s", dlerror());
    return CG_ERR_UNSUPPORTED;
  }
  Nccl n;
  n.lib = h;
  n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(h, "ncclAllReduce"));
  n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
  n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce || !n.AllGather ||
      !n.GetErrorString) {
    set_error("libnccl.so.2 lacks an expected symbol");
    return CG_ERR_UNSUPPORTED;
  }
  g_nccl = n;
  return CG_OK;
}
#define CG_NCCL(expr)                                                        \
  do {                                                                       \
    const int _r = (expr);                                                   \
    if (_r != ncclSuccess) {                                                 \
      set_error("NCCL error %d (%s) at %s", _r, g_nccl.GetErrorString(_r), #expr); \
      return CG_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

// What a peer needs of a rank's partial layer (device pointers valid in the READER's process).
struct PeerLayer {
  const float* pool;
  const uint64_t* block_keys;
  const uint8_t* has_data;
  const uint32_t* lists;   // [nranks][list_cap] pool slots, grouped by owner
  const uint32_t* counts;  // [nranks]
};
constexpr int kMaxRanks = 64;

struct IpcRecord {  // one per rank, all-gathered
  cudaIpcMemHandle_t pool, keys, flags, shared;
  unsigned long long off_pool, off_keys, off_flags, off_shared;  // pointer - allocation base
  unsigned long long max_blocks;
  float voxel_size;
  int pad;
};

}  // namespace cg

struct cg_comm {
  cg::ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int* d_token = nullptr;             // barrier word
  // binding of one partial layer
  const cg_layer* bound = nullptr;
  uint32_t* shared = nullptr;         // [nranks * list_cap] lists, then [nranks] counts (this rank's)
  size_t list_cap = 0;
  std::vector<void*> opened;          // peer mappings to close
  cg::PeerLayer* d_peers = nullptr;   // [nranks]
};

namespace cg {

static __host__ __device__ __forceinline__ uint32_t comm_owner_of(uint64_t key, uint32_t nranks) {
  return (hash_key(key ^ 0x9E3779B97F4A7C15ULL) >> 7) % nranks;  // == cg_block_owner
}

// slots of the partial layer's blocks, grouped by owner
__global__ void k_owner_lists(LayerView L, int n, uint32_t nranks, uint32_t list_cap,
                              uint32_t* __restrict__ lists, uint32_t* __restrict__ counts) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  const uint32_t o = comm_owner_of(L.block_keys[slot], nranks);
  const uint32_t pos = atomicAdd(&counts[o], 1u);
  if (pos < list_cap) lists[static_cast<size_t>(o) * list_cap + pos] = static_cast<uint32_t>(slot);
}

// The blocks rank `src` holds for this rank, read through the peer mapping and folded into the
// owner's layer: one CTA per block (grid-stride), 16-byte loads across NVLink.
__global__ void __launch_bounds__(256)
k_fold_pull(LayerView B, const PeerLayer* __restrict__ peers, int src, int me, uint32_t list_cap,
            unsigned long long* folded) {
  const PeerLayer P = peers[src];
  const uint32_t n = min(P.counts[me], list_cap);
  __shared__ int s_slot;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t a = P.lists[static_cast<size_t>(me) * list_cap + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      s_slot = -1;
      if (P.has_data[a]) {  // Block::mergeBlock: a source block without data is skipped
        const int e = B.insert_entry(P.block_keys[a]);
        s_slot = B.hash_vals[e];
        if (s_slot >= 0) {
          B.has_data[s_slot] = 1;
          B.updated[s_slot] = 1;
          atomicAdd(folded, 1ull);
        }
      }
    }
    __syncthreads();
    const int slot = s_slot;
    if (slot < 0) continue;
    const uint4* sd = reinterpret_cast<const uint4*>(P.pool + static_cast<size_t>(a) * (3 * kVoxelsPerBlock));
    const uint4* sw = sd + kVoxelsPerBlock / 4;
    const uint4* sc = sw + kVoxelsPerBlock / 4;
    uint4* dd = reinterpret_cast<uint4*>(B.dist_plane(slot));
    uint4* dw = reinterpret_cast<uint4*>(B.weight_plane(slot));
    uint4* dc = reinterpret_cast<uint4*>(B.color_plane(slot));
    for (int q = threadIdx.x; q < kVoxelsPerBlock / 4; q += blockDim.x) {
      const uint4 ad = sd[q], aw = sw[q], ac = sc[q];
      uint4 bd = dd[q], bw = dw[q], bc = dc[q];
      VoxelState v0{__uint_as_float(bd.x), __uint_as_float(bw.x), bc.x};
      VoxelState v1{__uint_as_float(bd.y), __uint_as_float(bw.y), bc.y};
      VoxelState v2{__uint_as_float(bd.z), __uint_as_float(bw.z), bc.z};
      VoxelState v3{__uint_as_float(bd.w), __uint_as_float(bw.w), bc.w};
      merge_voxel(__uint_as_float(ad.x), __uint_as_float(aw.x), ac.x, v0);
      merge_voxel(__uint_as_float(ad.y), __uint_as_float(aw.y), ac.y, v1);
      merge_voxel(__uint_as_float(ad.z), __uint_as_float(aw.z), ac.z, v2);
      merge_voxel(__uint_as_float(ad.w), __uint_as_float(aw.w), ac.w, v3);
      dd[q] = make_uint4(__float_as_uint(v0.d), __float_as_uint(v1.d), __float_as_uint(v2.d), __float_as_uint(v3.d));
      dw[q] = make_uint4(__float_as_uint(v0.w), __float_as_uint(v1.w), __float_as_uint(v2.w), __float_as_uint(v3.w));
      dc[q] = make_uint4(v0.c, v1.c, v2.c, v3.c);
    }
  }
}

// pointer -> (allocation base handle, offset): cudaMalloc may place small buffers inside a larger
// allocation, and an IPC handle always names the whole allocation
static int32_t ipc_of(const void* p, cudaIpcMemHandle_t* h, unsigned long long* off) {
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  static GetRange get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CG_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    get_range = reinterpret_cast<GetRange>(fn);
    if (!get_range) {
      set_error("cuMemGetAddressRange is not available");
      return CG_ERR_CUDA;
    }
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, reinterpret_cast<unsigned long long>(p)) != 0) {
    set_error("cuMemGetAddressRange failed");
    return CG_ERR_CUDA;
  }
  *off = reinterpret_cast<unsigned long long>(p) - base;
  CG_CUDA(cudaIpcGetMemHandle(h, reinterpret_cast<void*>(base)));
  return CG_OK;
}

static void unbind(cg_comm* c) {
  for (void* p : c->opened) cudaIpcCloseMemHandle(p);
  c->opened.clear();
  if (c->shared) cudaFree(c->shared);
  c->shared = nullptr;
  c->bound = nullptr;
}

static int32_t barrier(cg_context* ctx) {
  cg_comm* c = ctx->comm;
  CG_NCCL(g_nccl.AllReduce(c->d_token, c->d_token, 1, ncclInt32, ncclSum, c->comm, ctx->stream));
  return CG_OK;
}

// All ranks call this with their own partial layer (collective): lists + IPC handles exchanged,
// peer mappings opened.  Done once per partial layer (it stays bound until another one is used).
static int32_t bind_partial(cg_context* ctx, const cg_layer* partial) {
  cg_comm* c = ctx->comm;
  unbind(c);
  cudaStream_t s = ctx->stream;
  const int R = c->nranks;
  c->list_cap = partial->max_blocks;
  const size_t words = static_cast<size_t>(R) * c->list_cap + R;
  CG_CUDA(cudaMalloc(&c->shared, words * sizeof(uint32_t)));
  CG_CUDA(cudaMemsetAsync(c->shared, 0, words * sizeof(uint32_t), s));
  IpcRecord mine;
  memset(&mine, 0, sizeof(mine));
  int32_t rc;
  if ((rc = ipc_of(partial->v.pool, &mine.pool, &mine.off_pool))) return rc;
  if ((rc = ipc_of(partial->v.block_keys, &mine.keys, &mine.off_keys))) return rc;
  if ((rc = ipc_of(partial->v.has_data, &mine.flags, &mine.off_flags))) return rc;
  if ((rc = ipc_of(c->shared, &mine.shared, &mine.off_shared))) return rc;
  mine.max_blocks = partial->max_blocks;
  mine.voxel_size = partial->v.voxel_size;
  void* d_all = nullptr;
  CG_CUDA(cudaMalloc(&d_all, sizeof(IpcRecord) * R));
  CG_CUDA(cudaMemcpyAsync(static_cast<char*>(d_all) + sizeof(IpcRecord) * c->rank, &mine,
                          sizeof(IpcRecord), cudaMemcpyHostToDevice, s));
  CG_NCCL(g_nccl.AllGather(static_cast<char*>(d_all) + sizeof(IpcRecord) * c->rank, d_all,
                           sizeof(IpcRecord), ncclInt8, c->comm, s));
  std::vector<IpcRecord> all(R);
  CG_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(IpcRecord) * R, cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_all);
  std::vector<PeerLayer> peers(R);
  for (int r = 0; r < R; ++r) {
    if (all[r].max_blocks != partial->max_blocks || all[r].voxel_size != partial->v.voxel_size) {
      set_error("cg_gather_global: rank %d's partial layer differs (max_blocks / voxel size)", r);
      return CG_ERR_INVALID_ARG;
    }
    if (r == c->rank) {
      peers[r] = PeerLayer{partial->v.pool, partial->v.block_keys, partial->v.has_data, c->shared,
                           c->shared + static_cast<size_t>(R) * c->list_cap};
      continue;
    }
    // several buffers of a peer may live in one allocation: a handle can be opened once only
    const cudaIpcMemHandle_t* hs[4] = {&all[r].pool, &all[r].keys, &all[r].flags, &all[r].shared};
    const unsigned long long offs[4] = {all[r].off_pool, all[r].off_keys, all[r].off_flags,
                                        all[r].off_shared};
    char* mapped[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) {
      for (int j = 0; j < k; ++j)
        if (memcmp(hs[j], hs[k], sizeof(cudaIpcMemHandle_t)) == 0) mapped[k] = mapped[j];
      if (!mapped[k]) {
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, *hs[k], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          set_error("cudaIpcOpenMemHandle (rank %d's partial layer) failed: %s", r,
                    cudaGetErrorString(e));
          return CG_ERR_CUDA;
        }
        c->opened.push_back(p);
        mapped[k] = static_cast<char*>(p);
      }
    }
    const uint32_t* sh = reinterpret_cast<const uint32_t*>(mapped[3] + offs[3]);
    peers[r] = PeerLayer{reinterpret_cast<const float*>(mapped[0] + offs[0]),
                         reinterpret_cast<const uint64_t*>(mapped[1] + offs[1]),
                         reinterpret_cast<const uint8_t*>(mapped[2] + offs[2]), sh,
                         sh + static_cast<size_t>(R) * c->list_cap};
  }
  CG_CUDA(cudaMemcpyAsync(c->d_peers, peers.data(), sizeof(PeerLayer) * R, cudaMemcpyHostToDevice, s));
  CG_CUDA(cudaStreamSynchronize(s));
  c->bound = partial;
  return CG_OK;
}

}  // namespace cg

using namespace cg;

extern "C" {

int32_t cg_comm_get_unique_id(uint8_t id[CG_COMM_ID_BYTES]) {
  if (!id) return CG_ERR_INVALID_ARG;
  int32_t rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId u;
  CG_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, CG_COMM_ID_BYTES);
  return CG_OK;
}

int32_t cg_comm_init(cg_context* ctx, const uint8_t id[CG_COMM_ID_BYTES], int32_t rank,
                     int32_t nranks) {
  if (!ctx || !id || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) {
    set_error("cg_comm_init: invalid argument");
    return CG_ERR_INVALID_ARG;
  }
  if (ctx->comm) {
    set_error("cg_comm_init: the context already has a communicator");
    return CG_ERR_INVALID_ARG;
  }
  int32_t rc = load_nccl();
  if (rc) return rc;
  CG_CUDA(cudaSetDevice(ctx->device));
  cg_comm* c = new cg_comm;
  c->rank = rank;
  c->nranks = nranks;
  ncclUniqueId u;
  memcpy(u.internal, id, CG_COMM_ID_BYTES);
  const int r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return CG_ERR_CUDA;
  }
  if (cudaMalloc(&c->d_token, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&c->d_peers, sizeof(PeerLayer) * nranks) != cudaSuccess) {
    set_error("cg_comm_init: out of device memory");
    return CG_ERR_CUDA;
  }
  cudaMemset(c->d_token, 0, sizeof(int));
  ctx->comm = c;
  return CG_OK;
}

int32_t cg_comm_destroy(cg_context* ctx) {
  if (!ctx) return CG_ERR_INVALID_ARG;
  cg_comm* c = ctx->comm;
  if (!c) return CG_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  unbind(c);
  if (c->d_token) cudaFree(c->d_token);
  if (c->d_peers) cudaFree(c->d_peers);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  ctx->comm = nullptr;
  return CG_OK;
}

int32_t cg_gather_global(const cg_layer* partial, cg_layer* owned, uint64_t* blocks_folded) {
  if (!partial || !owned || partial->ctx != owned->ctx || partial == owned) {
    set_error("cg_gather_global: partial and owned must be two layers of one context");
    return CG_ERR_INVALID_ARG;
  }
  cg_context* ctx = owned->ctx;
  cg_comm* c = ctx->comm;
  if (!c) {
    set_error("cg_gather_global: call cg_comm_init first");
    return CG_ERR_INVALID_ARG;
  }
  if (partial->v.voxel_size != owned->v.voxel_size) {
    set_error("cg_gather_global: the layers' voxel sizes differ");
    return CG_ERR_INVALID_ARG;
  }
  CG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  int32_t rc;
  if (c->bound != partial && (rc = bind_partial(ctx, partial))) return rc;
  const int R = c->nranks;
  uint32_t* counts = c->shared + static_cast<size_t>(R) * c->list_cap;
  const int n = static_cast<int>(partial->num_blocks);
  ctx->own_launches += 1 + R;
  CG_CUDA(cudaMemsetAsync(counts, 0, R * sizeof(uint32_t), s));
  if (n > 0)
    k_owner_lists<<<grid_for(n, 256), 256, 0, s>>>(partial->v, n, static_cast<uint32_t>(R),
                                                   static_cast<uint32_t>(c->list_cap), c->shared,
                                                   counts);
  // every rank's partial layer and lists are complete (stream-ordered, no host wait)
  if ((rc = barrier(ctx))) return rc;
  CG_CUDA(cudaMemsetAsync(&ctx->d_counters->blocks_out, 0, sizeof(unsigned long long), s));
  for (int src = 0; src < R; ++src)  // ascending source rank: the fold order is fixed
    k_fold_pull<<<ctx->num_sms * 8, 256, 0, s>>>(owned->v, c->d_peers, src, c->rank,
                                                 static_cast<uint32_t>(c->list_cap),
                                                 &ctx->d_counters->blocks_out);
  // nobody clears or refills its partial layer while a peer still reads it
  if ((rc = barrier(ctx))) return rc;
  CallCounters cc;
  rc = finish_call(owned, &cc);
  if (blocks_folded) *blocks_folded = cc.blocks_out;
  return rc;
}

int32_t cg_project_submaps_sharded(const cg_layer* const* submaps, const float* poses,
                                   size_t num_submaps, cg_layer* partial, cg_layer* owned,
                                   cg_merge_stats* stats) {
  if (!partial || !owned) return CG_ERR_INVALID_ARG;
  int32_t rc = cg_layer_clear(partial);
  if (rc) return rc;
  if (num_submaps) {
    rc = cg_project_submaps(submaps, poses, num_submaps, partial, stats);
    if (rc) return rc;
  } else if (stats) {
    memset(stats, 0, sizeof(*stats));
  }
  return cg_gather_global(partial, owned, nullptr);
}

}  // extern "C"
