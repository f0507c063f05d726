// comm.cu — the server's global merge across the GPUs of one box, behind the C ABI (SURVEY.md
// §8b cg_comm_init / cg_gather_global, §8e).
//
// Every rank projects its submaps into a *partial* global layer.  The exchange is an owner-pull
// over NVLink peer memory (peer pointers opened through CUDA IPC):
//   k_build_holders  every rank reads the block keys of every rank's partial layer (8 B per block)
//                    and builds the same map  block -> set of ranks that hold it
//   k_list_owned     the owner of a block is one of its HOLDERS, picked by a hash of the index: a
//                    block only one rank holds never crosses NVLink, contested blocks spread
//                    evenly over their holders
//   k_fold_owned     one CTA per owned block: the holders' copies are read straight out of their
//                    block pools, in ascending rank order, and folded in shared memory with
//                    voxblox::mergeLayerAintoLayerB's aligned form (Block::mergeBlock /
//                    mergeVoxelAIntoVoxelB, R10; reference call site of that overload:
//                    coxgraph/src/server/submap_collection.cpp:31-33); the block is written once
// — gather and fold are ONE kernel, there is no packing pass, no staging copy and no host
// synchronisation between listing, exchange and fold.  NCCL (loaded at run time from
// libnccl.so.2, the one torch has loaded if there is one) is the bootstrap and the barrier only:
// unique id -> communicator, one all-gather of the IPC handles when a partial layer is bound, and a
// one-word all-reduce on the stream before and after the kernels (all partial layers complete /
// all peers done reading).  The fold order is fixed, so the result is deterministic and equal to
// the packed exchange of exchange.cu (hash owner, NCCL all-to-all) block for block.
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
  const uint32_t* shared;  // [0] number of blocks of the partial layer (written per call)
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
  uint32_t* shared = nullptr;         // peer-visible: [0] blocks of this rank's partial layer
  std::vector<void*> opened;          // peer mappings to close
  cg::PeerLayer* d_peers = nullptr;   // [nranks]
  // the holder map (scratch, rebuilt every call): block key -> set of ranks holding the block
  unsigned long long* hkeys = nullptr;
  unsigned long long* hmask = nullptr;
  uint32_t* hbase = nullptr;          // owned entries: first element of the block's slot list
  uint32_t* mine = nullptr;           // entries this rank owns
  uint32_t* slots = nullptr;          // pool slots of the holders' copies, per owned block
  uint32_t* counters = nullptr;       // [0] owned entries, [1] slot list cursor, [2] overflow
  size_t hcap = 0, list_cap = 0;
};

namespace cg {

__device__ __forceinline__ uint32_t holder_find(const unsigned long long* __restrict__ hkeys,
                                                uint32_t hmask_cap, unsigned long long key) {
  uint32_t h = hash_key(key) & hmask_cap;
  while (hkeys[h] != key) h = (h + 1) & hmask_cap;
  return h;
}
// the owner of a block: one of its holders, picked by a hash of the block index
__device__ __forceinline__ int pick_owner(unsigned long long key, unsigned long long mask) {
  int n = static_cast<int>((hash_key(key ^ 0x9E3779B97F4A7C15ULL) >> 7) %
                           static_cast<uint32_t>(__popcll(mask)));
  unsigned long long m = mask;
  while (n-- > 0) m &= m - 1;
  return __ffsll(static_cast<long long>(m)) - 1;
}

// Every block key of every rank's partial layer (read through the peer mappings) enters the map
// with its rank's bit.  Grid y = rank.
__global__ void k_build_holders(const PeerLayer* __restrict__ peers, unsigned long long* hkeys,
                                unsigned long long* hmask, uint32_t hcap_mask, uint32_t* counters) {
  const int r = blockIdx.y;
  const PeerLayer P = peers[r];
  const uint32_t n = P.shared[0];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long key = P.block_keys[i];
    uint32_t h = hash_key(key) & hcap_mask;
    for (uint32_t tries = 0;; ++tries) {
      const unsigned long long k = hkeys[h];
      if (k == key) break;
      if (k == kEmptyKey) {
        const unsigned long long old = atomicCAS(&hkeys[h], kEmptyKey, key);
        if (old == kEmptyKey || old == key) break;
      }
      h = (h + 1) & hcap_mask;
      if (tries > hcap_mask) {  // cannot happen: the map holds every block of every rank
        atomicExch(&counters[2], 1u);
        return;
      }
    }
    atomicOr(&hmask[h], 1ull << r);
  }
}
// entries this rank owns, each with room for its holders' slots
__global__ void k_list_owned(const unsigned long long* __restrict__ hkeys,
                             const unsigned long long* __restrict__ hmask, uint32_t hcap, int me,
                             uint32_t* __restrict__ hbase, uint32_t* __restrict__ mine,
                             uint32_t list_cap, uint32_t* counters) {
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < hcap; e += gridDim.x * blockDim.x) {
    const unsigned long long mask = hmask[e];
    if (!mask || pick_owner(hkeys[e], mask) != me) continue;
    const uint32_t pos = atomicAdd(&counters[0], 1u);
    const uint32_t base = atomicAdd(&counters[1], static_cast<uint32_t>(__popcll(mask)));
    if (pos < list_cap) {
      mine[pos] = e;
      hbase[e] = base;
    } else {
      atomicExch(&counters[2], 1u);
    }
  }
}
// the pool slot of every holder's copy of the blocks this rank owns.  Grid y = rank.
__global__ void k_holder_slots(const PeerLayer* __restrict__ peers,
                               const unsigned long long* __restrict__ hkeys,
                               const unsigned long long* __restrict__ hmask,
                               const uint32_t* __restrict__ hbase, uint32_t hcap_mask, int me,
                               uint32_t* __restrict__ slots, uint32_t slots_cap) {
  const int r = blockIdx.y;
  const PeerLayer P = peers[r];
  const uint32_t n = P.shared[0];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long key = P.block_keys[i];
    const uint32_t e = holder_find(hkeys, hcap_mask, key);
    const unsigned long long mask = hmask[e];
    if (pick_owner(key, mask) != me) continue;
    const uint32_t at = hbase[e] + __popcll(mask & ((1ull << r) - 1ull));
    if (at < slots_cap) slots[at] = i;
  }
}

// One CTA per owned block: the destination block (fresh blocks are in the default state) is staged
// in shared memory, every holder's copy is read through its mapping — 16-byte loads, across
// NVLink for the peers — and folded in ascending rank order, and the result is written once.
constexpr int kFoldThreads = 256;
__global__ void __launch_bounds__(kFoldThreads)
k_fold_owned(LayerView B, const PeerLayer* __restrict__ peers,
             const unsigned long long* __restrict__ hkeys,
             const unsigned long long* __restrict__ hmask, const uint32_t* __restrict__ hbase,
             const uint32_t* __restrict__ mine, const uint32_t* __restrict__ slots,
             const uint32_t* __restrict__ counters, uint32_t list_cap, int blocks_before,
             unsigned long long* folded) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* t_d = reinterpret_cast<uint4*>(smem_raw);  // three planes of 4096 words
  uint4* t_w = t_d + kVoxelsPerBlock / 4;
  uint4* t_c = t_w + kVoxelsPerBlock / 4;
  __shared__ int s_slot, s_fresh;
  const uint32_t n = min(counters[0], list_cap);
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t e = mine[i];
    const unsigned long long key = hkeys[e];
    unsigned long long mask = hmask[e];
    const uint32_t base = hbase[e];
    // Block::mergeBlock skips a source block without data: is there one with data at all?
    bool any = false;
    {
      unsigned long long m = mask;
      for (int j = 0; m; ++j, m &= m - 1)
        any = any || peers[__ffsll(static_cast<long long>(m)) - 1].has_data[slots[base + j]] != 0;
    }
    __syncthreads();  // the previous block's tile has been written out
    if (!any) continue;
    if (threadIdx.x == 0) {
      const int he = B.insert_entry(key);
      s_slot = B.hash_vals[he];
      s_fresh = 0;
      if (s_slot >= 0) {
        // a slot claimed during this call holds a default-constructed block (the pool keeps
        // unclaimed slots in that state): nothing to read
        s_fresh = s_slot >= blocks_before;
        B.has_data[s_slot] = 1;
        B.updated[s_slot] = 1;
        atomicAdd(folded, static_cast<unsigned long long>(__popcll(mask)));
      }
    }
    __syncthreads();
    const int slot = s_slot;
    if (slot < 0) continue;  // pool exhausted: flagged by insert_entry
    uint4* dd = reinterpret_cast<uint4*>(B.dist_plane(slot));
    uint4* dw = reinterpret_cast<uint4*>(B.weight_plane(slot));
    uint4* dc = reinterpret_cast<uint4*>(B.color_plane(slot));
    if (s_fresh) {
      for (int q = threadIdx.x; q < kVoxelsPerBlock / 4; q += kFoldThreads) {
        t_d[q] = make_uint4(0u, 0u, 0u, 0u);
        t_w[q] = make_uint4(0u, 0u, 0u, 0u);
        t_c[q] = make_uint4(kDefaultColor, kDefaultColor, kDefaultColor, kDefaultColor);
      }
    } else {
      for (int q = threadIdx.x; q < kVoxelsPerBlock / 4; q += kFoldThreads) {
        t_d[q] = dd[q];
        t_w[q] = dw[q];
        t_c[q] = dc[q];
      }
    }
    for (int j = 0; mask; ++j, mask &= mask - 1) {  // ascending rank: the fold order is fixed
      const PeerLayer P = peers[__ffsll(static_cast<long long>(mask)) - 1];
      const uint32_t a = slots[base + j];
      if (!P.has_data[a]) continue;
      const uint4* sd =
          reinterpret_cast<const uint4*>(P.pool + static_cast<size_t>(a) * (3 * kVoxelsPerBlock));
      const uint4* sw = sd + kVoxelsPerBlock / 4;
      const uint4* sc = sw + kVoxelsPerBlock / 4;
      // each thread folds the same quads in every round: no synchronisation between sources
      for (int q = threadIdx.x; q < kVoxelsPerBlock / 4; q += kFoldThreads) {
        const uint4 ad = sd[q], aw = sw[q], ac = sc[q];
        const uint4 bd = t_d[q], bw = t_w[q], bc = t_c[q];
        VoxelState v0{__uint_as_float(bd.x), __uint_as_float(bw.x), bc.x};
        VoxelState v1{__uint_as_float(bd.y), __uint_as_float(bw.y), bc.y};
        VoxelState v2{__uint_as_float(bd.z), __uint_as_float(bw.z), bc.z};
        VoxelState v3{__uint_as_float(bd.w), __uint_as_float(bw.w), bc.w};
        merge_voxel(__uint_as_float(ad.x), __uint_as_float(aw.x), ac.x, v0);
        merge_voxel(__uint_as_float(ad.y), __uint_as_float(aw.y), ac.y, v1);
        merge_voxel(__uint_as_float(ad.z), __uint_as_float(aw.z), ac.z, v2);
        merge_voxel(__uint_as_float(ad.w), __uint_as_float(aw.w), ac.w, v3);
        t_d[q] = make_uint4(__float_as_uint(v0.d), __float_as_uint(v1.d), __float_as_uint(v2.d),
                            __float_as_uint(v3.d));
        t_w[q] = make_uint4(__float_as_uint(v0.w), __float_as_uint(v1.w), __float_as_uint(v2.w),
                            __float_as_uint(v3.w));
        t_c[q] = make_uint4(v0.c, v1.c, v2.c, v3.c);
      }
    }
    for (int q = threadIdx.x; q < kVoxelsPerBlock / 4; q += kFoldThreads) {
      dd[q] = t_d[q];
      dw[q] = t_w[q];
      dc[q] = t_c[q];
    }
  }
}

__global__ void k_publish_blocks(uint32_t* shared, int n) { shared[0] = static_cast<uint32_t>(n); }
__global__ void k_comm_overflow(const uint32_t* counters, int32_t* err) {
  if (counters[2]) atomicOr(err, kErrPoolFull);
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
  void* bufs[] = {c->shared, c->hkeys, c->hmask, c->hbase, c->mine, c->slots, c->counters};
  for (void* b : bufs)
    if (b) cudaFree(b);
  c->shared = c->hbase = c->mine = c->slots = c->counters = nullptr;
  c->hkeys = c->hmask = nullptr;
  c->bound = nullptr;
}

static int32_t barrier(cg_context* ctx) {
  cg_comm* c = ctx->comm;
  CG_NCCL(g_nccl.AllReduce(c->d_token, c->d_token, 1, ncclInt32, ncclSum, c->comm, ctx->stream));
  return CG_OK;
}

// All ranks call this with their own partial layer (collective): IPC handles exchanged, peer
// mappings opened, scratch of the holder map sized for every block of every rank.  Done once per
// partial layer (it stays bound until another one is used).
static int32_t bind_partial(cg_context* ctx, const cg_layer* partial) {
  cg_comm* c = ctx->comm;
  unbind(c);
  cudaStream_t s = ctx->stream;
  const int R = c->nranks;
  c->list_cap = partial->max_blocks * static_cast<size_t>(R);  // blocks over all ranks
  c->hcap = 1024;
  while (c->hcap < 2 * c->list_cap) c->hcap <<= 1;
  CG_CUDA(cudaMalloc(&c->shared, 64));
  CG_CUDA(cudaMalloc(&c->hkeys, c->hcap * sizeof(unsigned long long)));
  CG_CUDA(cudaMalloc(&c->hmask, c->hcap * sizeof(unsigned long long)));
  CG_CUDA(cudaMalloc(&c->hbase, c->hcap * sizeof(uint32_t)));
  CG_CUDA(cudaMalloc(&c->mine, c->list_cap * sizeof(uint32_t)));
  CG_CUDA(cudaMalloc(&c->slots, c->list_cap * sizeof(uint32_t)));
  CG_CUDA(cudaMalloc(&c->counters, 4 * sizeof(uint32_t)));
  CG_CUDA(cudaMemsetAsync(c->shared, 0, 64, s));
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
      peers[r] = PeerLayer{partial->v.pool, partial->v.block_keys, partial->v.has_data, c->shared};
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
    peers[r] = PeerLayer{reinterpret_cast<const float*>(mapped[0] + offs[0]),
                         reinterpret_cast<const uint64_t*>(mapped[1] + offs[1]),
                         reinterpret_cast<const uint8_t*>(mapped[2] + offs[2]),
                         reinterpret_cast<const uint32_t*>(mapped[3] + offs[3])};
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
  // CG_TRACE_COMM=1: per-phase device times of the exchange on stderr (diagnostics only)
  static const bool trace = getenv("CG_TRACE_COMM") != nullptr;
  cudaEvent_t tev[6] = {};
  int tn = 0;
  auto mark = [&]() {
    if (trace && tn < 6) {
      cudaEventCreate(&tev[tn]);
      cudaEventRecord(tev[tn++], s);
    }
  };
  mark();
  const int R = c->nranks;
  const uint32_t hmask_cap = static_cast<uint32_t>(c->hcap - 1);
  const uint32_t list_cap = static_cast<uint32_t>(std::min<size_t>(c->list_cap, 0xFFFFFFF0u));
  ctx->own_launches += 6;
  k_publish_blocks<<<1, 1, 0, s>>>(c->shared, static_cast<int>(partial->num_blocks));
  CG_CUDA(cudaMemsetAsync(c->hkeys, 0xFF, c->hcap * sizeof(unsigned long long), s));
  CG_CUDA(cudaMemsetAsync(c->hmask, 0, c->hcap * sizeof(unsigned long long), s));
  CG_CUDA(cudaMemsetAsync(c->counters, 0, 4 * sizeof(uint32_t), s));
  CG_CUDA(cudaMemsetAsync(&ctx->d_counters->blocks_out, 0, sizeof(unsigned long long), s));
  // every rank's partial layer is complete and its block count published (stream-ordered)
  if ((rc = barrier(ctx))) return rc;
  mark();
  const dim3 per_rank(static_cast<unsigned>(ctx->num_sms), static_cast<unsigned>(R));
  k_build_holders<<<per_rank, 256, 0, s>>>(c->d_peers, c->hkeys, c->hmask, hmask_cap, c->counters);
  k_list_owned<<<ctx->num_sms * 4, 256, 0, s>>>(c->hkeys, c->hmask, static_cast<uint32_t>(c->hcap),
                                                c->rank, c->hbase, c->mine, list_cap, c->counters);
  k_holder_slots<<<per_rank, 256, 0, s>>>(c->d_peers, c->hkeys, c->hmask, c->hbase, hmask_cap,
                                          c->rank, c->slots, list_cap);
  mark();
  const size_t smem = 3 * kVoxelsPerBlock * sizeof(uint32_t);
  static bool attr_set[64] = {};
  if (ctx->device < 0 || ctx->device >= 64 || !attr_set[ctx->device]) {
    CG_CUDA(cudaFuncSetAttribute(k_fold_owned, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    if (ctx->device >= 0 && ctx->device < 64) attr_set[ctx->device] = true;
  }
  k_fold_owned<<<ctx->num_sms * 4, kFoldThreads, smem, s>>>(
      owned->v, c->d_peers, c->hkeys, c->hmask, c->hbase, c->mine, c->slots, c->counters, list_cap,
      static_cast<int>(owned->num_blocks), &ctx->d_counters->blocks_out);
  k_comm_overflow<<<1, 1, 0, s>>>(c->counters, owned->v.err);
  mark();
  // nobody clears or refills its partial layer while a peer still reads it
  if ((rc = barrier(ctx))) return rc;
  mark();
  CallCounters cc;
  rc = finish_call(owned, &cc);
  if (trace && tn == 5) {
    float t[4];
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
    fprintf(stderr, "[cg comm rank %d] reset+barrier %.3f  holders+list+slots %.3f  fold %.3f  barrier %.3f ms "
            "(partial %lld blocks, hash %zu)\n", c->rank, t[0], t[1], t[2], t[3],
            static_cast<long long>(partial->num_blocks), c->hcap);
  }
  for (int i = 0; i < tn; ++i) cudaEventDestroy(tev[i]);
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
