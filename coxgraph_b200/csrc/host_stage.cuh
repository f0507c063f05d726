// host_stage.cuh — pageable host buffers into the device at more than one core's memcpy rate.
//
// The drop-in call (TsdfIntegratorBase::integratePointCloud with std::vector clouds,
// coxgraph/include/coxgraph/map_comm/tsdf_recover.h:75) hands the library PAGEABLE memory.  A
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread (≈ 10 GB/s: a
// 640x480 cloud costs as much as fusing it).  Here a few worker threads copy the buffer chunk by
// chunk into a pinned bounce buffer while the calling thread queues the DMA of every finished
// chunk on the context's stream.  Pinned (or registered) inputs take the direct copy.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace cg {

struct HostStager;  // owned by the context; created on first use

// queues dst[0, bytes) <- src on `stream` (dst: device memory); returns after every chunk's DMA has
// been QUEUED (the bounce buffer is reused by the next call: callers synchronise the stream before
// that, as every integrate call does).  threads <= 0: plain cudaMemcpyAsync.
cudaError_t stage_to_device(HostStager** stager, void* dst, const void* src, size_t bytes,
                            cudaStream_t stream, int threads);
// start of a call that stages inputs: the bounce buffer is used from its beginning again
cudaError_t stage_begin(HostStager* stager, cudaStream_t stream);
void destroy_stager(HostStager* stager);

}  // namespace cg
