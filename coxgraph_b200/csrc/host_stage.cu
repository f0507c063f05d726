// host_stage.cu — see host_stage.cuh.
#include "host_stage.cuh"

#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace cg {

constexpr size_t kStageChunk = size_t(512) << 10;    // bytes per chunk
// smaller copies go the direct way: waking the workers costs ~0.1 ms, more than they save on one
// 640x480 cloud (measured: 1.02 ms per pageable frame direct, 1.23 ms through the pool)
constexpr size_t kStageMinBytes = size_t(16) << 20;

struct HostStager {
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv;
  bool stop = false;
  unsigned long long generation = 0;  // bumped per job (under m)
  // the job
  char* bounce = nullptr;             // pinned
  size_t bounce_cap = 0;
  size_t used = 0;                    // bytes of the bounce buffer taken by earlier copies of the call
  const char* src = nullptr;
  char* dst_host = nullptr;
  size_t bytes = 0, chunks = 0;
  std::atomic<size_t> next{0};
  std::vector<std::atomic<unsigned char>> done;
  std::atomic<int> active{0};         // workers inside the current job

  void work() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        active.fetch_add(1, std::memory_order_acq_rel);
      }
      for (;;) {
        const size_t i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= chunks) break;
        const size_t off = i * kStageChunk;
        const size_t n = bytes - off < kStageChunk ? bytes - off : kStageChunk;
        memcpy(dst_host + off, src + off, n);
        done[i].store(1, std::memory_order_release);
      }
      active.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
};

static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory reports an error on older runtimes
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

cudaError_t stage_to_device(HostStager** stager, void* dst, const void* src, size_t bytes,
                            cudaStream_t stream, int threads) {
  if (bytes == 0) return cudaSuccess;
  if (threads <= 0 || bytes < kStageMinBytes || !is_pageable(src))
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
  if (!*stager) {
    *stager = new HostStager();
    for (int t = 0; t < threads; ++t) (*stager)->workers.emplace_back([s = *stager] { s->work(); });
  }
  HostStager& S = **stager;
  // several copies of one call (points, colours) share the bounce buffer: `used` is reset by the
  // caller through stage_reset() semantics below — here: a copy that does not fit behind the
  // earlier ones waits for the stream (its DMAs) and starts over
  if (S.used + bytes > S.bounce_cap) {
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    S.used = 0;
    if (bytes > S.bounce_cap) {
      if (S.bounce) cudaFreeHost(S.bounce);
      S.bounce = nullptr;
      S.bounce_cap = 0;
      const size_t want = bytes + bytes / 4 + (size_t(4) << 20);
      e = cudaHostAlloc(reinterpret_cast<void**>(&S.bounce), want, cudaHostAllocDefault);
      if (e != cudaSuccess) {  // no pinned memory to be had: the plain copy still works
        cudaGetLastError();
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
      }
      S.bounce_cap = want;
    }
  }
  const size_t chunks = (bytes + kStageChunk - 1) / kStageChunk;
  {
    // a worker that woke up late for the previous job may still be looking at its fields: they are
    // rewritten only while nobody is inside a job (workers enter one under this mutex)
    std::unique_lock<std::mutex> lk(S.m);
    while (S.active.load(std::memory_order_acquire) != 0) {
      lk.unlock();
      std::this_thread::yield();
      lk.lock();
    }
    if (S.done.size() < chunks) {
      std::vector<std::atomic<unsigned char>> fresh(chunks + chunks / 2);
      S.done.swap(fresh);
    }
    for (size_t i = 0; i < chunks; ++i) S.done[i].store(0, std::memory_order_relaxed);
    S.src = static_cast<const char*>(src);
    S.dst_host = S.bounce + S.used;
    S.bytes = bytes;
    S.chunks = chunks;
    S.next.store(0, std::memory_order_relaxed);
    ++S.generation;
  }
  S.cv.notify_all();
  // queue the DMA of every chunk as soon as it is in the bounce buffer; if the workers are slow
  // to wake, the calling thread copies chunks itself
  cudaError_t err = cudaSuccess;
  for (size_t i = 0; i < chunks; ++i) {
    while (!S.done[i].load(std::memory_order_acquire)) {
      const size_t j = S.next.fetch_add(1, std::memory_order_relaxed);
      if (j < chunks) {
        const size_t off = j * kStageChunk;
        const size_t n = bytes - off < kStageChunk ? bytes - off : kStageChunk;
        memcpy(S.dst_host + off, S.src + off, n);
        S.done[j].store(1, std::memory_order_release);
      } else {
        std::this_thread::yield();
      }
    }
    const size_t off = i * kStageChunk;
    const size_t n = bytes - off < kStageChunk ? bytes - off : kStageChunk;
    if (err == cudaSuccess)
      err = cudaMemcpyAsync(static_cast<char*>(dst) + off, S.dst_host + off, n,
                            cudaMemcpyHostToDevice, stream);
  }
  // no worker may still be inside this job when the next one rewrites its fields
  while (S.active.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  S.used += (bytes + 255) & ~size_t(255);
  return err;
}

cudaError_t stage_begin(HostStager* S, cudaStream_t stream) {
  if (!S || S->used == 0) return cudaSuccess;
  // the bounce buffer is reused from its start: whatever an earlier call queued from it must have
  // been read (a no-op after a call that completed, which synchronises the stream)
  const cudaError_t e = cudaStreamSynchronize(stream);
  S->used = 0;
  return e;
}

void destroy_stager(HostStager* S) {
  if (!S) return;
  {
    std::lock_guard<std::mutex> lk(S->m);
    S->stop = true;
  }
  S->cv.notify_all();
  for (std::thread& t : S->workers) t.join();
  if (S->bounce) cudaFreeHost(S->bounce);
  delete S;
}

}  // namespace cg
