// CPU shim of the few CUDA runtime calls host_stage.cu uses (test harness only)
#pragma once
#include <stdlib.h>
#include <string.h>
typedef int cudaError_t; typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1 };
enum { cudaHostAllocDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1 };
struct cudaPointerAttributes { int type; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeUnregistered; return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { *p = malloc(n); return *p ? 0 : 1; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
