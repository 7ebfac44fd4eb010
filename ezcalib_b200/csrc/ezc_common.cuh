// ezc_common.cuh -- context, error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ezcalib_b200.h"

namespace ezc {

void set_error(const char* fmt, ...);

#define EZC_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::ezc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
      return EZC_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define EZC_CHECK(cond, code, ...)          \
  do {                                      \
    if (!(cond)) {                          \
      ::ezc::set_error(__VA_ARGS__);        \
      return (code);                        \
    }                                       \
  } while (0)

// Simple RAII device buffer (cudaMalloc / cudaFree), grow-only.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostPinned {
  void* p = nullptr;
  size_t cap = 0;
  ~HostPinned() { if (p) cudaFreeHost(p); }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace ezc

// The opaque context of the C ABI.
struct ezc_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  // multi-GPU plumbing: host callback that all-reduces a DEVICE buffer of doubles in place
  // (torch.distributed over NCCL in production, gloo/CPU in the tests).
  ezc_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  int rank = 0;
  int world = 1;
  uint64_t kernel_launches = 0;
  // per-context scratch pools of the detection pipelines (kept between calls; freed with the context)
  void* at_pools = nullptr;
  void (*at_pools_free)(void*) = nullptr;
  void* cb_pools = nullptr;
  void (*cb_pools_free)(void*) = nullptr;
};
