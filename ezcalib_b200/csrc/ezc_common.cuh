// ezc_common.cuh -- context, error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
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
// Process-wide cache of device allocations: cudaMalloc/cudaFree cost milliseconds (cudaFree also synchronises the
// device), which dominated the host-buffer (e2e) calls that create and destroy a solver / image batch per call.
// Blocks are handed back by ~DevBuf and reused best-fit (<= 25 % slack) by the next reserve().  Reuse is safe without a
// synchronise as long as producer and consumer run on the same stream (one context = one stream); entry points that
// destroy objects synchronise their stream first.  Never freed at process exit (the CUDA context may already be gone).
struct DevCache {
  std::mutex m;
  std::multimap<size_t, void*> blocks;
  size_t bytes = 0;
  size_t limit = size_t(24) << 30;
  void* take(size_t want) {
    std::lock_guard<std::mutex> g(m);
    auto it = blocks.lower_bound(want);
    if (it == blocks.end() || it->first > want + want / 4 + 4096) return nullptr;
    void* p = it->second;
    bytes -= it->first;
    last_size = it->first;
    blocks.erase(it);
    return p;
  }
  size_t last_size = 0;
  bool give(void* p, size_t sz) {
    std::lock_guard<std::mutex> g(m);
    if (bytes + sz > limit) return false;
    blocks.emplace(sz, p);
    bytes += sz;
    return true;
  }
  void flush() {
    std::lock_guard<std::mutex> g(m);
    for (auto& kv : blocks) cudaFree(kv.second);
    blocks.clear();
    bytes = 0;
  }
};
// One cache per device: a block cudaMalloc'ed on GPU a must never be handed to a context on GPU b.
inline DevCache& dev_cache(int device) {
  static DevCache* caches[64] = {};  // intentionally leaked
  static std::mutex m;
  std::lock_guard<std::mutex> g(m);
  if (device < 0 || device >= 64) device = 0;
  if (!caches[device]) caches[device] = new DevCache;
  return *caches[device];
}
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
  return d;
}
inline DevCache& dev_cache() { return dev_cache(current_device()); }
inline size_t dev_cache_bytes(int device) {
  DevCache& c = dev_cache(device);
  std::lock_guard<std::mutex> g(c.m);
  return c.bytes;
}
// number of live contexts per device; the last ezc_destroy on a device flushes its cache
inline int& live_contexts(int device) {
  static int n[64] = {};
  return n[(device < 0 || device >= 64) ? 0 : device];
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int dev = -1;  // device the block lives on
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && !dev_cache(dev).give(p, cap)) cudaFree(p);
    p = nullptr; cap = 0;
  }
  // Error paths destroy DevBufs while kernels may still be queued on `stream`: wait before the block can be recycled.
  void release_after(cudaStream_t stream) {
    if (p) cudaStreamSynchronize(stream);
    release();
  }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    dev = current_device();
    DevCache& c = dev_cache(dev);
    {
      std::lock_guard<std::mutex> g(c.m);
      auto it = c.blocks.lower_bound(bytes);
      if (it != c.blocks.end() && it->first <= bytes + bytes / 4 + 4096) {
        p = it->second; cap = it->first;
        c.bytes -= it->first;
        c.blocks.erase(it);
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and retry once
      cudaGetLastError();
      c.flush();
      e = cudaMalloc(&p, bytes);
    }
    if (e == cudaSuccess) cap = bytes; else p = nullptr;
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Declared AFTER the local DevBufs of an entry point: runs before their destructors on every return path (error returns
// included), so a block is never recycled while kernels that use it are still queued on the stream.
struct SyncOnExit {
  cudaStream_t s;
  explicit SyncOnExit(cudaStream_t stream) : s(stream) {}
  ~SyncOnExit() { cudaStreamSynchronize(s); }
};

// Pinned host buffers are expensive to create (cudaMallocHost / cudaFreeHost take milliseconds -- more than the 80 MB
// observation upload of a solver): released buffers are kept, like the device blocks above.
struct HostCache {
  std::mutex m;
  std::multimap<size_t, void*> blocks;
  size_t bytes = 0;
  size_t limit = size_t(256) << 20;
  void* take(size_t want, size_t& got) {
    std::lock_guard<std::mutex> g(m);
    auto it = blocks.lower_bound(want);
    if (it == blocks.end() || it->first > 2 * want + 4096) return nullptr;
    void* p = it->second;
    got = it->first;
    bytes -= it->first;
    blocks.erase(it);
    return p;
  }
  bool give(void* p, size_t sz) {
    std::lock_guard<std::mutex> g(m);
    if (bytes + sz > limit) return false;
    blocks.emplace(sz, p);
    bytes += sz;
    return true;
  }
};
inline HostCache& host_cache() {
  static HostCache* c = new HostCache;  // intentionally leaked
  return *c;
}
struct HostPinned {
  void* p = nullptr;
  size_t cap = 0;
  HostPinned() = default;
  HostPinned(const HostPinned&) = delete;
  HostPinned& operator=(const HostPinned&) = delete;
  ~HostPinned() { release(); }
  void release() {
    if (p && !host_cache().give(p, cap)) cudaFreeHost(p);
    p = nullptr; cap = 0;
  }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    size_t got = 0;
    if (void* q = host_cache().take(bytes, got)) { p = q; cap = got; return cudaSuccess; }
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes; else p = nullptr;
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
  // second stream for ezc_images_upload_async: H2D copies overlap the detection kernels of earlier sub-batches
  cudaStream_t copy_stream = nullptr;
  // library-owned peer-memory exchange (csrc/comm.cu)
  void* comm = nullptr;
  void (*comm_free)(void*) = nullptr;
};
