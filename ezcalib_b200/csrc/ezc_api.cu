// ezc_api.cu -- context management and error plumbing of the C ABI (include/ezcalib_b200.h).
#include "ezc_common.cuh"

namespace ezc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace ezc

extern "C" {

const char* ezc_last_error(void) { return ezc::g_err; }

ezc_status ezc_create(int32_t device, ezc_context** out) {
  EZC_CHECK(out != nullptr, EZC_ERR_INVALID, "ezc_create: out is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    ezc::set_error("ezc_create: no CUDA device available (%s); this library has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return EZC_ERR_CUDA;
  }
  EZC_CHECK(device >= 0 && device < n, EZC_ERR_INVALID, "ezc_create: device %d out of range [0,%d)", device, n);
  EZC_CUDA(cudaSetDevice(device));
  ezc_context* c = new ezc_context();
  c->device = device;
  cudaDeviceProp prop;
  EZC_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  EZC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  {
    std::lock_guard<std::mutex> g(ezc::dev_cache(device).m);
    ++ezc::live_contexts(device);
  }
  *out = c;
  return EZC_OK;
}

ezc_status ezc_trim_cache(int32_t device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return EZC_OK; }
  for (int d = 0; d < n && d < 64; ++d) {
    if (device >= 0 && d != device) continue;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d);
    cudaDeviceSynchronize();
    ezc::dev_cache(d).flush();
    cudaSetDevice(prev);
  }
  return EZC_OK;
}

void ezc_destroy(ezc_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);  // cached device blocks are handed to other contexts unsynchronised
  if (ctx->comm && ctx->comm_free) ctx->comm_free(ctx->comm);
  if (ctx->at_pools && ctx->at_pools_free) ctx->at_pools_free(ctx->at_pools);
  if (ctx->cb_pools && ctx->cb_pools_free) ctx->cb_pools_free(ctx->cb_pools);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  bool last = false;
  {
    std::lock_guard<std::mutex> g(ezc::dev_cache(ctx->device).m);
    last = --ezc::live_contexts(ctx->device) <= 0;
  }
  if (last) ezc::dev_cache(ctx->device).flush();  // other allocators in the process (torch, NCCL) get the memory back
  delete ctx;
}

ezc_status ezc_set_stream(ezc_context* ctx, void* cuda_stream) {
  EZC_CHECK(ctx != nullptr, EZC_ERR_INVALID, "ezc_set_stream: ctx is null");
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return EZC_OK;
}

ezc_status ezc_set_allreduce(ezc_context* ctx, ezc_allreduce_fn fn, void* user, int32_t rank, int32_t world) {
  EZC_CHECK(ctx != nullptr, EZC_ERR_INVALID, "ezc_set_allreduce: ctx is null");
  EZC_CHECK(world >= 1 && rank >= 0 && rank < world, EZC_ERR_INVALID, "ezc_set_allreduce: bad rank/world %d/%d", rank, world);
  EZC_CHECK(world == 1 || fn != nullptr, EZC_ERR_INVALID, "ezc_set_allreduce: world > 1 needs a callback");
  ctx->allreduce = fn;
  ctx->allreduce_user = user;
  ctx->rank = rank;
  ctx->world = world;
  return EZC_OK;
}

uint64_t ezc_kernel_launches(const ezc_context* ctx) { return ctx ? ctx->kernel_launches : 0; }

ezc_status ezc_synchronize(ezc_context* ctx) {
  EZC_CHECK(ctx != nullptr, EZC_ERR_INVALID, "ezc_synchronize: ctx is null");
  EZC_CUDA(cudaSetDevice(ctx->device));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}

}  // extern "C"
