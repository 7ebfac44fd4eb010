// microbench.cu -- roofline denominators measured on the device the bench runs on.
// MEASURED_PEAKS.json (driver-written) records HBM copy bandwidth and bf16 tensor peaks but no
// FP64 figure; the LM build is FP64-FMA bound, so its peak is measured here with
// a DFMA-saturating kernel (8 independent chains per thread, all SMs resident).
#include <algorithm>
#include "ezc_common.cuh"

namespace ezc {

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}


// FP64 tensor-core (DMMA m8n8k4) throughput: 8 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) { c[t][0] = seed + t; c[t][1] = seed - t; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace ezc

extern "C" ezc_status ezc_measure_fp64_peak(ezc_context* ctx, double* tflops_out) {
  EZC_CHECK(ctx && tflops_out, EZC_ERR_INVALID, "ezc_measure_fp64_peak: null argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  ezc::DevBuf buf;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  EZC_CUDA(buf.reserve(64));
  const int iters = 4096, grid = ctx->sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  EZC_CUDA(cudaEventCreate(&e0));
  EZC_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    EZC_CUDA(cudaEventRecord(e0, ctx->stream));
    ezc::fp64_peak_kernel<<<grid, block, 0, ctx->stream>>>(buf.as<double>(), iters, 1.0 + rep);
    EZC_CUDA(cudaEventRecord(e1, ctx->stream));
    EZC_CUDA(cudaEventSynchronize(e1));
    ctx->kernel_launches++;
    float ms = 0.f;
    EZC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * (double)grid * block;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return EZC_OK;
}

extern "C" ezc_status ezc_measure_dmma_peak(ezc_context* ctx, double* tflops_out) {
  EZC_CHECK(ctx && tflops_out, EZC_ERR_INVALID, "ezc_measure_dmma_peak: null argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  ezc::DevBuf buf;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  EZC_CUDA(buf.reserve(64));
  const int iters = 4096, grid = ctx->sm_count * 8, block = 256;
  cudaEvent_t e0, e1;
  EZC_CUDA(cudaEventCreate(&e0));
  EZC_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    EZC_CUDA(cudaEventRecord(e0, ctx->stream));
    ezc::dmma_peak_kernel<<<grid, block, 0, ctx->stream>>>(buf.as<double>(), iters, 1.0 + rep);
    EZC_CUDA(cudaEventRecord(e1, ctx->stream));
    EZC_CUDA(cudaEventSynchronize(e1));
    ctx->kernel_launches++;
    float ms = 0.f;
    EZC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 MMA = 8*8*4 FMA = 512 FLOP per warp
    const double flops = 512.0 * 8.0 * iters * (double)grid * (block / 32);
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return EZC_OK;
}
