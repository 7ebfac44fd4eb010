// scan_sort.cuh -- hand-written device primitives for the detection pipelines:
//   * exclusive prefix sum over int32 (3 kernels, fixed order => deterministic),
//   * STABLE least-significant-digit radix sort of a uint32 payload array whose 8-bit digit is computed
//     from the payload by a functor (stability is what reproduces std::stable_sort / raster order of
//     the reference: TagDetector.cc:237, std::map iteration TagDetector.cc:245-261).
#pragma once
#include "ezc_common.cuh"

namespace ezc {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

static __global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const int* __restrict__ in, long long n, int* __restrict__ sums) {
  __shared__ int sh[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile;
  int v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const long long i = base + (long long)threadIdx.x * kScanItems + k;
    if (i < n) v += in[i];
  }
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) s += sh[w];
    sums[blockIdx.x] = s;
  }
}

// single block: in-place exclusive scan of sums[nb]; writes the grand total to *total (if not null)
static __global__ void __launch_bounds__(1024) scan_sums_kernel(int* __restrict__ sums, int nb, int* __restrict__ total) {
  __shared__ int sh[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int inc = v;
    for (int m = 1; m < 32; m <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, m); if ((threadIdx.x & 31) >= m) inc += o; }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = sh[threadIdx.x];
      for (int m = 1; m < 32; m <<= 1) { const int o = __shfl_up_sync(0xffffffffu, w, m); if (threadIdx.x >= m) w += o; }
      sh[threadIdx.x] = w;
    }
    __syncthreads();
    const int wp = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
    const int c = carry;
    if (i < nb) sums[i] = c + wp + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + wp + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

static __global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, long long n,
                                                                   const int* __restrict__ sums) {
  __shared__ int sh[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int t = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < n) ? in[base + k] : 0; t += v[k]; }
  int inc = t;
  for (int m = 1; m < 32; m <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, m); if ((threadIdx.x & 31) >= m) inc += o; }
  if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
  __syncthreads();
  int wp = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wp += sh[w];
  int run = sums[blockIdx.x] + wp + inc - t;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

// out[i] = sum_{j<i} in[j]; out may alias in.  total (device int*, optional) receives the grand total.
inline ezc_status exclusive_scan_i32(ezc_context* ctx, const int* in, int* out, long long n, DevBuf& tmp, int* d_total) {
  if (n <= 0) {
    if (d_total) EZC_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), ctx->stream));
    return EZC_OK;
  }
  const int nb = (int)((n + kScanTile - 1) / kScanTile);
  EZC_CUDA(tmp.reserve(sizeof(int) * (size_t)nb));
  scan_block_sums_kernel<<<nb, kScanThreads, 0, ctx->stream>>>(in, n, tmp.as<int>());
  scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(tmp.as<int>(), nb, d_total);
  scan_apply_kernel<<<nb, kScanThreads, 0, ctx->stream>>>(in, out, n, tmp.as<int>());
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches += 3;
  return EZC_OK;
}

// ------------------------------------------------------------------------------------------ radix sort
constexpr int kSortWarps = 8;
constexpr int kSortChunk = 1024;  // items per warp per pass

template <class DigitFn>
__global__ void __launch_bounds__(kSortWarps * 32) radix_hist_kernel(const uint32_t* __restrict__ vals, int n, int n_warps,
                                                                       DigitFn digit, int* __restrict__ counts) {
  __shared__ int h[kSortWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kSortWarps + warp;
  for (int b = lane; b < 256; b += 32) h[warp][b] = 0;
  __syncwarp();
  if (gw < n_warps) {
    const int beg = gw * kSortChunk, end = min(n, beg + kSortChunk);
    for (int i = beg + lane; i < end; i += 32) atomicAdd(&h[warp][digit(vals[i])], 1);
    __syncwarp();
    for (int b = lane; b < 256; b += 32) counts[(size_t)b * n_warps + gw] = h[warp][b];
  }
}

template <class DigitFn>
__global__ void __launch_bounds__(kSortWarps * 32) radix_scatter_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ out,
                                                                          int n, int n_warps, DigitFn digit,
                                                                          const int* __restrict__ offsets) {
  __shared__ int off[kSortWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kSortWarps + warp;
  if (gw >= n_warps) return;
  for (int b = lane; b < 256; b += 32) off[warp][b] = offsets[(size_t)b * n_warps + gw];
  __syncwarp();
  const int beg = gw * kSortChunk, end = min(n, beg + kSortChunk);
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool act = i < end;
    uint32_t v = 0;
    int d = 256 + lane;  // inactive lanes get unique pseudo-digits
    if (act) { v = vals[i]; d = digit(v); }
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(m & ((1u << lane) - 1u));
    int base = 0;
    if (act) base = off[warp][d];
    __syncwarp();
    if (act) {
      out[base + rank] = v;
      if (rank == 0) off[warp][d] = base + __popc(m);
    }
    __syncwarp();
  }
}

// ---- key-value variant: the keys are materialised once and travel with the values, so a pass streams 4 B (histogram) and
// 8 B in / 8 B out (scatter) per item instead of re-deriving the key of every item through two dependent gathers in every
// pass (round-2 ncu: the four label passes moved 142 MB per image that way).  BITS-bit digits (9 -> 3 passes over 26 bits).
// FIRST: the pass over the per-item byte `cost` of the still unsorted sequence (vals_in = iota, not materialised).
template <int BITS, bool FIRST>
__global__ void __launch_bounds__(kSortWarps * 32) radix_hist_kv_kernel(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ cost,
                                                                          int n, int n_warps, int shift, int* __restrict__ counts) {
  constexpr int NB = 1 << BITS;
  __shared__ int h[kSortWarps][NB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kSortWarps + warp;
  for (int b = lane; b < NB; b += 32) h[warp][b] = 0;
  __syncwarp();
  if (gw < n_warps) {
    const int beg = gw * kSortChunk, end = min(n, beg + kSortChunk);
    for (int i = beg + lane; i < end; i += 32) atomicAdd(&h[warp][FIRST ? (int)cost[i] : (int)((keys[i] >> shift) & (NB - 1))], 1);
    __syncwarp();
    for (int b = lane; b < NB; b += 32) counts[(size_t)b * n_warps + gw] = h[warp][b];
  }
}

template <int BITS, bool FIRST>
__global__ void __launch_bounds__(kSortWarps * 32) radix_scatter_kv_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                             const uint8_t* __restrict__ cost, uint32_t* __restrict__ keys_out,
                                                                             uint32_t* __restrict__ vals_out, int n, int n_warps, int shift,
                                                                             const int* __restrict__ offsets) {
  constexpr int NB = 1 << BITS;
  __shared__ int off[kSortWarps][NB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kSortWarps + warp;
  if (gw >= n_warps) return;
  for (int b = lane; b < NB; b += 32) off[warp][b] = offsets[(size_t)b * n_warps + gw];
  __syncwarp();
  const int beg = gw * kSortChunk, end = min(n, beg + kSortChunk);
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool act = i < end;
    uint32_t k = 0, v = 0;
    int d = NB + lane;  // inactive lanes get unique pseudo-digits
    if (act) {
      k = keys_in[i];
      v = FIRST ? (uint32_t)i : vals_in[i];
      d = FIRST ? (int)cost[i] : (int)((k >> shift) & (NB - 1));
    }
    unsigned m;
    if (FIRST) {
      // peers by ballots over the digit's bits: on the strongly skewed cost byte MATCH.ANY throttled the MIO pipe
      // (ncu: mio_throttle 92 cycles per issue in this pass, 1 in the passes over the component number)
      m = __ballot_sync(0xffffffffu, act);
      if (!act) m = 1u << lane;
#pragma unroll
      for (int b = 0; b < BITS; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, act && ((d >> b) & 1));
        if (act) m &= ((d >> b) & 1) ? bal : ~bal;
      }
    } else {
      m = __match_any_sync(0xffffffffu, d);
    }
    const int rank = __popc(m & ((1u << lane) - 1u));
    int base = 0;
    if (act) base = off[warp][d];
    __syncwarp();
    if (act) {
      keys_out[base + rank] = k;
      vals_out[base + rank] = v;
      if (rank == 0) off[warp][d] = base + __popc(m);
    }
    __syncwarp();
  }
}

struct SortTemp {
  DevBuf counts, scan_tmp;
};

// One stable key-value pass (see above).
template <int BITS, bool FIRST>
inline ezc_status radix_pass_kv(ezc_context* ctx, const uint32_t* keys_in, const uint32_t* vals_in, const uint8_t* cost, uint32_t* keys_out,
                                uint32_t* vals_out, int n, int shift, SortTemp& t) {
  if (n <= 0) return EZC_OK;
  constexpr int NB = 1 << BITS;
  const int n_warps = (n + kSortChunk - 1) / kSortChunk;
  const int grid = (n_warps + kSortWarps - 1) / kSortWarps;
  EZC_CUDA(t.counts.reserve(sizeof(int) * NB * (size_t)n_warps));
  radix_hist_kv_kernel<BITS, FIRST><<<grid, kSortWarps * 32, 0, ctx->stream>>>(keys_in, cost, n, n_warps, shift, t.counts.as<int>());
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  if (ezc_status st = exclusive_scan_i32(ctx, t.counts.as<int>(), t.counts.as<int>(), (long long)NB * n_warps, t.scan_tmp, nullptr)) return st;
  radix_scatter_kv_kernel<BITS, FIRST><<<grid, kSortWarps * 32, 0, ctx->stream>>>(keys_in, vals_in, cost, keys_out, vals_out, n, n_warps, shift,
                                                                                t.counts.as<int>());
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return EZC_OK;
}

// One stable pass: reorders vals[n] into out[n] by the 8-bit digit digit(v).
template <class DigitFn>
inline ezc_status radix_pass(ezc_context* ctx, const uint32_t* vals, uint32_t* out, int n, DigitFn digit, SortTemp& t) {
  if (n <= 0) return EZC_OK;
  const int n_warps = (n + kSortChunk - 1) / kSortChunk;
  const int grid = (n_warps + kSortWarps - 1) / kSortWarps;
  EZC_CUDA(t.counts.reserve(sizeof(int) * 256 * (size_t)n_warps));
  radix_hist_kernel<<<grid, kSortWarps * 32, 0, ctx->stream>>>(vals, n, n_warps, digit, t.counts.as<int>());
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  if (ezc_status st = exclusive_scan_i32(ctx, t.counts.as<int>(), t.counts.as<int>(), 256LL * n_warps, t.scan_tmp, nullptr)) return st;
  radix_scatter_kernel<<<grid, kSortWarps * 32, 0, ctx->stream>>>(vals, out, n, n_warps, digit, t.counts.as<int>());
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return EZC_OK;
}

}  // namespace ezc
