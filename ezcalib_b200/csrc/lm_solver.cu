// lm_solver.cu -- B200-native Levenberg-Marquardt calibration solve (half (b) of the hot path).
//
// Replaces, behind the C ABI of include/ezcalib_b200.h, what the reference obtains from
// ceres::Solve on the problem built in EZCalibrator::computeCalibration
// (/root/reference/src/ezcalibrator.cpp:105-229) and runMultiCameraCalib (:631-844):
// trust-region LM, Cauchy(1) loss, SE3 left-perturbation pose blocks eliminated by a Schur
// complement onto the <= 12 shared intrinsic columns (Ceres rules: SURVEY.md 8a row B8).
//
// Device pipeline per LM iteration (all FP64):
//   lm_accumulate   one warp per view segment.  Phase 1: one lane per observation evaluates the
//                   residual + analytic Jacobian (lm_models.cuh), applies the Cauchy corrector and
//                   stores the weighted row pair [J_pose(6) | r | 0 | J_shared(k)] column-major in
//                   shared memory.  Phase 2: the (8+k)^2 Gram matrix [J r]^T [J r] is accumulated in
//                   8x8 tiles on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA).  Tile row 0
//                   (pose rows + residual row) is flushed per view into an 8 x 8*CG "strip"
//                   (H_pp, g_p, H_pc, per-view g_c); H_cc tiles stay in registers for the whole launch
//                   and are reduced once, in fixed order.  The cost sum comes out of the same pass.
//   lm_schur        one thread per view: Jacobi scaling, LM damping, 6x6 Cholesky, Y = L^-1 W,
//                   z = L^-1 g; fixed-order reduction of Y^T Y, Y^T z, g_c, |x|^2, gradient norm.
//   (host)          k x k reduced system (k <= 12): scale, damp, Cholesky, solve  [+ one all-reduce
//                   of the packed partial system over NVLink when views are sharded over GPUs]
//   lm_backsub      one thread per view: pose step, candidate pose exp(delta) T, model-cost terms.
//   lm_accumulate   again, AT THE CANDIDATE, into the alternate half of the double-buffered strips:
//                   its cost is the candidate cost; on acceptance the buffers are swapped and the next
//                   iteration's normal equations already exist (speculative evaluation).
// Accept / reject, radius update and termination follow Ceres' TrustRegionMinimizer on the host.
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <limits>
#include <new>

#include "ezc_common.cuh"
#include "lm_models.cuh"

namespace ezc {

constexpr int kCS = 68;            // shared-memory column stride in doubles (64 rows + 4 pad: conflict-free MMA fragment loads)
constexpr int kWarpsPerBlock = 4;  // lm_accumulate block = 128 threads
constexpr int kMaxK = 12;          // max shared columns (2 focal + 2 pp + 8 dist)
constexpr int kLYZ = 100;          // per-view Schur record: L(21) pad z(6) @22, Y(6*12) @28

// layout of the packed reduction buffer "red1" (doubles)
constexpr int R1_HCC = 0;                     // [12][12] upper triangle used
constexpr int R1_COST = 144;                  // cost at x
constexpr int R1_SCHUR = 145;                 // start of the lm_schur section
constexpr int R1_SIZE = 320;                   // 145 + S(<=78) b(k) gc(k) xnorm2 nres nused fail + one gmax slot per rank (<= 64)
constexpr int kMaxSlotRanks = 64;

__host__ __device__ inline int tile_index(int ti, int tj, int NT) { return ti * NT - ti * (ti - 1) / 2 + (tj - ti); }

// -------------------------------------------------------------------------------------------
// State of one ceres::Solve, resident in device memory: the trust-region loop (reduced-system solve, accept / reject,
// radius, termination tests) runs in two one-thread kernels, so the host only enqueues iterations and looks at `done`
// once per group of iterations instead of synchronising twice per iteration.
struct alignas(8) LmDevState {
  int cur, sb;               // which pose buffer / strip buffer holds the current point
  int done, termination, failed_start;
  int iter, n_success, n_invalid, n_jac, n_lin;
  int first, reuse_diagonal, last_successful, fresh, solver_ok, trace_pending;
  int n_residual_blocks;
  double radius, decrease_factor, x_cost, x_norm, initial_cost;
  double scale_c[kMaxK], diag_c[kMaxK], Hcc[kMaxK * kMaxK], gc[kMaxK], uc[kMaxK], dc[kMaxK];
  CamParams cam, cam_cand;
};

struct LmDevOpts {
  int max_num_iterations, max_invalid, trace_capacity;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double max_radius, min_radius, min_relative_decrease, min_diag, max_diag;
  double* trace_cost;
  double* trace_radius;
  int* trace_flag;
  int k, nf, n_slots;
  int cols[kMaxK];
};

// -------------------------------------------------------------------------------------------
struct AccumArgs {
  const float2* obs;
  const double* pts;
  const double* poses;
  const CamParams* cams;  // per-block cameras (multi-camera solve) or nullptr
  CamParams cam;
  int n_blocks, M, period;
  int seg_len, segs_per_block, n_segs;
  float wb, hb;
  int k;
  int colpos[kMaxK];      // full shared column -> Gram column (8 + active index) or -1
  double* strips;         // [n_segs][8*DA]
  double* part;           // [n_warps][16][32]  persistent tiles
  double* cost_part;      // [n_warps]
  // device-driven solve: st != nullptr -> poses / strips / camera come from the state (which = 0: current point,
  // 1: candidate); the launch is a no-op once the solve is done or when the linear solve of this iteration failed
  const LmDevState* st;
  int which;
  const double* poses2[2];
  double* strips2[2];
};

// FP64 tensor-core MMA (DMMA): D(8x8) += A(8x4, row) * B(4x8, col).  Fragments: a = A[lane>>2][lane&3],
// b = B[lane&3][lane>>2], c0/c1 = C[lane>>2][2*(lane&3) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODEL, bool MONO, int CG, bool PER_BLOCK_CAM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
lm_accumulate_kernel(const AccumArgs a) {
  constexpr int DA = 8 * CG;
  constexpr int NTILE = CG * (CG + 1) / 2;
  constexpr int NF = MONO ? 1 : 2;
  constexpr int KFULL = NF + 2 + num_dist(MODEL);
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* J = smem + warp * (DA * kCS);
  const int gw = blockIdx.x * kWarpsPerBlock + warp;
  const int nW = gridDim.x * kWarpsPerBlock;

  const double* poses = a.poses;
  double* strips_out = a.strips;
  CamParams cam_shared = a.cam;
  if (a.st) {
    if (a.st->done || (a.which && !a.st->solver_ok)) return;
    poses = a.poses2[a.which ? 1 - a.st->cur : a.st->cur];
    strips_out = a.strips2[a.which ? 1 - a.st->sb : a.st->sb];
    cam_shared = a.which ? a.st->cam_cand : a.st->cam;
  }
  for (int i = lane; i < DA * kCS; i += 32) J[i] = 0.0;

  // fragment addressing: column (8g + lane>>2), row (row0 + lane&3)
  const int fr = (lane >> 2) * kCS + (lane & 3);
  double acc[NTILE][2];
#pragma unroll
  for (int t = 0; t < NTILE; ++t) acc[t][0] = acc[t][1] = 0.0;
  double cost_prod = 1.0;
  int cost_exp = 0;

  for (int seg = gw; seg < a.n_segs; seg += nW) {
    const int blk = seg / a.segs_per_block;
    const int l0 = (seg - blk * a.segs_per_block) * a.seg_len;
    const int l1 = min(l0 + a.seg_len, a.M);
    double pose[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) pose[i] = poses[(size_t)blk * 7 + i];
    CamParams cam;
    if constexpr (PER_BLOCK_CAM) cam = a.cams[blk]; else cam = cam_shared;
    int nvalid = 0;

    for (int c0 = l0; c0 < l1; c0 += 32) {
      const int n = min(32, l1 - c0);
      const int nr = (n + 3) & ~3;  // rows are consumed four at a time (MMA k = 4)
      const int l = c0 + lane;
      bool valid = false;
      float2 o = make_float2(-1.f, -1.f);
      if (lane < n) {
        o = a.obs[(size_t)blk * a.M + l];
        valid = (o.x > -0.5f && o.y > -0.5f && o.x < a.wb && o.y < a.hb);
      }
      __syncwarp();
      if (lane < nr) {
        if (valid) {
          const double* P = a.pts + (size_t)(l % a.period) * 3;
          double rx, ry, w, half_rho, Jpx[6], Jpy[6], Jcx[12], Jcy[12];
          residual_jacobian_w<MODEL, MONO, true>(cam, pose, P[0], P[1], P[2], (double)o.x, (double)o.y, rx, ry, w, half_rho, Jpx, Jpy,
                                                 Jcx, Jcy);
          // running product of (1+s) with the exponent kept apart: cost = 0.5 (log(prod) + e ln 2) at the end
          cost_prod *= half_rho;
          { int e2; cost_prod = frexp(cost_prod, &e2); cost_exp += e2; }
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            J[j * kCS + lane] = Jpx[j];
            J[j * kCS + 32 + lane] = Jpy[j];
          }
          J[6 * kCS + lane] = w * rx;
          J[6 * kCS + 32 + lane] = w * ry;
#pragma unroll
          for (int jf = 0; jf < KFULL; ++jf) {
            const int c = a.colpos[jf];
            if (c >= 0) {
              J[c * kCS + lane] = Jcx[jf];
              J[c * kCS + 32 + lane] = Jcy[jf];
            }
          }
          ++nvalid;
        } else {
          const int ncol = 8 + a.k;
          for (int c = 0; c < ncol; ++c) {
            J[c * kCS + lane] = 0.0;
            J[c * kCS + 32 + lane] = 0.0;
          }
        }
      }
      __syncwarp();
      // ---- Gram update on the FP64 tensor cores: [J r]^T [J r], 8x8 tiles, 4 rows per MMA
      const int ks = nr >> 2;
      for (int j = 0; j < 2 * ks; ++j) {
        const int row0 = 4 * ((j < ks) ? j : (8 + j - ks));
        double f[CG];
#pragma unroll
        for (int g = 0; g < CG; ++g) f[g] = J[g * 8 * kCS + fr + row0];
        int t = 0;
#pragma unroll
        for (int gi = 0; gi < CG; ++gi)
#pragma unroll
          for (int gj = gi; gj < CG; ++gj) { dmma_m8n8k4(acc[t][0], acc[t][1], f[gi], f[gj]); ++t; }
      }
    }
    // ---- per-view tiles (tile row 0: pose rows, residual row 6, count row 7) -> strip of this segment
    for (int m = 16; m >= 1; m >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, m);
    double* strip = strips_out + (size_t)seg * (8 * DA);
#pragma unroll
    for (int gj = 0; gj < CG; ++gj) {
      double2 v = make_double2(acc[gj][0], acc[gj][1]);
      if (gj == 0 && lane == 31) v.y = (double)nvalid;  // element (7,7)
      *reinterpret_cast<double2*>(strip + (lane >> 2) * DA + gj * 8 + 2 * (lane & 3)) = v;
      acc[gj][0] = acc[gj][1] = 0.0;
    }
  }
  // ---- persistent tiles (H_cc) and cost of this warp
  double* pw = a.part + (size_t)gw * 384;
#pragma unroll
  for (int t = 0; t < NTILE; ++t) *reinterpret_cast<double2*>(pw + t * 64 + lane * 2) = make_double2(acc[t][0], acc[t][1]);
  double cost = 0.5 * (log(cost_prod) + (double)cost_exp * 0.693147180559945309417232121458);
  for (int m = 16; m >= 1; m >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, m);
  if (lane == 0) a.cost_part[gw] = cost;
}

// Sum the per-segment strips of a block (only when a block is split into several segments).
struct StripSel {   // device-driven solve: buffer pair + which point (see AccumArgs)
  const LmDevState* st;
  int which;
  const double* seg2[2];
  double* blk2[2];
};
__global__ void lm_strip_reduce_kernel(const double* __restrict__ seg_strips, int segs_per_block, int strip_len,
                                       double* __restrict__ blk_strips, const StripSel sel) {
  if (sel.st) {
    if (sel.st->done || (sel.which && !sel.st->solver_ok)) return;
    const int si = sel.which ? 1 - sel.st->sb : sel.st->sb;
    seg_strips = sel.seg2[si];
    blk_strips = sel.blk2[si];
  }
  const int blk = blockIdx.x;
  for (int e = threadIdx.x; e < strip_len; e += blockDim.x) {
    double v = 0.0;
    const double* p = seg_strips + (size_t)blk * segs_per_block * strip_len + e;
    for (int s = 0; s < segs_per_block; ++s) v += p[(size_t)s * strip_len];
    blk_strips[(size_t)blk * strip_len + e] = v;
  }
}

// Fixed-order reduction of the persistent H_cc tiles and the cost over all warps of lm_accumulate.
// grid = k(k+1)/2 + 1 blocks of 256 threads.
__global__ void lm_hcc_reduce_kernel(const double* __restrict__ part, const double* __restrict__ cost_part, int n_warps,
                                     int CG, int k, double* __restrict__ red1, const LmDevState* st, int which) {
  __shared__ double sh[256];
  if (st && (st->done || (which && !st->solver_ok))) return;
  const int nent = k * (k + 1) / 2;
  double v = 0.0;
  int out = R1_COST;
  if ((int)blockIdx.x < nent) {
    int a = 0, rem = blockIdx.x, len = k;
    while (rem >= len) { rem -= len; ++a; --len; }
    const int b = a + rem;
    const int R = 8 + a, Cc = 8 + b;
    const int tt = tile_index(R >> 3, Cc >> 3, CG);
    const int off = tt * 64 + ((R & 7) * 4 + ((Cc & 7) >> 1)) * 2 + (Cc & 1);
    for (int w = threadIdx.x; w < n_warps; w += 256) v += part[(size_t)w * 384 + off];
    out = R1_HCC + a * kMaxK + b;
  } else {
    for (int w = threadIdx.x; w < n_warps; w += 256) v += cost_part[w];
  }
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int m = 128; m >= 1; m >>= 1) {
    if ((int)threadIdx.x < m) sh[threadIdx.x] += sh[threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) red1[out] = sh[0];
}

// Generic fixed-order reduction of partial rows: in[nparts][nv] -> out[nv]; values with index >=
// max_from are max-reduced.  grid = nv blocks of 256 threads.
__global__ void lm_reduce_rows_kernel(const double* __restrict__ in, int nparts, int nv, int max_from,
                                      double* __restrict__ out, const LmDevState* st) {
  __shared__ double sh[256];
  if (st && (st->done || !st->solver_ok)) return;
  const int v = blockIdx.x;
  const bool is_max = v >= max_from;
  double acc = 0.0;
  for (int p = threadIdx.x; p < nparts; p += 256) {
    const double x = in[(size_t)p * nv + v];
    acc = is_max ? fmax(acc, x) : acc + x;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int m = 128; m >= 1; m >>= 1) {
    if ((int)threadIdx.x < m) sh[threadIdx.x] = is_max ? fmax(sh[threadIdx.x], sh[threadIdx.x + m]) : sh[threadIdx.x] + sh[threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[v] = sh[0];
}

// Block-wide fixed-order sum/max of one value per thread; result valid in thread 0.
__device__ __forceinline__ double block_reduce(double v, bool is_max, double* sh /*[warps]*/) {
  for (int m = 16; m >= 1; m >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, v, m);
    v = is_max ? fmax(v, o) : v + o;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < nw; ++w) r = is_max ? fmax(r, sh[w]) : r + sh[w];
  return r;
}

// -------------------------------------------------------------------------------------------
struct SchurArgs {
  const double* strips;   // [n_blocks][8*DA]
  const double* poses;
  int n_blocks, DA, k;
  double radius, min_diag, max_diag;
  int first;              // compute the Jacobi scaling of the pose columns (iteration 0 of a Solve)
  int reuse_diag;
  int undamped;           // covariance mode: no scaling, no damping
  double* scale_p;        // [n_blocks][6]
  double* diag_p;         // [n_blocks][6]
  double* lyz;            // [n_blocks][kLYZ]
  double* part;           // [grid][nv]
  int nv;                 // nS + 2k + 4 (+1 max)
  const LmDevState* st;   // device-driven solve: radius / first / reuse_diag and the current buffers come from the state
  const double* strips2[2];
  const double* poses2[2];
};

constexpr int kXS = 36;            // staging stride (32 rows + 4): == 4 mod 16 -> conflict-free MMA fragment loads
constexpr int kSchurPart = 3 * 64 + kMaxK + 6;  // per-warp partial: Gram tiles of [Y z], g_c sums, 5 scalars

// One lane per view does the 6x6 work (scaling, damping, Cholesky, forward substitutions); the reduction
// S = sum_views Y^T Y and Y^T z is a Gram matrix of the stacked rows [Y | z] (6 rows per view) and is accumulated on the
// FP64 tensor cores: after computing row i of its view, every lane stages it in shared memory and the warp folds the 32
// rows into persistent 8x8 DMMA tiles.  KC = number of shared columns the instantiation can hold (7 -> one 8-column
// group [Y(7) z], 12 -> two groups).  z sits in the last column (ZC) of the padded range.
template <int KC>
__global__ void __launch_bounds__(128) lm_schur_kernel(const SchurArgs a) {
  constexpr int KG = (KC + 1 + 7) / 8;
  constexpr int NT = KG * (KG + 1) / 2;
  constexpr int ZC = 8 * KG - 1;
  __shared__ double xs_all[4][8 * KG * kXS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* xs = xs_all[warp];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = a.k, DA = a.DA;
  const bool live = b < a.n_blocks;
  const double* strips_in = a.strips;
  const double* poses_in = a.poses;
  double radius = a.radius;
  int first = a.first, reuse_diag = a.reuse_diag;
  if (a.st) {
    if (a.st->done) return;
    strips_in = a.strips2[a.st->sb];
    poses_in = a.poses2[a.st->cur];
    radius = a.st->radius; first = a.st->first; reuse_diag = a.st->reuse_diagonal;
  }
  for (int i = lane; i < 8 * KG * kXS; i += 32) xs[i] = 0.0;

  double Wm[6][KC];   // W rows, overwritten in place by Y rows
  double H[6][6], g[6], sc[6], lm2[6], L[6][6], z[6], gc[KC];
  double cnt = 0.0;
  double xn2 = 0.0, nres = 0.0, nused = 0.0, fail = 0.0, gmax = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) H[i][j] = (i == j) ? 1.0 : 0.0;
    g[i] = 0.0; sc[i] = 1.0; lm2[i] = 0.0;
#pragma unroll
    for (int c = 0; c < KC; ++c) Wm[i][c] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < KC; ++c) gc[c] = 0.0;
  if (live) {
    const double* st = strips_in + (size_t)b * 8 * DA;
    // 16-byte loads: every 32-byte sector of the strip rows is consumed by two back-to-back loads of the same lane
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double2* r2 = reinterpret_cast<const double2*>(st + i * DA);
      double row[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const double2 v = r2[j]; row[2 * j] = v.x; row[2 * j + 1] = v.y; }
#pragma unroll
      for (int j = i; j < 6; ++j) { H[i][j] = row[j]; H[j][i] = row[j]; }
      g[i] = row[6];
#pragma unroll
      for (int j = 0; j < (KC + 1) / 2; ++j) {
        if (2 * j < k) {
          const double2 v = r2[4 + j];
          Wm[i][2 * j] = v.x;
          if (2 * j + 1 < KC) Wm[i][2 * j + 1] = v.y;   // zero padding of the strip when 2j+1 == k
        }
      }
    }
    {
      const double2* r2 = reinterpret_cast<const double2*>(st + 6 * DA);
#pragma unroll
      for (int j = 0; j < (KC + 1) / 2; ++j) {
        if (2 * j < k) {
          const double2 v = r2[4 + j];
          gc[2 * j] = v.x;
          if (2 * j + 1 < KC) gc[2 * j + 1] = v.y;
        }
      }
    }
    cnt = st[7 * DA + 7];
    if (a.undamped) {
      // sc = 1, lm2 = 0
    } else if (first) {
#pragma unroll
      for (int i = 0; i < 6; ++i) { sc[i] = 1.0 / (1.0 + sqrt(H[i][i])); a.scale_p[(size_t)b * 6 + i] = sc[i]; }
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) sc[i] = a.scale_p[(size_t)b * 6 + i];
    }
    if (!a.undamped) {
      if (!reuse_diag) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double d = fmin(fmax(sc[i] * sc[i] * H[i][i], a.min_diag), a.max_diag);
          a.diag_p[(size_t)b * 6 + i] = d;
          lm2[i] = d / radius;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) lm2[i] = a.diag_p[(size_t)b * 6 + i] / radius;
      }
    }
  }
  // V = S H S + D^2 ; Cholesky (lower)
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = sc[j] * sc[j] * H[j][j] + lm2[j];
#pragma unroll
    for (int m = 0; m < j; ++m) d -= L[j][m] * L[j][m];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    d = sqrt(d);
    L[j][j] = d;
    const double id = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double sacc = sc[i] * sc[j] * H[i][j];
#pragma unroll
      for (int m = 0; m < j; ++m) sacc -= L[i][m] * L[j][m];
      L[i][j] = sacc * id;
    }
  }
  if (live && cnt > 0.0 && !ok) fail = 1.0;

  double acc[NT][2];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
  const int fr = (lane >> 2) * kXS + (lane & 3);
  __syncwarp();
  // row i of z = L^-1 (s g) and Y = L^-1 (s W), staged and folded into the Gram tiles
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double il = 1.0 / L[i][i];
    {
      double sacc = sc[i] * g[i];
#pragma unroll
      for (int m = 0; m < i; ++m) sacc -= L[i][m] * z[m];
      z[i] = sacc * il;
    }
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      double sacc = sc[i] * Wm[i][c];
#pragma unroll
      for (int m = 0; m < i; ++m) sacc -= L[i][m] * Wm[m][c];
      Wm[i][c] = sacc * il;
      xs[c * kXS + lane] = Wm[i][c];
    }
    xs[ZC * kXS + lane] = z[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double f[KG];
#pragma unroll
      for (int q = 0; q < KG; ++q) f[q] = xs[q * 8 * kXS + fr + 4 * j];
      int t = 0;
#pragma unroll
      for (int gi = 0; gi < KG; ++gi)
#pragma unroll
        for (int gj = gi; gj < KG; ++gj) { dmma_m8n8k4(acc[t][0], acc[t][1], f[gi], f[gj]); ++t; }
    }
    __syncwarp();
  }
  if (live) {
    // record for lm_backsub: L(21) | pad | z(6) @22 | Y rows @28 + 12 i   (16-byte aligned rows)
    double* rec = a.lyz + (size_t)b * kLYZ;
    {
      int q = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) rec[q++] = L[i][j];
    }
#pragma unroll
    for (int i = 0; i < 6; i += 2) *reinterpret_cast<double2*>(rec + 22 + i) = make_double2(z[i], z[i + 1]);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < (KC + 1) / 2; ++j)
        if (2 * j < k) *reinterpret_cast<double2*>(rec + 28 + i * kMaxK + 2 * j) = make_double2(Wm[i][2 * j], (2 * j + 1 < KC) ? Wm[i][2 * j + 1] : 0.0);
    if (cnt > 0.0) {
      nres = cnt; nused = 1.0;
      const double* x = poses_in + (size_t)b * 7;
      double ng[6], out[7];
#pragma unroll
      for (int i = 0; i < 6; ++i) ng[i] = -g[i];
      se3_plus(x, ng, out);
#pragma unroll
      for (int i = 0; i < 7; ++i) { xn2 += x[i] * x[i]; gmax = fmax(gmax, fabs(x[i] - out[i])); }
    }
  }
  // ---- per-warp partials (fixed order): Gram tiles, g_c sums, scalars
  const int gw = blockIdx.x * 4 + warp;
  double* pw = a.part + (size_t)gw * kSchurPart;
#pragma unroll
  for (int t = 0; t < NT; ++t) *reinterpret_cast<double2*>(pw + t * 64 + lane * 2) = make_double2(acc[t][0], acc[t][1]);
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    double v = gc[c];
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) pw[192 + c] = v;
  }
  for (int m = 16; m >= 1; m >>= 1) {
    xn2 += __shfl_xor_sync(0xffffffffu, xn2, m);
    nres += __shfl_xor_sync(0xffffffffu, nres, m);
    nused += __shfl_xor_sync(0xffffffffu, nused, m);
    fail += __shfl_xor_sync(0xffffffffu, fail, m);
    gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, m));
  }
  if (lane == 0) {
    pw[192 + kMaxK + 0] = xn2; pw[192 + kMaxK + 1] = nres; pw[192 + kMaxK + 2] = nused; pw[192 + kMaxK + 3] = fail;
    pw[192 + kMaxK + 4] = gmax;
  }
}

// Fixed-order reduction of the per-warp Schur partials into the packed layout
//   [S upper triangle (k(k+1)/2) | Y^T z (k) | g_c (k) | |x|^2, n_res, n_used, fail | gmax(max)].
// grid = nv blocks of 256 threads.
__global__ void lm_schur_reduce_kernel(const double* __restrict__ part, int n_warps, int k, int KG, double* __restrict__ out, int rank,
                                       int n_slots, const LmDevState* st) {
  __shared__ double sh[256];
  if (st && st->done) return;
  const int nS = k * (k + 1) / 2;
  const int q = blockIdx.x;
  const int ZC = 8 * KG - 1;
  int off;
  bool is_max = false;
  if (q < nS + k) {
    int R, Cc;
    if (q < nS) {
      int a = 0, rem = q, len = k;
      while (rem >= len) { rem -= len; ++a; --len; }
      R = a; Cc = a + rem;
    } else {
      R = q - nS; Cc = ZC;
    }
    const int tt = tile_index(R >> 3, Cc >> 3, KG);
    off = tt * 64 + ((R & 7) * 4 + ((Cc & 7) >> 1)) * 2 + (Cc & 1);
  } else if (q < nS + 2 * k) {
    off = 192 + (q - nS - k);
  } else {
    off = 192 + kMaxK + (q - nS - 2 * k);
    is_max = (q - nS - 2 * k) == 4;
  }
  double v = 0.0;
  for (int w = threadIdx.x; w < n_warps; w += 256) {
    const double x = part[(size_t)w * kSchurPart + off];
    v = is_max ? fmax(v, x) : v + x;
  }
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int m = 128; m >= 1; m >>= 1) {
    if ((int)threadIdx.x < m) sh[threadIdx.x] = is_max ? fmax(sh[threadIdx.x], sh[threadIdx.x + m]) : sh[threadIdx.x] + sh[threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (is_max) {
      // the maximum travels through the SUM all-reduce of the packed buffer: one slot per rank, zero in the others
      for (int r = 0; r < n_slots; ++r) out[q + r] = (r == rank) ? sh[0] : 0.0;
    } else {
      out[q] = sh[0];
    }
  }
}

struct BacksubArgs {
  const double* strips;
  const double* poses;
  double* cand_poses;
  const double* scale_p;
  const double* lyz;
  int n_blocks, DA, k;
  double uc[kMaxK];      // s_c * yhat_c  (delta_c = -uc)
  double* step_p;        // [n_blocks][6] tangent step actually applied
  double* part;          // [grid][4]
  const LmDevState* st;  // device-driven solve: uc and the buffers come from the state
  const double* strips2[2];
  double* poses2[2];
};

__global__ void __launch_bounds__(128) lm_backsub_kernel(const BacksubArgs a) {
  __shared__ double sh[8];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = a.k, DA = a.DA;
  double gd = 0.0, dHd = 0.0, step2 = 0.0, cn2 = 0.0;
  const double* strips_in = a.strips;
  const double* poses_in = a.poses;
  double* cand_out = a.cand_poses;
  const double* ucp = a.uc;
  if (a.st) {
    if (a.st->done || !a.st->solver_ok) return;
    strips_in = a.strips2[a.st->sb];
    poses_in = a.poses2[a.st->cur];
    cand_out = a.poses2[1 - a.st->cur];
    ucp = a.st->uc;
  }
  if (b < a.n_blocks) {
    const double* st = strips_in + (size_t)b * 8 * DA;
    const double* rec = a.lyz + (size_t)b * kLYZ;
    const double* x = poses_in + (size_t)b * 7;
    double* xc = cand_out + (size_t)b * 7;
    const double cnt = st[7 * DA + 7];
    if (cnt > 0.0) {
      double L[6][6], tt[6], y[6], d[6];
      {
        // 16-byte loads: L (21 values + pad) and z (6) are 14 double2 from the 800-byte aligned record
        const double2* r2 = reinterpret_cast<const double2*>(rec);
        double lin[28];
#pragma unroll
        for (int j = 0; j < 14; ++j) { const double2 v = r2[j]; lin[2 * j] = v.x; lin[2 * j + 1] = v.y; }
        int q = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) L[i][j] = lin[q++];
#pragma unroll
        for (int i = 0; i < 6; ++i) tt[i] = lin[22 + i];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double2* y2 = reinterpret_cast<const double2*>(rec + 28 + i * kMaxK);
        double sacc = tt[i];
#pragma unroll
        for (int j = 0; j < kMaxK / 2; ++j) {
          if (2 * j < k) {
            const double2 v = y2[j];
            sacc -= v.x * ucp[2 * j];
            if (2 * j + 1 < k) sacc -= v.y * ucp[2 * j + 1];
          }
        }
        tt[i] = sacc;
      }
#pragma unroll
      for (int i = 5; i >= 0; --i) {
        double sacc = tt[i];
#pragma unroll
        for (int m = i + 1; m < 6; ++m) sacc -= L[m][i] * y[m];
        y[i] = sacc / L[i][i];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) { d[i] = -y[i] * a.scale_p[(size_t)b * 6 + i]; a.step_p[(size_t)b * 6 + i] = d[i]; }
      double out[7];
      se3_plus(x, d, out);
#pragma unroll
      for (int i = 0; i < 7; ++i) { xc[i] = out[i]; const double e = x[i] - out[i]; step2 += e * e; cn2 += out[i] * out[i]; }
      // model cost terms in unscaled space: g.delta and delta^T H delta (pose-pose + 2 pose-shared)
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        gd += st[i * DA + 6] * d[i];
        double hrow = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) hrow += ((i <= j) ? st[i * DA + j] : st[j * DA + i]) * d[j];
        double wrow = 0.0;
        for (int c = 0; c < k; ++c) wrow += st[i * DA + 8 + c] * (-ucp[c]);
        dHd += d[i] * (hrow + 2.0 * wrow);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 7; ++i) xc[i] = x[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) a.step_p[(size_t)b * 6 + i] = 0.0;
    }
  }
  double* pout = a.part + (size_t)blockIdx.x * 4;
  double v = block_reduce(gd, false, sh);    if (threadIdx.x == 0) pout[0] = v;
  v = block_reduce(dHd, false, sh);          if (threadIdx.x == 0) pout[1] = v;
  v = block_reduce(step2, false, sh);        if (threadIdx.x == 0) pout[2] = v;
  v = block_reduce(cn2, false, sh);          if (threadIdx.x == 0) pout[3] = v;
}

// -------------------------------------------------------------------------------------------
// The trust-region loop of ceres::Solve on the device (one thread; k <= 12).  Same rules, same order of evaluation as the
// host loop they replace (kept below as lm_solve_host_loop for A/B runs, EZC_LM_HOST_LOOP=1).
__device__ __forceinline__ double* dev_shared_param(CamParams& c, int nf, int j) {
  if (j < nf) return &c.focal[j];
  if (j < nf + 2) return &c.pp[j - nf];
  return &c.dist[j - nf - 2];
}
__device__ __forceinline__ void dev_record(const LmDevState* st, const LmDevOpts& o, int it, double cost, int flag) {
  if (it < o.trace_capacity) {
    if (o.trace_cost) o.trace_cost[it] = cost;
    if (o.trace_radius) o.trace_radius[it] = st->radius;
    if (o.trace_flag) o.trace_flag[it] = flag;
  }
}
__device__ __forceinline__ void dev_take_hcc(LmDevState* st, const double* x2, int k) {
  st->x_cost = x2[R1_COST];
  for (int a = 0; a < k; ++a)
    for (int b = a; b < k; ++b) st->Hcc[a * k + b] = st->Hcc[b * k + a] = x2[R1_HCC + a * kMaxK + b];
}
__device__ __forceinline__ void dev_loop_top(LmDevState* st, const LmDevOpts& o) {
  // TrustRegionMinimizer::FinalizeIterationAndCheckIfMinimizerCanContinue: max iterations, then (in lm_step, once the
  // gradient of this point is known) the gradient tolerance, then the minimum radius
  if (st->iter >= o.max_num_iterations) { st->termination = 3; st->done = 1; }
  else if (st->radius <= o.min_radius) { st->termination = 4; st->done = 1; }
}

// after the first Jacobian evaluation (H_cc + cost of the starting point, all-reduced, in x2)
__device__ void lm_init_compute(LmDevState* st, const double* x2, const LmDevOpts& o) {
  dev_take_hcc(st, x2, o.k);
  st->n_jac = 1;
  st->initial_cost = st->x_cost;
  if (!isfinite(st->x_cost)) { st->failed_start = 1; st->termination = 5; st->done = 1; return; }
  dev_loop_top(st, o);
}

// after lm_schur + reduction (packet x1, all-reduced): gradient test, reduced system, candidate shared parameters
__device__ void lm_step_compute(LmDevState* st, const double* x1, const LmDevOpts& o) {
  if (st->done) return;
  const int k = o.k, nf = o.nf, nS = k * (k + 1) / 2;
  const double* Ssum = x1;
  const double* bsum = x1 + nS;
  const double* gcs = x1 + nS + k;
  const double xn2_p = x1[nS + 2 * k + 0], nres = x1[nS + 2 * k + 1], fail = x1[nS + 2 * k + 3];
  double gmax = 0.0;
  for (int r = 0; r < o.n_slots; ++r) gmax = fmax(gmax, x1[nS + 2 * k + 4 + r]);
  ++st->n_lin;
  if (st->fresh) {
    for (int c = 0; c < k; ++c) st->gc[c] = gcs[c];
    if (st->first) {
      st->n_residual_blocks = (int)nres;
      for (int c = 0; c < k; ++c) st->scale_c[c] = 1.0 / (1.0 + sqrt(st->Hcc[c * k + c]));
      double xs = xn2_p;
      for (int c = 0; c < k; ++c) { const double v = *dev_shared_param(st->cam, nf, o.cols[c]); xs += v * v; }
      st->x_norm = sqrt(xs);
      dev_record(st, o, 0, st->x_cost, 0);
    } else if (st->trace_pending >= 0) {
      dev_record(st, o, st->trace_pending, st->x_cost, 1);
    }
    st->trace_pending = -1;
    st->fresh = 0;
  }
  st->first = 0;
  for (int c = 0; c < k; ++c) gmax = fmax(gmax, fabs(st->gc[c]));
  if (st->last_successful && gmax <= o.gradient_tolerance) { st->termination = 2; st->done = 1; return; }
  ++st->iter;
  if (!st->reuse_diagonal)
    for (int c = 0; c < k; ++c)
      st->diag_c[c] = fmin(fmax(st->scale_c[c] * st->scale_c[c] * st->Hcc[c * k + c], o.min_diag), o.max_diag);
  st->reuse_diagonal = 1;
  bool ok = (fail == 0.0);
  for (int c = 0; c < kMaxK; ++c) { st->uc[c] = 0.0; st->dc[c] = 0.0; }
  if (ok && k > 0) {
    double S[kMaxK * kMaxK], rhs[kMaxK];
    int q = 0;
    for (int a = 0; a < k; ++a)
      for (int b = a; b < k; ++b) {
        const double v = st->scale_c[a] * st->scale_c[b] * (st->Hcc[a * k + b] - Ssum[q]);
        S[a * k + b] = S[b * k + a] = v;
        ++q;
      }
    for (int c = 0; c < k; ++c) {
      S[c * k + c] += st->diag_c[c] / st->radius;
      rhs[c] = st->scale_c[c] * (st->gc[c] - bsum[c]);
    }
    // in-place Cholesky (lower) + solve.  One thread: the FP64 divisions / square roots are ~400-cycle dependent chains, so
    // each column takes ONE reciprocal (the host loop divided every entry) -- last-bit differences only.
    double inv[kMaxK];
    for (int j = 0; j < k && ok; ++j) {
      double d = S[j * k + j];
      for (int m = 0; m < j; ++m) d -= S[j * k + m] * S[j * k + m];
      if (!(d > 0.0)) { ok = false; break; }
      d = sqrt(d);
      S[j * k + j] = d;
      const double id = 1.0 / d;
      inv[j] = id;
      for (int i = j + 1; i < k; ++i) {
        double v = S[i * k + j];
        for (int m = 0; m < j; ++m) v -= S[i * k + m] * S[j * k + m];
        S[i * k + j] = v * id;
      }
    }
    if (ok) {
      for (int i = 0; i < k; ++i) { double v = rhs[i]; for (int m = 0; m < i; ++m) v -= S[i * k + m] * rhs[m]; rhs[i] = v * inv[i]; }
      for (int i = k - 1; i >= 0; --i) { double v = rhs[i]; for (int m = i + 1; m < k; ++m) v -= S[m * k + i] * rhs[m]; rhs[i] = v * inv[i]; }
      for (int c = 0; c < k; ++c) {
        st->uc[c] = st->scale_c[c] * rhs[c];
        st->dc[c] = -st->uc[c];
        if (!isfinite(st->uc[c])) ok = false;
      }
    }
  }
  st->solver_ok = ok ? 1 : 0;
  if (ok) {
    st->cam_cand = st->cam;
    for (int c = 0; c < k; ++c) *dev_shared_param(st->cam_cand, nf, o.cols[c]) += st->dc[c];
    if (nf == 1) st->cam_cand.focal[1] = st->cam_cand.focal[0];
  }
}

// after the speculative evaluation at the candidate (packet x2: H_cc, cost, model-cost sums; all-reduced)
__device__ void lm_decide_compute(LmDevState* st, const double* x2, const LmDevOpts& o) {
  if (st->done) return;
  const int k = o.k, nf = o.nf;
  bool step_valid = false;
  double model_cost_change = 0.0, cand_cost = 0.0, step2 = 0.0, cn2 = 0.0;
  if (st->solver_ok) {
    ++st->n_jac;
    cand_cost = x2[R1_COST];
    double gd = x2[R1_SCHUR + 0], dHd = x2[R1_SCHUR + 1];
    step2 = x2[R1_SCHUR + 2]; cn2 = x2[R1_SCHUR + 3];
    for (int a = 0; a < k; ++a) {
      gd += st->gc[a] * st->dc[a];
      double row = 0.0;
      for (int b = 0; b < k; ++b) row += st->Hcc[a * k + b] * st->dc[b];
      dHd += st->dc[a] * row;
      step2 += st->dc[a] * st->dc[a];
      const double v = *dev_shared_param(st->cam_cand, nf, o.cols[a]);
      cn2 += v * v;
    }
    model_cost_change = -gd - 0.5 * dHd;
    step_valid = model_cost_change > 0.0;
    if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
  }
  if (!step_valid) {
    ++st->n_invalid;
    if (st->n_invalid >= o.max_invalid) { dev_record(st, o, st->iter, st->x_cost, 3); st->termination = 5; st->done = 1; return; }
    st->radius = st->radius / st->decrease_factor; st->decrease_factor *= 2.0; st->reuse_diagonal = 1;
    dev_record(st, o, st->iter, st->x_cost, 3);
    st->last_successful = 0;
    dev_loop_top(st, o);
    return;
  }
  st->n_invalid = 0;
  const double step_norm = sqrt(step2);
  if (step_norm <= o.parameter_tolerance * (st->x_norm + o.parameter_tolerance)) {
    dev_record(st, o, st->iter, st->x_cost, 2); st->termination = 1; st->done = 1; return;
  }
  const double cost_change = st->x_cost - cand_cost;
  if (fabs(cost_change) <= o.function_tolerance * st->x_cost) {
    dev_record(st, o, st->iter, st->x_cost, 2); st->termination = 0; st->done = 1; return;
  }
  const double rel = (cand_cost >= DBL_MAX) ? -DBL_MAX : cost_change / model_cost_change;
  if (rel > o.min_relative_decrease) {
    st->cam = st->cam_cand;
    st->cur = 1 - st->cur;
    st->sb = 1 - st->sb;
    st->x_norm = sqrt(cn2);
    const double t21 = 2.0 * rel - 1.0;
    st->radius = st->radius / fmax(1.0 / 3.0, 1.0 - t21 * t21 * t21);
    st->radius = fmin(o.max_radius, st->radius);
    st->decrease_factor = 2.0; st->reuse_diagonal = 0;
    ++st->n_success;
    st->last_successful = 1;
    dev_take_hcc(st, x2, k);   // x_cost = cand_cost (non-finite was excluded above), H_cc of the new point
    dev_record(st, o, st->iter, st->x_cost, 1);
    st->trace_pending = st->iter;
    st->fresh = 1;
  } else {
    st->radius = st->radius / st->decrease_factor; st->decrease_factor *= 2.0; st->reuse_diagonal = 1;
    st->last_successful = 0;
    dev_record(st, o, st->iter, st->x_cost, 2);
  }
  dev_loop_top(st, o);
}

// One warp stages the state and the packet in shared memory (coalesced), lane 0 runs the control code there (a chain of
// dependent global loads / stores costs ~0.5 us each; in shared memory the whole step is a few microseconds), the warp
// writes the state back.
template <int WHICH>
__global__ void __launch_bounds__(32) lm_control_kernel(LmDevState* st, const double* __restrict__ packet, int packet_len, const LmDevOpts o) {
  __shared__ LmDevState sh;
  __shared__ double px[512];
  static_assert(sizeof(LmDevState) % 8 == 0, "state is copied as doubles");
  constexpr int NW = sizeof(LmDevState) / 8;
  const int lane = threadIdx.x;
  if (WHICH != 0 && st->done) return;
  for (int i = lane; i < NW; i += 32) reinterpret_cast<double*>(&sh)[i] = reinterpret_cast<const double*>(st)[i];
  for (int i = lane; i < packet_len; i += 32) px[i] = packet[i];
  __syncwarp();
  if (lane == 0) {
    if (WHICH == 0) lm_init_compute(&sh, px, o);
    else if (WHICH == 1) lm_step_compute(&sh, px, o);
    else lm_decide_compute(&sh, px, o);
  }
  __syncwarp();
  for (int i = lane; i < NW; i += 32) reinterpret_cast<double*>(st)[i] = reinterpret_cast<const double*>(&sh)[i];
}

// -------------------------------------------------------------------------------------------
struct CostArgs {
  const float2* obs;
  const double* pts;
  const double* poses;
  const CamParams* cams;
  CamParams cam;
  long long n_obs;  // n_blocks * M
  int M, period;
  float wb, hb;
};

// Per-block RMSE of the plain reprojection error: one warp per block (view).
template <int MODEL, bool MONO, bool PER_BLOCK_CAM>
__global__ void __launch_bounds__(128) lm_rmse_kernel(const CostArgs a, int n_blocks, double* __restrict__ rmse) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_blocks) return;
  double acc = 0.0;
  int n = 0;
  for (int l = lane; l < a.M; l += 32) {
    const float2 o = a.obs[(size_t)w * a.M + l];
    if (!(o.x > -0.5f && o.y > -0.5f && o.x < a.wb && o.y < a.hb)) continue;
    const double* P = a.pts + (size_t)(l % a.period) * 3;
    double rx, ry;
    if constexpr (PER_BLOCK_CAM) residual<MODEL, MONO>(a.cams[w], a.poses + (size_t)w * 7, P[0], P[1], P[2], (double)o.x, (double)o.y, rx, ry);
    else residual<MODEL, MONO>(a.cam, a.poses + (size_t)w * 7, P[0], P[1], P[2], (double)o.x, (double)o.y, rx, ry);
    const double nrm = sqrt(rx * rx + ry * ry);  // (px_obs - px_proj).norm()
    acc += nrm * nrm;
    ++n;
  }
  for (int m = 16; m >= 1; m >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, m); n += __shfl_xor_sync(0xffffffffu, n, m); }
  if (lane == 0) rmse[w] = n > 0 ? sqrt(acc / (double)n) : DBL_MAX;
}

// Single-observation debug kernel (test hook behind ezc_lm_residual_jacobian).
template <int MODEL, bool MONO>
__global__ void lm_single_obs_kernel(CamParams cam, const double* __restrict__ in /*pose7 wpt3 u v*/, double* __restrict__ out) {
  constexpr int NF = MONO ? 1 : 2;
  constexpr int KFULL = NF + 2 + num_dist(MODEL);
  double rx, ry, Jpx[6], Jpy[6], Jcx[12], Jcy[12];
  residual_jacobian<MODEL, MONO>(cam, in, in[7], in[8], in[9], in[10], in[11], rx, ry, Jpx, Jpy, Jcx, Jcy);
  constexpr int D = KFULL + 6;
  out[0] = rx; out[1] = ry;
  for (int j = 0; j < KFULL; ++j) { out[2 + j] = Jcx[j]; out[2 + D + j] = Jcy[j]; }
  for (int j = 0; j < 6; ++j) { out[2 + KFULL + j] = Jpx[j]; out[2 + D + KFULL + j] = Jpy[j]; }
}

// ------------------------------------------------------------------------------------------
// Dispatch tables
// ------------------------------------------------------------------------------------------
typedef void (*AccumFn)(const AccumArgs);
typedef void (*RmseFn)(const CostArgs, int, double*);
typedef void (*SingleFn)(CamParams, const double*, double*);

template <int MODEL, bool MONO>
static AccumFn accum_fn(int NT, bool per_block) {
  if (per_block) {
    switch (NT) { case 1: return lm_accumulate_kernel<MODEL, MONO, 1, true>; default: return nullptr; }
  }
  switch (NT) {
    case 1: return lm_accumulate_kernel<MODEL, MONO, 1, false>;
    case 2: return lm_accumulate_kernel<MODEL, MONO, 2, false>;
    case 3: return lm_accumulate_kernel<MODEL, MONO, 3, false>;
  }
  return nullptr;
}

#define EZC_MODEL_SWITCH(model, mono, EXPR)                                         \
  switch ((model) * 2 + ((mono) ? 1 : 0)) {                                         \
    case 0: { constexpr int MD = RAD1; constexpr bool MN = false; EXPR; } break;    \
    case 1: { constexpr int MD = RAD1; constexpr bool MN = true; EXPR; } break;     \
    case 2: { constexpr int MD = RAD2; constexpr bool MN = false; EXPR; } break;    \
    case 3: { constexpr int MD = RAD2; constexpr bool MN = true; EXPR; } break;     \
    case 4: { constexpr int MD = RAD3; constexpr bool MN = false; EXPR; } break;    \
    case 5: { constexpr int MD = RAD3; constexpr bool MN = true; EXPR; } break;     \
    case 6: { constexpr int MD = RADTAN4; constexpr bool MN = false; EXPR; } break; \
    case 7: { constexpr int MD = RADTAN4; constexpr bool MN = true; EXPR; } break;  \
    case 8: { constexpr int MD = RADTAN5; constexpr bool MN = false; EXPR; } break; \
    case 9: { constexpr int MD = RADTAN5; constexpr bool MN = true; EXPR; } break;  \
    case 10: { constexpr int MD = RADTAN8; constexpr bool MN = false; EXPR; } break;\
    case 11: { constexpr int MD = RADTAN8; constexpr bool MN = true; EXPR; } break; \
    case 12: { constexpr int MD = KB4; constexpr bool MN = false; EXPR; } break;    \
    case 13: { constexpr int MD = KB4; constexpr bool MN = true; EXPR; } break;     \
    default: break;                                                                 \
  }

static AccumFn get_accum(int model, bool mono, int NT, bool per_block) {
  AccumFn f = nullptr;
  EZC_MODEL_SWITCH(model, mono, (f = accum_fn<MD, MN>(NT, per_block)));
  return f;
}
static RmseFn get_rmse(int model, bool mono, bool per_block) {
  RmseFn f = nullptr;
  EZC_MODEL_SWITCH(model, mono, (f = per_block ? lm_rmse_kernel<MD, MN, true> : lm_rmse_kernel<MD, MN, false>));
  return f;
}
static SingleFn get_single(int model, bool mono) {
  SingleFn f = nullptr;
  EZC_MODEL_SWITCH(model, mono, (f = lm_single_obs_kernel<MD, MN>));
  return f;
}

// Dense in-place Cholesky / solve for the k x k reduced system (host, k <= 12).
static bool host_cholesky(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int m = 0; m < j; ++m) d -= A[j * n + m] * A[j * n + m];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int m = 0; m < j; ++m) s -= A[i * n + m] * A[j * n + m];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
static void host_chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int m = 0; m < i; ++m) s -= L[i * n + m] * b[m]; b[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int m = i + 1; m < n; ++m) s -= L[m * n + i] * b[m]; b[i] = s / L[i * n + i]; }
}

}  // namespace ezc

using namespace ezc;

// ============================================================================================
// Solver object
// ============================================================================================
struct ezc_lm_solver {
  ezc_context* ctx = nullptr;
  ezc_lm_problem prob{};
  int nf = 1, nd = 1, kfull = 4;
  bool per_block_cam = false;
  bool collective = false;   // this solve exchanges its packets with the other ranks of the context
  int world = 1, rank = 0;   // effective world / rank of THIS solve (1 / 0 when it is local)
  const float2* d_obs = nullptr;
  const double* d_pts = nullptr;
  DevBuf obs_buf, pts_buf;
  DevBuf poses[2];
  DevBuf cams[2];            // per-block CamParams (multi-camera) x / candidate
  CamParams cam{}, cam_cand{};
  std::vector<CamParams> h_cams;
  int cur = 0;
  int seg_len = 0, segs_per_block = 1, n_segs = 0;
  int acc_grid = 0, acc_warps = 0;
  // strips / packed H_cc+cost are double-buffered: [sb] belongs to the current point, [1-sb] receives the speculative
  // evaluation at the candidate (which becomes the current one when the step is accepted)
  DevBuf seg_strips[2], blk_strips[2], red1[2];
  int sb = 0;
  DevBuf part, cost_part, scale_p, diag_p, lyz, step_p;
  DevBuf part2, part3, red2, gmaxbuf, rmse, dbg;
  HostPinned h_red;
  HostPinned h_stage;   // pose download staging for pageable destinations
  // device-driven solve: state, the two fixed exchange buffers (Schur packet | candidate H_cc + cost + model terms; these are
  // what the collectives reduce), device-side trace
  DevBuf dstate, xchg, trace_c, trace_r, trace_f;
  HostPinned h_state;
  // state of the current variable pattern
  int k = 0, NT = 2, DA = 8;
  int cols[kMaxK];
  int colpos[kMaxK];
  bool have_params = false;
};

namespace {

ezc_status launch_check(ezc_context* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    return EZC_ERR_CUDA;
  }
  ctx->kernel_launches++;
  return EZC_OK;
}

void setup_pattern(ezc_lm_solver* s, int opt_focal, int opt_pp, int opt_dist) {
  int k = 0;
  for (int j = 0; j < kMaxK; ++j) s->colpos[j] = -1;
  if (!s->per_block_cam) {
    if (opt_focal) for (int i = 0; i < s->nf; ++i) s->cols[k++] = i;
    if (opt_pp) for (int i = 0; i < 2; ++i) s->cols[k++] = s->nf + i;
    if (opt_dist) for (int i = 0; i < s->nd; ++i) s->cols[k++] = s->nf + 2 + i;
  }
  s->k = k;
  for (int c = 0; c < k; ++c) s->colpos[s->cols[c]] = 8 + c;
  s->NT = (8 + k + 7) / 8;  // 8-column groups of the Gram matrix (MMA tiles)
  s->DA = 8 * s->NT;
}

double* shared_param(CamParams& c, int nf, int j) {
  if (j < nf) return &c.focal[j];
  if (j < nf + 2) return &c.pp[j - nf];
  return &c.dist[j - nf - 2];
}

ezc_status alloc_state(ezc_lm_solver* s) {
  ezc_context* ctx = s->ctx;
  const int nb = s->prob.n_blocks, M = s->prob.obs_per_block;
  // segments: a block (view) is processed by one warp unless it is large (multi-camera solve)
  s->seg_len = M <= 1024 ? M : 512;
  s->segs_per_block = (M + s->seg_len - 1) / s->seg_len;
  s->n_segs = nb * s->segs_per_block;
  const int max_blocks = ctx->sm_count * 4;
  s->acc_grid = std::max(1, std::min((s->n_segs + kWarpsPerBlock - 1) / kWarpsPerBlock, max_blocks));
  s->acc_warps = s->acc_grid * kWarpsPerBlock;
  const size_t strip_max = 8 * 24;
  for (int b = 0; b < 2; ++b) {
    EZC_CUDA(s->seg_strips[b].reserve(sizeof(double) * strip_max * std::max(s->n_segs, 1)));
    if (s->segs_per_block > 1) EZC_CUDA(s->blk_strips[b].reserve(sizeof(double) * strip_max * nb));
    EZC_CUDA(s->red1[b].reserve(sizeof(double) * R1_SIZE));
    EZC_CUDA(cudaMemsetAsync(s->red1[b].p, 0, sizeof(double) * R1_SIZE, ctx->stream));
  }
  EZC_CUDA(s->part.reserve(sizeof(double) * 384 * s->acc_warps));
  EZC_CUDA(s->cost_part.reserve(sizeof(double) * s->acc_warps));
  EZC_CUDA(s->poses[0].reserve(sizeof(double) * 7 * std::max(nb, 1)));
  EZC_CUDA(s->poses[1].reserve(sizeof(double) * 7 * std::max(nb, 1)));
  EZC_CUDA(s->scale_p.reserve(sizeof(double) * 6 * std::max(nb, 1)));
  EZC_CUDA(s->diag_p.reserve(sizeof(double) * 6 * std::max(nb, 1)));
  EZC_CUDA(s->step_p.reserve(sizeof(double) * 6 * std::max(nb, 1)));
  EZC_CUDA(s->lyz.reserve(sizeof(double) * kLYZ * std::max(nb, 1)));
  const int g2 = (nb + 127) / 128 + 1;
  EZC_CUDA(s->part2.reserve(sizeof(double) * kSchurPart * 4 * g2));
  EZC_CUDA(s->part3.reserve(sizeof(double) * 4 * g2));
  EZC_CUDA(s->red2.reserve(sizeof(double) * 8));
  EZC_CUDA(s->gmaxbuf.reserve(sizeof(double) * 2));
  EZC_CUDA(s->rmse.reserve(sizeof(double) * std::max(nb, 1)));
  EZC_CUDA(s->dbg.reserve(sizeof(double) * 64));
  EZC_CUDA(s->h_red.reserve(sizeof(double) * (R1_SIZE + 16)));
  EZC_CUDA(s->dstate.reserve(sizeof(LmDevState)));
  EZC_CUDA(s->xchg.reserve(sizeof(double) * 512));
  EZC_CUDA(cudaMemsetAsync(s->xchg.p, 0, sizeof(double) * 512, ctx->stream));
  EZC_CUDA(s->h_state.reserve(sizeof(LmDevState)));
  if (s->per_block_cam) {
    EZC_CUDA(s->cams[0].reserve(sizeof(CamParams) * std::max(nb, 1)));
    EZC_CUDA(s->cams[1].reserve(sizeof(CamParams) * std::max(nb, 1)));
  }
  return EZC_OK;
}

ezc_status create_common(ezc_context* ctx, const ezc_lm_problem* p, bool device_ptrs, ezc_lm_solver** out) {
  EZC_CHECK(ctx && p && out, EZC_ERR_INVALID, "ezc_lm_create: null argument");
  *out = nullptr;
  EZC_CHECK(p->model >= 0 && p->model < NUM_MODELS, EZC_ERR_INVALID, "ezc_lm_create: unknown model %d", p->model);
  EZC_CHECK(p->n_blocks >= 0 && p->obs_per_block > 0 && p->pts_period > 0, EZC_ERR_INVALID, "ezc_lm_create: bad sizes");
  EZC_CHECK(p->n_blocks == 0 || (p->obs && p->pts), EZC_ERR_INVALID, "ezc_lm_create: null data pointer");
  EZC_CUDA(cudaSetDevice(ctx->device));
  ezc_lm_solver* s = new (std::nothrow) ezc_lm_solver();
  EZC_CHECK(s, EZC_ERR_INVALID, "out of host memory");
  s->ctx = ctx;
  s->prob = *p;
  s->nf = p->use_mono_focal ? 1 : 2;
  s->nd = num_dist(p->model);
  s->kfull = s->nf + 2 + s->nd;
  s->per_block_cam = p->cams_per_block != 0;
  s->collective = ctx->world > 1 && ctx->allreduce && (p->collective > 0 || (p->collective == 0 && p->n_blocks_global != p->n_blocks));
  s->world = s->collective ? ctx->world : 1;
  s->rank = s->collective ? ctx->rank : 0;
  const bool timing = getenv("EZC_LM_TIMING") != nullptr;
  const auto tc0 = std::chrono::steady_clock::now();
  ezc_status st = alloc_state(s);
  if (st != EZC_OK) { delete s; return st; }
  if (timing) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "[lm create] alloc_state %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count()); }
  const size_t n_obs = (size_t)p->n_blocks * p->obs_per_block;
  if (device_ptrs) {
    s->d_obs = reinterpret_cast<const float2*>(p->obs);
    s->d_pts = p->pts;
  } else {
    cudaError_t e = s->obs_buf.reserve(std::max<size_t>(n_obs, 1) * sizeof(float2));
    if (e == cudaSuccess) e = s->pts_buf.reserve((size_t)p->pts_period * 3 * sizeof(double));
    if (e == cudaSuccess && n_obs) e = cudaMemcpyAsync(s->obs_buf.p, p->obs, n_obs * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->pts_buf.p, p->pts, (size_t)p->pts_period * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { set_error("ezc_lm_create: upload failed: %s", cudaGetErrorString(e)); delete s; return EZC_ERR_CUDA; }
    s->d_obs = s->obs_buf.as<float2>();
    s->d_pts = s->pts_buf.as<double>();
  }
  if (timing) fprintf(stderr, "[lm create] total %.3f ms (%zu observation bytes)\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count(), n_obs * sizeof(float2));
  *out = s;
  return EZC_OK;
}

// lm_accumulate (+ strip reduce + H_cc reduce) at the parameters (poses[pose_idx], cam) into strip / red1 buffer `buf`.
constexpr int kX1 = 0, kX2 = 256;   // offsets of the two exchange packets inside ezc_lm_solver::xchg

// dev_which < 0: host-driven (explicit pose buffer / camera / strip buffer, result in red1[buf]);
// dev_which = 0 | 1: device-driven at the current point | the candidate, result in xchg + kX2.
ezc_status run_accumulate(ezc_lm_solver* s, int pose_idx, const CamParams& cam, int buf, int dev_which = -1) {
  ezc_context* ctx = s->ctx;
  AccumArgs a{};
  const LmDevState* dst = dev_which >= 0 ? s->dstate.as<LmDevState>() : nullptr;
  a.st = dst; a.which = dev_which < 0 ? 0 : dev_which;
  for (int b = 0; b < 2; ++b) { a.poses2[b] = s->poses[b].as<double>(); a.strips2[b] = s->seg_strips[b].as<double>(); }
  a.obs = s->d_obs; a.pts = s->d_pts; a.poses = s->poses[pose_idx].as<double>();
  a.cams = s->per_block_cam ? s->cams[0].as<CamParams>() : nullptr;
  a.cam = cam;
  a.n_blocks = s->prob.n_blocks; a.M = s->prob.obs_per_block; a.period = s->prob.pts_period;
  a.seg_len = s->seg_len; a.segs_per_block = s->segs_per_block; a.n_segs = s->n_segs;
  a.wb = (float)s->prob.img_w - 0.5f; a.hb = (float)s->prob.img_h - 0.5f;
  a.k = s->k;
  for (int j = 0; j < kMaxK; ++j) a.colpos[j] = s->colpos[j];
  a.strips = s->seg_strips[buf].as<double>();
  a.part = s->part.as<double>();
  a.cost_part = s->cost_part.as<double>();
  AccumFn fn = get_accum(s->prob.model, s->prob.use_mono_focal != 0, s->NT, s->per_block_cam);
  EZC_CHECK(fn, EZC_ERR_INVALID, "no accumulate kernel for this configuration");
  const size_t smem = sizeof(double) * kWarpsPerBlock * s->DA * kCS;
  if (smem > 48 * 1024) EZC_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (s->n_segs > 0) {
    fn<<<s->acc_grid, kWarpsPerBlock * 32, smem, ctx->stream>>>(a);
    if (ezc_status st = launch_check(ctx, "lm_accumulate")) return st;
    if (s->segs_per_block > 1) {
      StripSel sel{};
      sel.st = dst; sel.which = a.which;
      for (int b = 0; b < 2; ++b) { sel.seg2[b] = s->seg_strips[b].as<double>(); sel.blk2[b] = s->blk_strips[b].as<double>(); }
      lm_strip_reduce_kernel<<<s->prob.n_blocks, 192, 0, ctx->stream>>>(s->seg_strips[buf].as<double>(), s->segs_per_block,
                                                                         8 * s->DA, s->blk_strips[buf].as<double>(), sel);
      if (ezc_status st = launch_check(ctx, "lm_strip_reduce")) return st;
    }
  } else {
    EZC_CUDA(cudaMemsetAsync(s->part.p, 0, sizeof(double) * 384 * s->acc_warps, ctx->stream));
    EZC_CUDA(cudaMemsetAsync(s->cost_part.p, 0, sizeof(double) * s->acc_warps, ctx->stream));
  }
  double* hcc_out = dst ? s->xchg.as<double>() + kX2 : s->red1[buf].as<double>();
  lm_hcc_reduce_kernel<<<s->k * (s->k + 1) / 2 + 1, 256, 0, ctx->stream>>>(s->part.as<double>(), s->cost_part.as<double>(),
                                                                           s->acc_warps, s->NT, s->k, hcc_out, dst, a.which);
  return launch_check(ctx, "lm_hcc_reduce");
}

const double* block_strips(const ezc_lm_solver* s) {
  return s->segs_per_block > 1 ? s->blk_strips[s->sb].as<double>() : s->seg_strips[s->sb].as<double>();
}

ezc_status run_schur(ezc_lm_solver* s, double radius, bool first, bool reuse, bool undamped, double min_diag, double max_diag,
                     bool dev = false) {
  ezc_context* ctx = s->ctx;
  const int nb = s->prob.n_blocks;
  const int k = s->k;
  const int nS = k * (k + 1) / 2;
  const int nv = nS + 2 * k + 5;
  SchurArgs a{};
  a.strips = block_strips(s); a.poses = s->poses[s->cur].as<double>();
  a.n_blocks = nb; a.DA = s->DA; a.k = k;
  a.radius = radius; a.min_diag = min_diag; a.max_diag = max_diag;
  a.first = first; a.reuse_diag = reuse; a.undamped = undamped;
  a.scale_p = s->scale_p.as<double>(); a.diag_p = s->diag_p.as<double>(); a.lyz = s->lyz.as<double>();
  a.part = s->part2.as<double>(); a.nv = nv;
  const LmDevState* dst = dev ? s->dstate.as<LmDevState>() : nullptr;
  a.st = dst;
  for (int b = 0; b < 2; ++b) {
    a.strips2[b] = s->segs_per_block > 1 ? s->blk_strips[b].as<double>() : s->seg_strips[b].as<double>();
    a.poses2[b] = s->poses[b].as<double>();
  }
  const int grid = std::max(1, (nb + 127) / 128);
  if (k <= 7) lm_schur_kernel<7><<<grid, 128, 0, ctx->stream>>>(a);
  else lm_schur_kernel<12><<<grid, 128, 0, ctx->stream>>>(a);
  if (ezc_status st = launch_check(ctx, "lm_schur")) return st;
  // per-warp partials -> red1[R1_SCHUR ..] in fixed order (last value, the gradient max-norm, is max-reduced)
  const bool slots = s->world <= kMaxSlotRanks;
  double* sr_out = dev ? s->xchg.as<double>() + kX1 : s->red1[s->sb].as<double>() + R1_SCHUR;
  lm_schur_reduce_kernel<<<nv, 256, 0, ctx->stream>>>(s->part2.as<double>(), grid * 4, k, k <= 7 ? 1 : 2, sr_out,
                                                      slots ? s->rank : 0, slots ? std::max(s->world, 1) : 1, dst);
  return launch_check(ctx, "lm_schur_reduce");
}

ezc_status allreduce(ezc_lm_solver* s, double* dev, int count, int op) {
  ezc_context* ctx = s->ctx;
  if (s->collective) {
    const int32_t rc = ctx->allreduce(ctx->allreduce_user, dev, count, op, ctx->stream);
    EZC_CHECK(rc == 0, EZC_ERR_COLLECTIVE, "all-reduce callback failed (%d); abort the process group", (int)rc);
  }
  return EZC_OK;
}

ezc_status upload_cams(ezc_lm_solver* s, int which) {
  if (!s->per_block_cam || s->prob.n_blocks == 0) return EZC_OK;
  EZC_CUDA(cudaMemcpyAsync(s->cams[which].p, s->h_cams.data(), sizeof(CamParams) * s->prob.n_blocks, cudaMemcpyHostToDevice, s->ctx->stream));
  return EZC_OK;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

int32_t ezc_num_dist(int32_t model) { return (model >= 0 && model < NUM_MODELS) ? num_dist(model) : -1; }

void ezc_lm_default_options(ezc_lm_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->opt_focal = o->opt_pp = o->opt_dist = 1;
  o->max_num_iterations = 1000;  // /root/reference/src/ezcalibrator.cpp:172
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
}

ezc_status ezc_lm_create(ezc_context* ctx, const ezc_lm_problem* problem, ezc_lm_solver** out) {
  return create_common(ctx, problem, false, out);
}
ezc_status ezc_lm_create_device(ezc_context* ctx, const ezc_lm_problem* problem, ezc_lm_solver** out) {
  return create_common(ctx, problem, true, out);
}
void ezc_lm_destroy(ezc_lm_solver* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  delete s;
}

ezc_status ezc_lm_set_params(ezc_lm_solver* s, const double* focal, const double* pp, const double* dist, const double* poses) {
  EZC_CHECK(s && focal && pp && dist && (poses || s->prob.n_blocks == 0), EZC_ERR_INVALID, "ezc_lm_set_params: null argument");
  EZC_CUDA(cudaSetDevice(s->ctx->device));
  const int nb = s->prob.n_blocks;
  if (s->per_block_cam) {
    s->h_cams.assign(nb, CamParams{});
    for (int b = 0; b < nb; ++b) {
      for (int i = 0; i < s->nf; ++i) s->h_cams[b].focal[i] = focal[b * s->nf + i];
      if (s->nf == 1) s->h_cams[b].focal[1] = focal[b];
      for (int i = 0; i < 2; ++i) s->h_cams[b].pp[i] = pp[b * 2 + i];
      for (int i = 0; i < s->nd; ++i) s->h_cams[b].dist[i] = dist[b * s->nd + i];
    }
    if (ezc_status st = upload_cams(s, 0)) return st;
    if (ezc_status st = upload_cams(s, 1)) return st;
  } else {
    s->cam = CamParams{};
    for (int i = 0; i < s->nf; ++i) s->cam.focal[i] = focal[i];
    if (s->nf == 1) s->cam.focal[1] = focal[0];
    s->cam.pp[0] = pp[0]; s->cam.pp[1] = pp[1];
    for (int i = 0; i < s->nd; ++i) s->cam.dist[i] = dist[i];
  }
  s->cur = 0;
  const bool timing = getenv("EZC_LM_TIMING") != nullptr;
  const auto ts0 = std::chrono::steady_clock::now();
  if (nb) EZC_CUDA(cudaMemcpyAsync(s->poses[0].p, poses, sizeof(double) * 7 * nb, cudaMemcpyHostToDevice, s->ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(s->ctx->stream));
  if (timing) {
    cudaPointerAttributes at{};
    cudaPointerGetAttributes(&at, poses);
    fprintf(stderr, "[lm set_params] pose upload %.3f ms (%zu bytes, host memory type %d)\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count(), sizeof(double) * 7 * (size_t)nb, (int)at.type);
  }
  s->have_params = true;
  return EZC_OK;
}

ezc_status ezc_lm_get_params(ezc_lm_solver* s, double* focal, double* pp, double* dist, double* poses) {
  EZC_CHECK(s && s->have_params, EZC_ERR_INVALID, "ezc_lm_get_params: no parameters set");
  EZC_CUDA(cudaSetDevice(s->ctx->device));
  const int nb = s->prob.n_blocks;
  if (s->per_block_cam) {
    for (int b = 0; b < nb; ++b) {
      if (focal) for (int i = 0; i < s->nf; ++i) focal[b * s->nf + i] = s->h_cams[b].focal[i];
      if (pp) for (int i = 0; i < 2; ++i) pp[b * 2 + i] = s->h_cams[b].pp[i];
      if (dist) for (int i = 0; i < s->nd; ++i) dist[b * s->nd + i] = s->h_cams[b].dist[i];
    }
  } else {
    if (focal) for (int i = 0; i < s->nf; ++i) focal[i] = s->cam.focal[i];
    if (pp) { pp[0] = s->cam.pp[0]; pp[1] = s->cam.pp[1]; }
    if (dist) for (int i = 0; i < s->nd; ++i) dist[i] = s->cam.dist[i];
  }
  if (poses && nb) {
    // a pageable destination makes the driver stage the copy at a fraction of the link rate: stage it ourselves through
    // a (cached) pinned buffer and finish with a host memcpy
    const size_t bytes = sizeof(double) * 7 * (size_t)nb;
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, poses) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    void* dst = poses;
    if (!pinned) {
      EZC_CUDA(s->h_stage.reserve(bytes));
      dst = s->h_stage.p;
    }
    EZC_CUDA(cudaMemcpyAsync(dst, s->poses[s->cur].p, bytes, cudaMemcpyDeviceToHost, s->ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(s->ctx->stream));
    if (!pinned) std::memcpy(poses, dst, bytes);
  }
  return EZC_OK;
}

static ezc_status lm_solve_device(ezc_lm_solver* s, const ezc_lm_options* opt, ezc_lm_summary* summary);
static ezc_status lm_solve_host_loop(ezc_lm_solver* s, const ezc_lm_options* opt, ezc_lm_summary* summary);

ezc_status ezc_lm_solve(ezc_lm_solver* s, const ezc_lm_options* opt, ezc_lm_summary* summary) {
  EZC_CHECK(s && opt && summary, EZC_ERR_INVALID, "ezc_lm_solve: null argument");
  EZC_CHECK(s->have_params, EZC_ERR_INVALID, "ezc_lm_solve: call ezc_lm_set_params first");
  static const bool host_loop = getenv("EZC_LM_HOST_LOOP") != nullptr;   // round-1 driver, kept for A/B runs
  if (host_loop || s->world > kMaxSlotRanks) return lm_solve_host_loop(s, opt, summary);
  return lm_solve_device(s, opt, summary);
}

// The solve with the trust-region loop on the device: per iteration the host enqueues
//   lm_schur -> lm_schur_reduce -> [all-reduce] -> lm_step -> lm_backsub -> lm_reduce_rows -> lm_accumulate(candidate)
//   -> (lm_strip_reduce) -> lm_hcc_reduce -> [all-reduce] -> lm_decide
// and reads the state back once per group of iterations; launches after `done` return immediately.
static ezc_status lm_solve_device(ezc_lm_solver* s, const ezc_lm_options* opt, ezc_lm_summary* summary) {
  ezc_context* ctx = s->ctx;
  EZC_CUDA(cudaSetDevice(ctx->device));
  std::memset(summary, 0, sizeof(*summary));
  setup_pattern(s, opt->opt_focal, opt->opt_pp, opt->opt_dist);
  const int k = s->k, nb = s->prob.n_blocks;
  const int nS = k * (k + 1) / 2;
  const int n_slots = std::max(s->world, 1);
  LmDevOpts o{};
  o.max_num_iterations = opt->max_num_iterations; o.max_invalid = opt->max_num_consecutive_invalid_steps;
  o.function_tolerance = opt->function_tolerance; o.gradient_tolerance = opt->gradient_tolerance;
  o.parameter_tolerance = opt->parameter_tolerance; o.max_radius = opt->max_trust_region_radius;
  o.min_radius = opt->min_trust_region_radius; o.min_relative_decrease = opt->min_relative_decrease;
  o.min_diag = opt->min_lm_diagonal; o.max_diag = opt->max_lm_diagonal;
  o.k = k; o.nf = s->nf; o.n_slots = n_slots;
  for (int c = 0; c < kMaxK; ++c) o.cols[c] = c < k ? s->cols[c] : 0;
  const int cap = std::max(0, opt->trace_capacity);
  o.trace_capacity = cap;
  if (cap > 0) {
    if (opt->trace_cost) { EZC_CUDA(s->trace_c.reserve(sizeof(double) * cap)); o.trace_cost = s->trace_c.as<double>(); }
    if (opt->trace_radius) { EZC_CUDA(s->trace_r.reserve(sizeof(double) * cap)); o.trace_radius = s->trace_r.as<double>(); }
    if (opt->trace_flag) { EZC_CUDA(s->trace_f.reserve(sizeof(int) * cap)); o.trace_flag = s->trace_f.as<int>(); }
  }
  LmDevState* hs = s->h_state.as<LmDevState>();
  std::memset(hs, 0, sizeof(*hs));
  hs->cur = s->cur; hs->sb = s->sb;
  hs->n_success = 1; hs->termination = 3;
  hs->first = 1; hs->fresh = 1; hs->trace_pending = -1;
  hs->radius = opt->initial_trust_region_radius; hs->decrease_factor = 2.0;
  hs->cam = s->cam; hs->cam_cand = s->cam;
  LmDevState* ds = s->dstate.as<LmDevState>();
  EZC_CUDA(cudaMemcpyAsync(ds, hs, sizeof(*hs), cudaMemcpyHostToDevice, ctx->stream));
  double* x1 = s->xchg.as<double>() + kX1;
  double* x2 = s->xchg.as<double>() + kX2;

  if (ezc_status st = run_accumulate(s, 0, s->cam, 0, 0)) return st;
  if (ezc_status st = allreduce(s, x2, R1_SCHUR, 0)) return st;   // H_cc + cost of this rank's views
  lm_control_kernel<0><<<1, 32, 0, ctx->stream>>>(ds, x2, R1_SCHUR, o);
  if (ezc_status st = launch_check(ctx, "lm_init")) return st;

  const int group = opt->max_num_iterations < 4 ? std::max(1, opt->max_num_iterations) : 4;
  const int grid_b = std::max(1, (nb + 127) / 128);
  for (;;) {
    for (int g = 0; g < group; ++g) {
      if (ezc_status st = run_schur(s, 0.0, false, false, false, opt->min_lm_diagonal, opt->max_lm_diagonal, true)) return st;
      if (ezc_status st = allreduce(s, x1, nS + 2 * k + 4 + n_slots, 0)) return st;
      lm_control_kernel<1><<<1, 32, 0, ctx->stream>>>(ds, x1, nS + 2 * k + 4 + n_slots, o);
      if (ezc_status st = launch_check(ctx, "lm_step")) return st;
      BacksubArgs ba{};
      ba.scale_p = s->scale_p.as<double>(); ba.lyz = s->lyz.as<double>();
      ba.n_blocks = nb; ba.DA = s->DA; ba.k = k;
      ba.step_p = s->step_p.as<double>(); ba.part = s->part3.as<double>();
      ba.st = ds;
      for (int b = 0; b < 2; ++b) {
        ba.strips2[b] = s->segs_per_block > 1 ? s->blk_strips[b].as<double>() : s->seg_strips[b].as<double>();
        ba.poses2[b] = s->poses[b].as<double>();
      }
      lm_backsub_kernel<<<grid_b, 128, 0, ctx->stream>>>(ba);
      if (ezc_status st = launch_check(ctx, "lm_backsub")) return st;
      lm_reduce_rows_kernel<<<4, 256, 0, ctx->stream>>>(s->part3.as<double>(), grid_b, 4, 4, x2 + R1_SCHUR, ds);
      if (ezc_status st = launch_check(ctx, "lm_reduce_rows")) return st;
      if (ezc_status st = run_accumulate(s, 0, s->cam, 0, 1)) return st;
      if (ezc_status st = allreduce(s, x2, R1_SCHUR + 4, 0)) return st;
      lm_control_kernel<2><<<1, 32, 0, ctx->stream>>>(ds, x2, R1_SCHUR + 4, o);
      if (ezc_status st = launch_check(ctx, "lm_decide")) return st;
    }
    EZC_CUDA(cudaMemcpyAsync(hs, ds, sizeof(*hs), cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hs->done) break;
  }
  s->cam = hs->cam; s->cam_cand = hs->cam_cand;
  s->cur = hs->cur; s->sb = hs->sb;
  summary->initial_cost = hs->initial_cost;
  summary->n_residual_blocks = hs->n_residual_blocks;
  summary->jacobian_evaluations = hs->n_jac;
  summary->linear_solves = hs->n_lin;
  summary->final_cost = hs->x_cost;
  summary->termination = hs->termination;
  if (hs->failed_start) {
    summary->iterations = 0; summary->num_successful = 0;
    set_error("ezc_lm_solve: non-finite cost at the starting point");
    return EZC_ERR_NUMERIC;
  }
  summary->iterations = hs->iter + 1;
  summary->num_successful = hs->n_success;
  const int nrec = std::min(cap, hs->iter + 1);
  if (nrec > 0) {
    if (opt->trace_cost) EZC_CUDA(cudaMemcpyAsync(opt->trace_cost, s->trace_c.p, sizeof(double) * nrec, cudaMemcpyDeviceToHost, ctx->stream));
    if (opt->trace_radius) EZC_CUDA(cudaMemcpyAsync(opt->trace_radius, s->trace_r.p, sizeof(double) * nrec, cudaMemcpyDeviceToHost, ctx->stream));
    if (opt->trace_flag) EZC_CUDA(cudaMemcpyAsync(opt->trace_flag, s->trace_f.p, sizeof(int) * nrec, cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return EZC_OK;
}

static ezc_status lm_solve_host_loop(ezc_lm_solver* s, const ezc_lm_options* opt, ezc_lm_summary* summary) {
  ezc_context* ctx = s->ctx;
  EZC_CUDA(cudaSetDevice(ctx->device));
  std::memset(summary, 0, sizeof(*summary));
  setup_pattern(s, opt->opt_focal, opt->opt_pp, opt->opt_dist);
  const int k = s->k, nf = s->nf, nb = s->prob.n_blocks;
  const int nS = k * (k + 1) / 2;
  double* h = s->h_red.as<double>();

  double scale_c[kMaxK], diag_c[kMaxK], Hcc[kMaxK * kMaxK], gc[kMaxK];
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false, first = true, last_successful = false;
  double x_cost = 0.0, x_norm = 0.0;
  int iter = 0, n_success = 1, n_invalid = 0, termination = 3;
  int n_jac = 0, n_lin = 0;
  int trace_pending = -1;

  auto record = [&](int it, double cost, int flag) {
    if (it < opt->trace_capacity) {
      if (opt->trace_cost) opt->trace_cost[it] = cost;
      if (opt->trace_radius) opt->trace_radius[it] = radius;
      if (opt->trace_flag) opt->trace_flag[it] = flag;
    }
  };

  if (ezc_status st = run_accumulate(s, s->cur, s->cam, s->sb)) return st;
  if (ezc_status st_ar = allreduce(s, s->red1[s->sb].as<double>(), R1_SCHUR, 0)) return st_ar;  // H_cc + cost of this rank's views
  ++n_jac;
  bool fresh_jacobian = true;

  for (;;) {
    if (iter >= opt->max_num_iterations) { termination = 3; break; }
    if (radius <= opt->min_trust_region_radius) { termination = 4; break; }
    // ---- per-view Schur pieces for the current radius; packed partial system -> host
    if (ezc_status st = run_schur(s, radius, first, reuse_diagonal, false, opt->min_lm_diagonal, opt->max_lm_diagonal)) return st;
    ++n_lin;
    const int n1 = R1_SCHUR + nS + 2 * k + 4;
    // ONE collective per synchronisation point: the gradient max-norm rides in per-rank slots of the summed buffer
    const bool slots = s->world <= kMaxSlotRanks;
    const int n_slots = slots ? std::max(s->world, 1) : 1;
    if (slots) {
      if (ezc_status st_ar = allreduce(s, s->red1[s->sb].as<double>() + R1_SCHUR, n1 - R1_SCHUR + n_slots, 0)) return st_ar;
    } else {
      if (ezc_status st_ar = allreduce(s, s->red1[s->sb].as<double>() + R1_SCHUR, n1 - R1_SCHUR, 0)) return st_ar;
      if (ezc_status st_ar = allreduce(s, s->red1[s->sb].as<double>() + n1, 1, 1)) return st_ar;
    }
    EZC_CUDA(cudaMemcpyAsync(h, s->red1[s->sb].p, sizeof(double) * (n1 + n_slots), cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    const double* hs = h + R1_SCHUR;
    const double* Ssum = hs;
    const double* bsum = hs + nS;
    const double* gcs = hs + nS + k;
    const double xn2_p = hs[nS + 2 * k + 0];
    const double nres = hs[nS + 2 * k + 1];
    const double fail = hs[nS + 2 * k + 3];
    double gmax_p = 0.0;
    for (int r = 0; r < n_slots; ++r) gmax_p = std::max(gmax_p, hs[nS + 2 * k + 4 + r]);
    if (fresh_jacobian) {
      x_cost = h[R1_COST];
      for (int a = 0; a < k; ++a) for (int b = a; b < k; ++b) Hcc[a * k + b] = Hcc[b * k + a] = h[R1_HCC + a * kMaxK + b];
      for (int c = 0; c < k; ++c) gc[c] = gcs[c];
      if (first) {
        summary->initial_cost = x_cost;
        summary->n_residual_blocks = (int)nres;
        if (!std::isfinite(x_cost)) {
          summary->iterations = 0; summary->num_successful = 0; summary->termination = 5;
          summary->final_cost = x_cost; summary->jacobian_evaluations = n_jac; summary->linear_solves = n_lin;
          set_error("ezc_lm_solve: non-finite cost at the starting point");
          return EZC_ERR_NUMERIC;
        }
        for (int c = 0; c < k; ++c) scale_c[c] = 1.0 / (1.0 + std::sqrt(Hcc[c * k + c]));
        double xs = xn2_p;
        for (int c = 0; c < k; ++c) { const double v = *shared_param(s->cam, nf, s->cols[c]); xs += v * v; }
        x_norm = std::sqrt(xs);
        record(0, x_cost, 0);
      } else if (trace_pending >= 0) {
        record(trace_pending, x_cost, 1);
      }
      trace_pending = -1;
      fresh_jacobian = false;
    }
    first = false;
    // gradient tolerance (checked only after a successful step, like Ceres)
    double gmax = gmax_p;
    for (int c = 0; c < k; ++c) gmax = std::max(gmax, std::fabs(gc[c]));
    if (last_successful && gmax <= opt->gradient_tolerance) { termination = 2; break; }
    ++iter;

    // ---- reduced system on the host
    if (!reuse_diagonal)
      for (int c = 0; c < k; ++c)
        diag_c[c] = std::min(std::max(scale_c[c] * scale_c[c] * Hcc[c * k + c], opt->min_lm_diagonal), opt->max_lm_diagonal);
    reuse_diagonal = true;
    bool solver_ok = (fail == 0.0);
    double uc[kMaxK] = {0}, dc[kMaxK] = {0};
    if (solver_ok && k > 0) {
      double S[kMaxK * kMaxK], rhs[kMaxK];
      int q = 0;
      for (int a = 0; a < k; ++a)
        for (int b = a; b < k; ++b) {
          const double v = scale_c[a] * scale_c[b] * (Hcc[a * k + b] - Ssum[q]);
          S[a * k + b] = S[b * k + a] = v;
          ++q;
        }
      for (int c = 0; c < k; ++c) {
        S[c * k + c] += diag_c[c] / radius;
        rhs[c] = scale_c[c] * (gc[c] - bsum[c]);
      }
      if (!host_cholesky(S, k)) solver_ok = false;
      else {
        host_chol_solve(S, k, rhs);  // yhat_c
        for (int c = 0; c < k; ++c) { uc[c] = scale_c[c] * rhs[c]; dc[c] = -uc[c]; if (!std::isfinite(uc[c])) solver_ok = false; }
      }
    }
    bool step_valid = false;
    double model_cost_change = 0.0, cand_cost = 0.0, step2 = 0.0, cn2 = 0.0;
    if (solver_ok) {
      BacksubArgs ba{};
      ba.strips = block_strips(s); ba.poses = s->poses[s->cur].as<double>(); ba.cand_poses = s->poses[1 - s->cur].as<double>();
      ba.scale_p = s->scale_p.as<double>(); ba.lyz = s->lyz.as<double>();
      ba.n_blocks = nb; ba.DA = s->DA; ba.k = k;
      for (int c = 0; c < kMaxK; ++c) ba.uc[c] = uc[c];
      ba.step_p = s->step_p.as<double>(); ba.part = s->part3.as<double>();
      const int grid = std::max(1, (nb + 127) / 128);
      lm_backsub_kernel<<<grid, 128, 0, ctx->stream>>>(ba);
      if (ezc_status st = launch_check(ctx, "lm_backsub")) return st;
      // model-cost sums land right behind the candidate's H_cc / cost block (the Schur section of the alternate buffer is
      // free until that buffer becomes current): one all-reduce and one copy cover both
      lm_reduce_rows_kernel<<<4, 256, 0, ctx->stream>>>(s->part3.as<double>(), grid, 4, 4, s->red1[1 - s->sb].as<double>() + R1_SCHUR, nullptr);
      if (ezc_status st = launch_check(ctx, "lm_reduce_rows")) return st;
      // candidate shared parameters
      s->cam_cand = s->cam;
      for (int c = 0; c < k; ++c) *shared_param(s->cam_cand, nf, s->cols[c]) += dc[c];
      if (nf == 1) s->cam_cand.focal[1] = s->cam_cand.focal[0];
      // speculative evaluation at the candidate: residuals, Jacobian and normal-equation blocks in one pass.  Its cost
      // decides acceptance; when the step is accepted (the common case) the blocks are already the next iteration's.
      const int alt = 1 - s->sb;
      if (ezc_status st = run_accumulate(s, 1 - s->cur, s->cam_cand, alt)) return st;
      ++n_jac;
      if (ezc_status st_ar = allreduce(s, s->red1[alt].as<double>(), R1_SCHUR + 4, 0)) return st_ar;
      EZC_CUDA(cudaMemcpyAsync(h, s->red1[alt].as<double>() + R1_COST, sizeof(double) * 5, cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaStreamSynchronize(ctx->stream));
      double gd = h[1], dHd = h[2];
      step2 = h[3]; cn2 = h[4]; cand_cost = h[0];
      for (int a = 0; a < k; ++a) {
        gd += gc[a] * dc[a];
        double row = 0.0;
        for (int b = 0; b < k; ++b) row += Hcc[a * k + b] * dc[b];
        dHd += dc[a] * row;
        step2 += dc[a] * dc[a];
        const double v = *shared_param(s->cam_cand, nf, s->cols[a]);
        cn2 += v * v;
      }
      // model_cost_change = -(J step)^T (f + J step / 2) = -g.delta - delta^T H delta / 2
      model_cost_change = -gd - 0.5 * dHd;
      step_valid = model_cost_change > 0.0;
      if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    }
    if (!step_valid) {
      ++n_invalid;
      if (n_invalid >= opt->max_num_consecutive_invalid_steps) { record(iter, x_cost, 3); termination = 5; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid == StepRejected(0)
      record(iter, x_cost, 3);
      last_successful = false;
      continue;
    }
    n_invalid = 0;
    const double step_norm = std::sqrt(step2);
    if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) { record(iter, x_cost, 2); termination = 1; break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= opt->function_tolerance * x_cost) { record(iter, x_cost, 2); termination = 0; break; }
    double rel = (cand_cost >= std::numeric_limits<double>::max()) ? std::numeric_limits<double>::lowest() : cost_change / model_cost_change;
    if (rel > opt->min_relative_decrease) {
      s->cam = s->cam_cand;
      s->cur = 1 - s->cur;
      s->sb = 1 - s->sb;
      x_norm = std::sqrt(cn2);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(opt->max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      ++n_success;
      last_successful = true;
      x_cost = cand_cost;
      record(iter, x_cost, 1);
      trace_pending = iter;
      fresh_jacobian = true;  // the candidate's blocks (buffer sb) are the current ones now
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      last_successful = false;
      record(iter, x_cost, 2);
    }
  }
  summary->iterations = iter + 1;
  summary->num_successful = n_success;
  summary->termination = termination;
  summary->final_cost = x_cost;
  summary->jacobian_evaluations = n_jac;
  summary->linear_solves = n_lin;
  return EZC_OK;
}

ezc_status ezc_lm_frame_rmse(ezc_lm_solver* s, double* rmse_out) {
  EZC_CHECK(s && rmse_out && s->have_params, EZC_ERR_INVALID, "ezc_lm_frame_rmse: bad argument");
  ezc_context* ctx = s->ctx;
  EZC_CUDA(cudaSetDevice(ctx->device));
  const int nb = s->prob.n_blocks;
  if (nb == 0) return EZC_OK;
  CostArgs ca{};
  ca.obs = s->d_obs; ca.pts = s->d_pts; ca.poses = s->poses[s->cur].as<double>();
  ca.cams = s->per_block_cam ? s->cams[s->cur].as<CamParams>() : nullptr;
  ca.cam = s->cam;
  ca.n_obs = (long long)nb * s->prob.obs_per_block; ca.M = s->prob.obs_per_block; ca.period = s->prob.pts_period;
  ca.wb = (float)s->prob.img_w - 0.5f; ca.hb = (float)s->prob.img_h - 0.5f;
  RmseFn rf = get_rmse(s->prob.model, s->prob.use_mono_focal != 0, s->per_block_cam);
  const int grid = (nb * 32 + 127) / 128;
  rf<<<grid, 128, 0, ctx->stream>>>(ca, nb, s->rmse.as<double>());
  if (ezc_status st = launch_check(ctx, "lm_rmse")) return st;
  EZC_CUDA(cudaMemcpyAsync(rmse_out, s->rmse.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}

ezc_status ezc_lm_covariance(ezc_lm_solver* s, double* cov_out) {
  EZC_CHECK(s && cov_out && s->have_params, EZC_ERR_INVALID, "ezc_lm_covariance: bad argument");
  EZC_CHECK(!s->per_block_cam, EZC_ERR_INVALID, "ezc_lm_covariance: shared intrinsics only");
  ezc_context* ctx = s->ctx;
  EZC_CUDA(cudaSetDevice(ctx->device));
  setup_pattern(s, 1, 1, 1);
  const int k = s->k, nS = k * (k + 1) / 2;
  if (ezc_status st = run_accumulate(s, s->cur, s->cam, s->sb)) return st;
  if (ezc_status st = run_schur(s, 1.0, false, false, true, 0.0, 0.0)) return st;
  const int n1 = R1_SCHUR + nS + 2 * k + 4;
  if (ezc_status st_ar = allreduce(s, s->red1[s->sb].as<double>(), n1, 0)) return st_ar;
  double* h = s->h_red.as<double>();
  EZC_CUDA(cudaMemcpyAsync(h, s->red1[s->sb].p, sizeof(double) * n1, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  double S[kMaxK * kMaxK];
  int q = 0;
  for (int a = 0; a < k; ++a)
    for (int b = a; b < k; ++b) { S[a * k + b] = S[b * k + a] = h[R1_HCC + a * kMaxK + b] - h[R1_SCHUR + q]; ++q; }
  EZC_CHECK(host_cholesky(S, k), EZC_ERR_NUMERIC, "ezc_lm_covariance: reduced system is rank deficient");
  for (int c = 0; c < k; ++c) {
    double col[kMaxK] = {0};
    col[c] = 1.0;
    host_chol_solve(S, k, col);
    for (int a = 0; a < k; ++a) cov_out[a * k + c] = col[a];
  }
  return EZC_OK;
}

ezc_status ezc_lm_mono(ezc_context* ctx, const ezc_lm_problem* problem, double* focal, double* pp, double* dist,
                       double* poses, double* frame_rmse_out, ezc_lm_summary summaries[3]) {
  EZC_CHECK(ctx && problem && focal && pp && dist && poses && summaries, EZC_ERR_INVALID, "ezc_lm_mono: null argument");
  ezc_lm_solver* s = nullptr;
  ezc_status st = ezc_lm_create(ctx, problem, &s);
  if (st != EZC_OK) return st;
  st = ezc_lm_set_params(s, focal, pp, dist, poses);
  ezc_lm_options o;
  ezc_lm_default_options(&o);
  // stage 1: focal + poses ; stage 2: + distortion ; stage 3: + principal point
  // (/root/reference/src/ezcalibrator.cpp:192-223)
  const int masks[3][3] = {{1, 0, 0}, {1, 0, 1}, {1, 1, 1}};
  for (int i = 0; i < 3 && st == EZC_OK; ++i) {
    o.opt_focal = masks[i][0]; o.opt_pp = masks[i][1]; o.opt_dist = masks[i][2];
    st = ezc_lm_solve(s, &o, &summaries[i]);
  }
  if (st == EZC_OK && frame_rmse_out) st = ezc_lm_frame_rmse(s, frame_rmse_out);
  if (st == EZC_OK) st = ezc_lm_get_params(s, focal, pp, dist, poses);
  ezc_lm_destroy(s);
  return st;
}

ezc_status ezc_lm_multicam(ezc_context* ctx, const ezc_lm_problem* problem, const double* focal, const double* pp, const double* dist,
                           double* extrinsics, ezc_lm_summary* summary) {
  EZC_CHECK(ctx && problem && focal && pp && dist && summary, EZC_ERR_INVALID, "ezc_lm_multicam: null argument");
  EZC_CHECK(problem->cams_per_block != 0, EZC_ERR_INVALID, "ezc_lm_multicam: the problem must carry one constant camera per block");
  EZC_CHECK(extrinsics || problem->n_blocks == 0, EZC_ERR_INVALID, "ezc_lm_multicam: null extrinsics");
  ezc_lm_solver* s = nullptr;
  ezc_status st = ezc_lm_create(ctx, problem, &s);
  if (st != EZC_OK) return st;
  st = ezc_lm_set_params(s, focal, pp, dist, extrinsics);
  ezc_lm_options o;
  ezc_lm_default_options(&o);
  // every intrinsic block is SetParameterBlockConstant (/root/reference/src/ezcalibrator.cpp:664-670); one trust region for all
  // extrinsic blocks, like the single ceres::Problem of the reference (:787-822)
  o.opt_focal = o.opt_pp = o.opt_dist = 0;
  if (st == EZC_OK) st = ezc_lm_solve(s, &o, summary);
  if (st == EZC_OK) st = ezc_lm_get_params(s, nullptr, nullptr, nullptr, extrinsics);
  ezc_lm_destroy(s);
  return st;
}

ezc_status ezc_lm_residual_jacobian(ezc_context* ctx, int32_t model, int32_t mono, const double* focal, const double* pp,
                                    const double* dist, const double* pose7, const double* wpt3, double u, double v,
                                    double* r, double* J) {
  EZC_CHECK(ctx && focal && pp && dist && pose7 && wpt3 && r && J, EZC_ERR_INVALID, "ezc_lm_residual_jacobian: null argument");
  EZC_CHECK(model >= 0 && model < NUM_MODELS, EZC_ERR_INVALID, "unknown model %d", model);
  EZC_CUDA(cudaSetDevice(ctx->device));
  const int nf = mono ? 1 : 2, nd = num_dist(model);
  CamParams cam{};
  for (int i = 0; i < nf; ++i) cam.focal[i] = focal[i];
  if (nf == 1) cam.focal[1] = focal[0];
  cam.pp[0] = pp[0]; cam.pp[1] = pp[1];
  for (int i = 0; i < nd; ++i) cam.dist[i] = dist[i];
  double in[12];
  for (int i = 0; i < 7; ++i) in[i] = pose7[i];
  for (int i = 0; i < 3; ++i) in[7 + i] = wpt3[i];
  in[10] = u; in[11] = v;
  DevBuf buf;
  EZC_CUDA(buf.reserve(sizeof(double) * 96));
  double* d = buf.as<double>();
  EZC_CUDA(cudaMemcpyAsync(d, in, sizeof(in), cudaMemcpyHostToDevice, ctx->stream));
  get_single(model, mono != 0)<<<1, 1, 0, ctx->stream>>>(cam, d, d + 16);
  if (ezc_status st = launch_check(ctx, "lm_single_obs")) return st;
  const int D = nf + 2 + nd + 6;
  double out[2 + 2 * 18];
  EZC_CUDA(cudaMemcpyAsync(out, d + 16, sizeof(double) * (2 + 2 * D), cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  r[0] = out[0]; r[1] = out[1];
  for (int i = 0; i < 2 * D; ++i) J[i] = out[2 + i];
  return EZC_OK;
}

ezc_status ezc_lm_accumulate_only(ezc_lm_solver* s, int32_t opt_focal, int32_t opt_pp, int32_t opt_dist) {
  EZC_CHECK(s && s->have_params, EZC_ERR_INVALID, "ezc_lm_accumulate_only: bad argument");
  EZC_CUDA(cudaSetDevice(s->ctx->device));
  setup_pattern(s, opt_focal, opt_pp, opt_dist);
  return run_accumulate(s, s->cur, s->cam, s->sb);
}

ezc_status ezc_lm_normal_equations(ezc_lm_solver* s, int32_t opt_focal, int32_t opt_pp, int32_t opt_dist, double* H_cc,
                                   double* g_c, double* H_pp, double* H_pc, double* g_p, double* cost) {
  EZC_CHECK(s && s->have_params, EZC_ERR_INVALID, "ezc_lm_normal_equations: bad argument");
  ezc_context* ctx = s->ctx;
  EZC_CUDA(cudaSetDevice(ctx->device));
  setup_pattern(s, opt_focal, opt_pp, opt_dist);
  const int k = s->k, nb = s->prob.n_blocks, DA = s->DA, nS = k * (k + 1) / 2;
  if (ezc_status st = run_accumulate(s, s->cur, s->cam, s->sb)) return st;
  if (ezc_status st = run_schur(s, 1.0, false, false, true, 0.0, 0.0)) return st;
  std::vector<double> h(R1_SIZE), st((size_t)nb * 8 * DA);
  EZC_CUDA(cudaMemcpyAsync(h.data(), s->red1[s->sb].p, sizeof(double) * R1_SIZE, cudaMemcpyDeviceToHost, ctx->stream));
  if (nb) EZC_CUDA(cudaMemcpyAsync(st.data(), block_strips(s), sizeof(double) * st.size(), cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cost) *cost = h[R1_COST];
  if (H_cc) for (int a = 0; a < k; ++a) for (int b = a; b < k; ++b) H_cc[a * k + b] = H_cc[b * k + a] = h[R1_HCC + a * kMaxK + b];
  if (g_c) for (int c = 0; c < k; ++c) g_c[c] = h[R1_SCHUR + nS + k + c];
  for (int b = 0; b < nb; ++b) {
    const double* sp = &st[(size_t)b * 8 * DA];
    if (H_pp) for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) H_pp[b * 36 + i * 6 + j] = H_pp[b * 36 + j * 6 + i] = sp[i * DA + j];
    if (g_p) for (int i = 0; i < 6; ++i) g_p[b * 6 + i] = sp[i * DA + 6];
    if (H_pc) for (int i = 0; i < 6; ++i) for (int c = 0; c < k; ++c) H_pc[((size_t)b * 6 + i) * k + c] = sp[i * DA + 8 + c];
  }
  return EZC_OK;
}

}  // extern "C"
