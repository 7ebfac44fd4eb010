// p3p.cu -- batched EZCalibrator::computeP3PRansac (/root/reference/src/ezcalibrator.cpp:560-628): the step between
// detection and the LM solve.  One warp per view; the mathematics is csrc/p3p_core.h (shared with the host harness of the
// tests), this file only distributes it over the lanes:
//   * compaction of the in-image points (ballot + prefix popcount),
//   * the RANSAC loop: lane 0 owns the cv::RNG stream and the minimal solver, all lanes classify the points of each
//     hypothesis (ballot counts), every lane keeps the same loop state,
//   * EPnP: barycentric coordinates and the 78 entries of M^T M spread over the lanes (each entry summed over the points
//     in order, exactly like the serial host harness), the small dense algebra (Jacobi SVDs, Gauss-Newton) on lane 0.
// Compiled with -fmad=false: the product must round like OpenCV's scalar x86 code.
#include "ezc_common.cuh"
#include "p3p_core.h"

namespace {

struct P3pArgs {
  const float* corners;   // [n_views][n_pts][2]
  const double* target;   // [n_pts][3]
  int n_views, n_pts, n_target_pts;
  double fx, fy, cx, cy;
  float wb, hb;
  int max_iters, min_points;
  double reproj_thr, confidence;
  uint8_t* ok;            // [n_views]
  double* rvec_tvec;      // [n_views][6]
  double* pose7;          // [n_views][7]
  int32_t* n_inliers;     // [n_views]
  int32_t* n_iters;       // [n_views]
  uint8_t* mask;          // [n_views][n_pts]
};

__host__ __device__ inline size_t align8(size_t x) { return (x + 7) & ~size_t(7); }

// shared memory of one view
__host__ __device__ inline size_t p3p_smem_bytes(int n_pts) {
  size_t b = 0;
  b += align8(sizeof(double) * 3 * n_pts);   // pw
  b += align8(sizeof(double) * 2 * n_pts);   // us
  b += align8(sizeof(double) * 4 * n_pts);   // alphas
  b += align8(sizeof(p3p::Epnp));
  b += align8(sizeof(double) * 78);          // M^T M upper triangle
  b += align8(sizeof(double) * 16);          // hypothesis R, t, flag
  b += align8(sizeof(float) * 3 * n_pts);    // X (float model points of the in-image observations)
  b += align8(sizeof(float) * 2 * n_pts);    // x (normalised image points)
  b += align8(sizeof(int) * n_pts);          // source index
  b += align8(2 * (size_t)n_pts);            // two masks
  return b;
}

__global__ void __launch_bounds__(32) p3p_ransac_kernel(P3pArgs a) {
  using namespace p3p;
  const int view = blockIdx.x;
  const int lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* sp = smem;
  auto take = [&](size_t bytes) { unsigned char* r = sp; sp += align8(bytes); return r; };
  double* pw = (double*)take(sizeof(double) * 3 * a.n_pts);
  double* us = (double*)take(sizeof(double) * 2 * a.n_pts);
  double* al = (double*)take(sizeof(double) * 4 * a.n_pts);
  Epnp& E = *(Epnp*)take(sizeof(Epnp));
  double* acc = (double*)take(sizeof(double) * 78);
  double* hyp = (double*)take(sizeof(double) * 16);
  float* X = (float*)take(sizeof(float) * 3 * a.n_pts);
  float* x = (float*)take(sizeof(float) * 2 * a.n_pts);
  int* src = (int*)take(sizeof(int) * a.n_pts);
  unsigned char* mask_a = take(a.n_pts);
  unsigned char* mask_b = mask_a + a.n_pts;

  const float* corners = a.corners + (size_t)view * a.n_pts * 2;
  uint8_t* out_mask = a.mask ? a.mask + (size_t)view * a.n_pts : nullptr;

  // 1. in-image observations, normalised with the prior intrinsics (ezcalibrator.cpp:576-584)
  int m = 0;
  for (int base = 0; base < a.n_pts; base += 32) {
    const int i = base + lane;
    bool in = false;
    float u = 0, v = 0;
    if (i < a.n_pts) {
      u = corners[2 * i]; v = corners[2 * i + 1];
      in = point_in_image(u, v, a.wb, a.hb);
      if (out_mask) out_mask[i] = 0;
    }
    const unsigned bal = __ballot_sync(FULL, in);
    if (in) {
      const int k = m + __popc(bal & ((1u << lane) - 1u));
      float nx, ny;
      normalise_pixel(a.fx, a.fy, a.cx, a.cy, u, v, &nx, &ny);
      x[2 * k] = nx; x[2 * k + 1] = ny;
      for (int c = 0; c < 3; ++c) X[3 * k + c] = (float)a.target[3 * i + c];
      src[k] = i;
    }
    m += __popc(bal);
  }
  __syncwarp();
  if (lane == 0) { a.ok[view] = 0; a.n_inliers[view] = 0; a.n_iters[view] = 0; }
  if ((m < a.min_points && a.n_target_pts >= a.min_points) || m < 4) return;

  double* R = hyp;
  double* t = hyp + 9;
  if (m == 4) {   // cv::solvePnPRansac: model_points == npoints -> the minimal solver alone, all four points inliers
    if (lane == 0) {
      double Xs[12], xs[8];
      for (int j = 0; j < 4; ++j) { for (int c = 0; c < 3; ++c) Xs[3 * j + c] = X[3 * j + c]; xs[2 * j] = x[2 * j]; xs[2 * j + 1] = x[2 * j + 1]; }
      if (ransac_hypothesis(Xs, xs, R, t)) {
        double* rt = a.rvec_tvec + (size_t)view * 6;
        rodrigues_to_vector(R, rt);
        for (int c = 0; c < 3; ++c) rt[3 + c] = t[c];
        reference_pose7(rt, rt + 3, a.pose7 + (size_t)view * 7);
        if (out_mask) for (int j = 0; j < 4; ++j) out_mask[src[j]] = 1;
        a.n_inliers[view] = 4;
        a.ok[view] = 1;
      }
    }
    return;
  }

  // 2. RANSAC (RANSACPointSetRegistrator::run)
  CvRng rng((uint64_t)-1);   // only lane 0 advances it
  int niters = a.max_iters > 1 ? a.max_iters : 1;
  const float thr2 = (float)(a.reproj_thr * a.reproj_thr);
  unsigned char* best = mask_a;
  unsigned char* cur = mask_b;
  int best_count = 0;
  int iter = 0;
  for (; iter < niters; ++iter) {
    if (lane == 0) {
      int idx[4];
      for (int i = 0; i < 4; ++i) {
        int k;
        for (;;) {
          k = rng.uniform(0, m);
          bool dup = false;
          for (int j = 0; j < i; ++j) dup |= idx[j] == k;
          if (!dup) break;
        }
        idx[i] = k;
      }
      double Xs[12], xs[8];
      for (int j = 0; j < 4; ++j) {
        for (int c = 0; c < 3; ++c) Xs[3 * j + c] = X[3 * idx[j] + c];
        xs[2 * j] = x[2 * idx[j]]; xs[2 * j + 1] = x[2 * idx[j] + 1];
      }
      hyp[12] = ransac_hypothesis(Xs, xs, R, t) ? 1.0 : 0.0;
    }
    __syncwarp();
    const bool have = hyp[12] != 0.0;
    if (have) {
      int cnt = 0;
      for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        bool in = false;
        if (i < m) {
          in = is_inlier(R, t, X + 3 * i, x[2 * i], x[2 * i + 1], thr2);
          cur[i] = in;
        }
        cnt += __popc(__ballot_sync(FULL, in));
      }
      if (cnt > (best_count > 3 ? best_count : 3)) {
        unsigned char* tmp = best; best = cur; cur = tmp;
        best_count = cnt;
        niters = ransac_update_num_iters(a.confidence, (double)(m - cnt) / m, 4, niters);
      }
    }
    __syncwarp();
  }
  if (lane == 0) a.n_iters[view] = iter;
  if (best_count <= 0) return;

  // 3. inliers -> doubles, in order (cv::solvePnP(opoints_inliers, ipoints_inliers, ..., SOLVEPNP_EPNP))
  int n = 0;
  for (int base = 0; base < m; base += 32) {
    const int i = base + lane;
    const bool in = i < m && best[i];
    const unsigned bal = __ballot_sync(FULL, in);
    if (in) {
      const int k = n + __popc(bal & ((1u << lane) - 1u));
      for (int c = 0; c < 3; ++c) pw[3 * k + c] = (double)X[3 * i + c];
      us[2 * k] = (double)x[2 * i]; us[2 * k + 1] = (double)x[2 * i + 1];
      if (out_mask) out_mask[src[i]] = 1;
    }
    n += __popc(bal);
  }
  __syncwarp();
  if (lane == 0) epnp_control_points(pw, n, E);
  __syncwarp();
  for (int i = lane; i < n; i += 32) epnp_alphas(E, pw + 3 * i, al + 4 * i);
  __syncwarp();
  // M^T M, upper triangle: entry k = (r, c); rows of M per point: M1[3j] = a_j, M1[3j+2] = -a_j u; M2[3j+1] = a_j, M2[3j+2] = -a_j v
  for (int k = lane; k < 78; k += 32) {
    int r = 0, rem = k;
    while (rem >= 12 - r) { rem -= 12 - r; ++r; }
    const int c = r + rem;
    const int jr = r / 3, kr = r % 3, jc = c / 3, kc = c % 3;
    double s = 0;
    for (int i = 0; i < n; ++i) {
      const double ar = al[4 * i + jr], ac = al[4 * i + jc];
      const double u = us[2 * i], v = us[2 * i + 1];
      const double m1r = kr == 0 ? ar * 1.0 : kr == 1 ? 0.0 : ar * (0.0 - u);
      const double m1c = kc == 0 ? ac * 1.0 : kc == 1 ? 0.0 : ac * (0.0 - u);
      const double m2r = kr == 0 ? 0.0 : kr == 1 ? ar * 1.0 : ar * (0.0 - v);
      const double m2c = kc == 0 ? 0.0 : kc == 1 ? ac * 1.0 : ac * (0.0 - v);
      s += m1r * m1c + m2r * m2c;
    }
    acc[k] = s;
  }
  __syncwarp();
  if (lane == 0) {
    epnp_null_space(acc, E);
    epnp_solve(E, pw, us, n, R, t);
    double* rt = a.rvec_tvec + (size_t)view * 6;
    rodrigues_to_vector(R, rt);
    for (int c = 0; c < 3; ++c) rt[3 + c] = t[c];
    a.n_inliers[view] = n;
    if (!(n < a.min_points && a.n_target_pts >= a.min_points)) {
      reference_pose7(rt, rt + 3, a.pose7 + (size_t)view * 7);
      a.ok[view] = 1;
    }
  }
}

}  // namespace

extern "C" {

ezc_status ezc_p3p_ransac_batch(ezc_context* ctx, const float* corners, const double* target, int32_t n_views, int32_t n_pts,
                                int32_t n_target_pts, const double* focal, int32_t n_focal, const double* pp, int32_t width,
                                int32_t height, int32_t max_iterations, double reproj_threshold, double confidence,
                                int32_t min_points, uint8_t* ok_out, double* poses7_out, int32_t* n_inliers_out,
                                uint8_t* inlier_mask_out, double* rvec_tvec_out, int32_t* n_iterations_out) {
  EZC_CHECK(ctx && corners && target && focal && pp && ok_out && poses7_out, EZC_ERR_INVALID, "ezc_p3p_ransac_batch: null argument");
  EZC_CHECK(n_views >= 0 && n_pts > 0 && (n_focal == 1 || n_focal == 2), EZC_ERR_INVALID, "ezc_p3p_ransac_batch: bad sizes");
  EZC_CHECK(focal[0] > 0 && focal[n_focal - 1] > 0 && reproj_threshold > 0, EZC_ERR_INVALID, "ezc_p3p_ransac_batch: bad intrinsics / threshold");
  EZC_CUDA(cudaSetDevice(ctx->device));
  if (n_views == 0) return EZC_OK;
  const size_t smem = p3p_smem_bytes(n_pts);
  EZC_CHECK(smem <= 200 * 1024, EZC_ERR_CAPACITY, "ezc_p3p_ransac_batch: %d points per view exceed the shared-memory working set", n_pts);
  ezc::DevBuf d_corners, d_target, d_ok, d_rt, d_pose, d_ninl, d_nit, d_mask;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  const size_t nc = (size_t)n_views * n_pts * 2;
  EZC_CUDA(d_corners.reserve(sizeof(float) * nc));
  EZC_CUDA(d_target.reserve(sizeof(double) * 3 * n_pts));
  EZC_CUDA(d_ok.reserve(n_views));
  EZC_CUDA(d_rt.reserve(sizeof(double) * 6 * n_views));
  EZC_CUDA(d_pose.reserve(sizeof(double) * 7 * n_views));
  EZC_CUDA(d_ninl.reserve(sizeof(int32_t) * n_views));
  EZC_CUDA(d_nit.reserve(sizeof(int32_t) * n_views));
  if (inlier_mask_out) EZC_CUDA(d_mask.reserve((size_t)n_views * n_pts));
  EZC_CUDA(cudaMemcpyAsync(d_corners.p, corners, sizeof(float) * nc, cudaMemcpyHostToDevice, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(d_target.p, target, sizeof(double) * 3 * n_pts, cudaMemcpyHostToDevice, ctx->stream));
  EZC_CUDA(cudaMemsetAsync(d_rt.p, 0, sizeof(double) * 6 * n_views, ctx->stream));
  EZC_CUDA(cudaMemsetAsync(d_pose.p, 0, sizeof(double) * 7 * n_views, ctx->stream));
  P3pArgs a;
  a.corners = d_corners.as<float>();
  a.target = d_target.as<double>();
  a.n_views = n_views; a.n_pts = n_pts; a.n_target_pts = n_target_pts;
  a.fx = focal[0]; a.fy = focal[n_focal - 1]; a.cx = pp[0]; a.cy = pp[1];
  a.wb = (float)width - 0.5f; a.hb = (float)height - 0.5f;
  a.max_iters = max_iterations; a.min_points = min_points;
  a.reproj_thr = reproj_threshold; a.confidence = confidence;
  a.ok = d_ok.as<uint8_t>(); a.rvec_tvec = d_rt.as<double>(); a.pose7 = d_pose.as<double>();
  a.n_inliers = d_ninl.as<int32_t>(); a.n_iters = d_nit.as<int32_t>();
  a.mask = inlier_mask_out ? d_mask.as<uint8_t>() : nullptr;
  if (smem > 48 * 1024) EZC_CUDA(cudaFuncSetAttribute(p3p_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p3p_ransac_kernel<<<n_views, 32, smem, ctx->stream>>>(a);
  ctx->kernel_launches++;
  EZC_CUDA(cudaGetLastError());
  EZC_CUDA(cudaMemcpyAsync(ok_out, d_ok.p, n_views, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(poses7_out, d_pose.p, sizeof(double) * 7 * n_views, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_inliers_out) EZC_CUDA(cudaMemcpyAsync(n_inliers_out, d_ninl.p, sizeof(int32_t) * n_views, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_iterations_out) EZC_CUDA(cudaMemcpyAsync(n_iterations_out, d_nit.p, sizeof(int32_t) * n_views, cudaMemcpyDeviceToHost, ctx->stream));
  if (inlier_mask_out) EZC_CUDA(cudaMemcpyAsync(inlier_mask_out, d_mask.p, (size_t)n_views * n_pts, cudaMemcpyDeviceToHost, ctx->stream));
  if (rvec_tvec_out) EZC_CUDA(cudaMemcpyAsync(rvec_tvec_out, d_rt.p, sizeof(double) * 6 * n_views, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}

}  // extern "C"
