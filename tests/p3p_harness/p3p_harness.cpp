// Host harness: compiles the product's csrc/p3p_core.h for the CPU (the same source the CUDA kernel csrc/p3p.cu
// uses, one warp per view there, one plain loop here) so that the restatement of cv::solvePnPRansac(AP3P) + EPnP can
// be checked against the real cv2 without a GPU (tests/test_p3p_cpu.py).  Test infrastructure only.
#include <vector>
#include <cstring>
#include "../../ezcalib_b200/csrc/p3p_core.h"

extern "C" {

// EZCalibrator::computeP3PRansac for one view.  corners[n_pts][2] float pixels, target[n_pts][3] doubles.
// Returns 1 on success.  rvec_tvec[6] = what cv::solvePnPRansac returns; pose7 = the reference's SE3 after
// Rodrigues -> float -> fitToSO3; mask[n_pts] = inlier flags per TARGET point.
int p3p_harness_view(const float* corners, const double* target, int n_pts, int n_target_pts, double fx, double fy, double cx,
                     double cy, int width, int height, int max_iters, double reproj_thr, double confidence, int min_points,
                     double* rvec_tvec, double* pose7, unsigned char* mask, int* n_inliers, int* n_iters) {
  using namespace p3p;
  const float wb = (float)width - 0.5f, hb = (float)height - 0.5f;
  std::vector<float> X, x;
  std::vector<int> src;
  for (int i = 0; i < n_pts; ++i) {
    mask[i] = 0;
    const float u = corners[2 * i], v = corners[2 * i + 1];
    if (!point_in_image(u, v, wb, hb)) continue;
    float nx, ny;
    normalise_pixel(fx, fy, cx, cy, u, v, &nx, &ny);
    x.push_back(nx); x.push_back(ny);
    for (int k = 0; k < 3; ++k) X.push_back((float)target[3 * i + k]);
    src.push_back(i);
  }
  const int m = (int)src.size();
  *n_inliers = 0; *n_iters = 0;
  if (m < min_points && n_target_pts >= min_points) return 0;
  if (m < 4) return 0;
  std::vector<unsigned char> best(m, 0), cur(m, 0);
  int best_count = 0;
  double R[9], t[3];
  if (m == 4) {
    double Xs[12], xs[8];
    for (int j = 0; j < 4; ++j) { for (int k = 0; k < 3; ++k) Xs[3 * j + k] = X[3 * j + k]; xs[2 * j] = x[2 * j]; xs[2 * j + 1] = x[2 * j + 1]; }
    if (!ransac_hypothesis(Xs, xs, R, t)) return 0;
    for (int i = 0; i < 4; ++i) mask[src[i]] = 1;
    *n_inliers = 4;
    rodrigues_to_vector(R, rvec_tvec);
    for (int k = 0; k < 3; ++k) rvec_tvec[3 + k] = t[k];
    reference_pose7(rvec_tvec, rvec_tvec + 3, pose7);
    return 1;
  }
  CvRng rng((uint64_t)-1);
  int niters = max_iters > 1 ? max_iters : 1;
  const float thr2 = (float)(reproj_thr * reproj_thr);
  int iter = 0;
  for (; iter < niters; ++iter) {
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
    for (int j = 0; j < 4; ++j) { for (int k = 0; k < 3; ++k) Xs[3 * j + k] = X[3 * idx[j] + k]; xs[2 * j] = x[2 * idx[j]]; xs[2 * j + 1] = x[2 * idx[j] + 1]; }
    if (!ransac_hypothesis(Xs, xs, R, t)) continue;
    int cnt = 0;
    for (int i = 0; i < m; ++i) { cur[i] = is_inlier(R, t, &X[3 * i], x[2 * i], x[2 * i + 1], thr2); cnt += cur[i]; }
    if (cnt > (best_count > 3 ? best_count : 3)) {
      best.swap(cur);
      best_count = cnt;
      niters = ransac_update_num_iters(confidence, (double)(m - cnt) / m, 4, niters);
    }
  }
  *n_iters = iter;
  if (best_count <= 0) return 0;
  std::vector<double> pw, us;
  for (int i = 0; i < m; ++i)
    if (best[i]) { for (int k = 0; k < 3; ++k) pw.push_back(X[3 * i + k]); us.push_back(x[2 * i]); us.push_back(x[2 * i + 1]); mask[src[i]] = 1; }
  const int n = best_count;
  *n_inliers = n;
  Epnp E;
  epnp_control_points(pw.data(), n, E);
  double acc[78];
  for (int k = 0; k < 78; ++k) acc[k] = 0;
  for (int i = 0; i < n; ++i) epnp_accumulate(E, &pw[3 * i], us[2 * i], us[2 * i + 1], acc);
  epnp_null_space(acc, E);
  epnp_solve(E, pw.data(), us.data(), n, R, t);
  rodrigues_to_vector(R, rvec_tvec);
  for (int k = 0; k < 3; ++k) rvec_tvec[3 + k] = t[k];
  if (n < min_points && n_target_pts >= min_points) return 0;
  reference_pose7(rvec_tvec, rvec_tvec + 3, pose7);
  return 1;
}

}  // extern "C"
