// p3p_core.h -- the serial mathematics of EZCalibrator::computeP3PRansac
// (/root/reference/src/ezcalibrator.cpp:560-628), shared unchanged by the CUDA kernel (csrc/p3p.cu, one warp per view)
// and by a host harness (tests/p3p_harness) so that it can be checked against the real OpenCV on a CPU.
//
// The reference calls cv::solvePnPRansac(identity K, no distortion, 1000 iterations, threshold 4 / mean focal,
// confidence 0.999, SOLVEPNP_AP3P).  OpenCV is an un-vendored dependency (4.13 in this image); what it does for these
// arguments, restated from the published algorithm and pinned against cv2 by tests/test_p3p_*.py:
//   * RANSACPointSetRegistrator::run: cv::RNG seeded with (uint64)-1, 4 distinct indices per sample (rejection sampling),
//     minimal solver on points 0..2 with point 3 picking the solution (smallest reprojection error), inliers =
//     float(|x - proj|^2) <= float(thr^2), best model = strictly more inliers, RANSACUpdateNumIters after every improvement;
//   * the minimal solver only influences the inlier mask; the returned pose is cv::solvePnP(inliers, SOLVEPNP_EPNP):
//     Lepetit/Moreno-Noguer/Fua EPnP exactly as modules/calib3d/src/epnp.cpp spells it -- control points from the PCA of
//     the model points, barycentric coordinates through an SVD pseudo-inverse, the 12x12 M^T M null space, three beta
//     initialisations each followed by 5 Gauss-Newton steps (Householder QR), Horn/Arun alignment, best reprojection error;
//   * every SVD is OpenCV's one-sided Jacobi (JacobiSVDImpl_), INCLUDING its treatment of zero singular values (left
//     vectors filled from cv::RNG(0x12345678) and orthogonalised).  That matters here: the reference's targets are
//     planar (z = 1), so the 4th control point coincides with the centroid, three columns of M are exactly zero and three
//     of the four "null vectors" EPnP uses are those pseudo-random vectors.  The sign convention of the Jacobi vectors fixes
//     the control points.  With all of it restated the pose equals cv2's to ~1e-14 (tests).
// The minimal solver itself is Grunert's quartic (Haralick et al. 1994) with Newton-polished roots, not Ke & Roumeliotis'
// AP3P: both return the same (up to four) poses of the same three points.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define P3P_HD __host__ __device__ inline
#else
#define P3P_HD inline
#endif

namespace p3p {

// cv::RNG (multiply-with-carry), modules/core/include/opencv2/core/operations.hpp
struct CvRng {
  uint64_t state;
  P3P_HD explicit CvRng(uint64_t s) : state(s) {}
  P3P_HD unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  P3P_HD int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// cv::RANSACUpdateNumIters
P3P_HD int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.0); p = fmin(p, 1.0);
  ep = fmax(ep, 0.0); ep = fmin(ep, 1.0);
  double num = fmax(1.0 - p, DBL_MIN);
  double denom = 1.0 - pow(1.0 - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

// OpenCV JacobiSVDImpl_<double> (modules/core/src/lapack.cpp).  At: n1 rows of length m (row stride lda); on entry rows
// 0..n-1 hold the COLUMNS of the m x n source matrix (m >= n); on exit row i is the i-th left singular vector.
// Vt (n x n, stride ldv, may be null): right singular vectors as rows.  W[n]: singular values, descending.
P3P_HD void jacobi_svd(double* At, int lda, double* W, double* Vt, int ldv, int m, int n, int n1) {
  const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
  const int max_iter = m > 30 ? m : 30;
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) { const double t = At[i * lda + k]; sd += t * t; }
    W[i] = sd;
    if (Vt) {
      for (int k = 0; k < n; ++k) Vt[i * ldv + k] = 0;
      Vt[i * ldv + i] = 1;
    }
  }
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double* Ai = At + i * lda;
        double* Aj = At + j * lda;
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (int k = 0; k < m; ++k) {
          const double t0 = c * Ai[k] + s * Aj[k];
          const double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0; Aj[k] = t1;
          a += t0 * t0; b += t1 * t1;
        }
        W[i] = a; W[j] = b;
        changed = true;
        if (Vt) {
          double* Vi = Vt + i * ldv;
          double* Vj = Vt + j * ldv;
          for (int k = 0; k < n; ++k) {
            const double t0 = c * Vi[k] + s * Vj[k];
            const double t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0; Vj[k] = t1;
          }
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) { const double t = At[i * lda + k]; sd += t * t; }
    W[i] = sqrt(sd);
  }
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i]; W[i] = W[j]; W[j] = t;
      for (int k = 0; k < m; ++k) { t = At[i * lda + k]; At[i * lda + k] = At[j * lda + k]; At[j * lda + k] = t; }
      if (Vt)
        for (int k = 0; k < n; ++k) { t = Vt[i * ldv + k]; Vt[i * ldv + k] = Vt[j * ldv + k]; Vt[j * ldv + k] = t; }
    }
  }
  CvRng rng(0x12345678);
  for (int i = 0; i < n1; ++i) {
    double sd = i < n ? W[i] : 0;
    for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
      // zero singular value: random vector, made orthogonal to the previous left vectors, normalised
      const double val0 = 1.0 / m;
      for (int k = 0; k < m; ++k) At[i * lda + k] = (rng.next() & 256) != 0 ? val0 : -val0;
      for (int it = 0; it < 2; ++it)
        for (int j = 0; j < i; ++j) {
          sd = 0;
          for (int k = 0; k < m; ++k) sd += At[i * lda + k] * At[j * lda + k];
          double asum = 0;
          for (int k = 0; k < m; ++k) {
            const double t = At[i * lda + k] - sd * At[j * lda + k];
            At[i * lda + k] = t;
            asum += fabs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (int k = 0; k < m; ++k) At[i * lda + k] *= asum;
        }
      sd = 0;
      for (int k = 0; k < m; ++k) { const double t = At[i * lda + k]; sd += t * t; }
      sd = sqrt(sd);
    }
    const double s = sd > minval ? 1 / sd : 0.;
    for (int k = 0; k < m; ++k) At[i * lda + k] *= s;
  }
}

// cv::solve(A, b, x, DECOMP_SVD) for an m x n system (m >= n <= 5, m <= 6): x = V diag(1/w) U^T b with OpenCV's cut-off.
P3P_HD void svd_solve(const double* A, int m, int n, const double* b, double* x) {
  double At[5 * 6], Vt[5 * 5], W[5];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) At[j * m + i] = A[i * n + j];
  jacobi_svd(At, m, W, Vt, n, m, n, n);
  double thr = 0;
  for (int i = 0; i < n; ++i) thr += W[i];
  thr *= DBL_EPSILON * 2;
  for (int j = 0; j < n; ++j) x[j] = 0;
  for (int i = 0; i < n; ++i) {
    if (fabs(W[i]) <= thr) continue;
    double s = 0;
    for (int k = 0; k < m; ++k) s += At[i * m + k] * b[k];
    s /= W[i];
    for (int j = 0; j < n; ++j) x[j] += Vt[i * n + j] * s;
  }
}

// cv::invert(A, Ainv, DECOMP_SVD) for 3x3 (row-major): pseudo-inverse with the same cut-off.
P3P_HD void svd_inv3(const double* A, double* X) {
  double At[9], Vt[9], W[3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) At[j * 3 + i] = A[i * 3 + j];
  jacobi_svd(At, 3, W, Vt, 3, 3, 3, 3);
  double thr = (W[0] + W[1] + W[2]) * DBL_EPSILON * 2;
  for (int i = 0; i < 9; ++i) X[i] = 0;
  for (int i = 0; i < 3; ++i) {
    if (fabs(W[i]) <= thr) continue;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) X[r * 3 + c] += Vt[i * 3 + r] * At[i * 3 + c] / W[i];
  }
}

// epnp::qr_solve (Householder QR, including the quirk that the column maximum skips the last row).
// A: nr x nc row-major (destroyed), b[nr] (destroyed).  Returns false (x untouched) when it meets an all-zero column.
P3P_HD bool qr_solve(double* A, double* b, double* x, int nr, int nc) {
  double A1[4], A2[4];
  for (int k = 0; k < nc; ++k) {
    double eta = fabs(A[k * nc + k]);
    for (int i = k + 1; i < nr; ++i) {
      const double elt = fabs(A[(i - 1) * nc + k]);
      if (eta < elt) eta = elt;
    }
    if (eta == 0) return false;
    const double inv_eta = 1. / eta;
    double sum2 = 0;
    for (int i = k; i < nr; ++i) {
      A[i * nc + k] *= inv_eta;
      sum2 += A[i * nc + k] * A[i * nc + k];
    }
    double sigma = sqrt(sum2);
    if (A[k * nc + k] < 0) sigma = -sigma;
    A[k * nc + k] += sigma;
    A1[k] = sigma * A[k * nc + k];
    A2[k] = -eta * sigma;
    for (int j = k + 1; j < nc; ++j) {
      double sum = 0;
      for (int i = k; i < nr; ++i) sum += A[i * nc + k] * A[i * nc + j];
      const double tau = sum / A1[k];
      for (int i = k; i < nr; ++i) A[i * nc + j] -= tau * A[i * nc + k];
    }
  }
  for (int j = 0; j < nc; ++j) {
    double tau = 0;
    for (int i = j; i < nr; ++i) tau += A[i * nc + j] * b[i];
    tau /= A1[j];
    for (int i = j; i < nr; ++i) b[i] -= tau * A[i * nc + j];
  }
  x[nc - 1] = b[nc - 1] / A2[nc - 1];
  for (int i = nc - 2; i >= 0; --i) {
    double sum = 0;
    for (int j = i + 1; j < nc; ++j) sum += A[i * nc + j] * x[j];
    x[i] = (b[i] - sum) / A2[i];
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Minimal solver: Grunert's quartic.  Real roots of a4 x^4 + .. + a0 by Ferrari's resolvent, then Newton polish.
P3P_HD double cbrt_signed(double x) { return x < 0 ? -pow(-x, 1.0 / 3.0) : pow(x, 1.0 / 3.0); }

P3P_HD int solve_quartic_real(double a4, double a3, double a2, double a1, double a0, double* roots) {
  int n = 0;
  if (fabs(a4) < 1e-300) return 0;
  const double b = a3 / a4, c = a2 / a4, d = a1 / a4, e = a0 / a4;
  // depressed quartic y^4 + p y^2 + q y + r, x = y - b/4
  const double b2 = b * b;
  const double p = c - 0.375 * b2;
  const double q = d - 0.5 * b * c + 0.125 * b2 * b;
  const double r = e - 0.25 * b * d + 0.0625 * b2 * c - (3.0 / 256.0) * b2 * b2;
  double y[4];
  if (fabs(q) < 1e-14 * (1.0 + fabs(p) * sqrt(fabs(p)))) {
    // biquadratic
    const double disc = p * p - 4 * r;
    if (disc >= 0) {
      const double sq = sqrt(disc);
      const double z1 = 0.5 * (-p + sq), z2 = 0.5 * (-p - sq);
      if (z1 >= 0) { y[n++] = sqrt(z1); y[n++] = -sqrt(z1); }
      if (z2 >= 0) { y[n++] = sqrt(z2); y[n++] = -sqrt(z2); }
    }
  } else {
    // resolvent cubic 8 m^3 + 8 p m^2 + (2 p^2 - 8 r) m - q^2 = 0: take a positive real root m
    const double ca = p, cb = 0.25 * p * p - r, cc = -0.125 * q * q;   // m^3 + ca m^2 + cb m + cc
    const double Q = (3 * cb - ca * ca) / 9.0;
    const double R = (9 * ca * cb - 27 * cc - 2 * ca * ca * ca) / 54.0;
    const double D = Q * Q * Q + R * R;
    double m;
    if (D >= 0) {
      const double sD = sqrt(D);
      m = cbrt_signed(R + sD) + cbrt_signed(R - sD) - ca / 3.0;
    } else {
      const double th = acos(R / sqrt(-Q * Q * Q));
      const double sq = 2 * sqrt(-Q);
      const double m0 = sq * cos(th / 3.0) - ca / 3.0;
      const double m1 = sq * cos((th + 2 * 3.14159265358979323846) / 3.0) - ca / 3.0;
      const double m2 = sq * cos((th + 4 * 3.14159265358979323846) / 3.0) - ca / 3.0;
      m = fmax(m0, fmax(m1, m2));
    }
    // polish m on the cubic
    for (int it = 0; it < 3; ++it) {
      const double f = ((m + ca) * m + cb) * m + cc, df = (3 * m + 2 * ca) * m + cb;
      if (df != 0) m -= f / df;
    }
    if (m > 0) {
      const double s = sqrt(2 * m);
      const double t1 = -(2 * p + 2 * m) - 2 * q / s;   // discriminants of the two quadratics
      const double t2 = -(2 * p + 2 * m) + 2 * q / s;
      // y^2 + s y + (p/2 + m - q/(2 s)) = 0  and  y^2 - s y + (p/2 + m + q/(2 s)) = 0
      if (t2 >= 0) { y[n++] = 0.5 * (-s + sqrt(t2)); y[n++] = 0.5 * (-s - sqrt(t2)); }
      if (t1 >= 0) { y[n++] = 0.5 * (s + sqrt(t1)); y[n++] = 0.5 * (s - sqrt(t1)); }
    }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - 0.25 * b;
    for (int it = 0; it < 4; ++it) {
      const double f = (((a4 * x + a3) * x + a2) * x + a1) * x + a0;
      const double df = ((4 * a4 * x + 3 * a3) * x + 2 * a2) * x + a1;
      if (df != 0 && isfinite(f / df)) x -= f / df;
    }
    roots[i] = x;
  }
  return n;
}

P3P_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// orthonormal frame of three points: e1 along P1-P0, e3 normal, e2 = e3 x e1 (columns of F)
P3P_HD bool frame3(const double* P0, const double* P1, const double* P2, double F[9]) {
  double e1[3] = {P1[0] - P0[0], P1[1] - P0[1], P1[2] - P0[2]};
  double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
  if (!(n1 > 0)) return false;
  for (int i = 0; i < 3; ++i) e1[i] /= n1;
  const double d[3] = {P2[0] - P0[0], P2[1] - P0[1], P2[2] - P0[2]};
  double e3[3];
  cross3(e1, d, e3);
  const double n3 = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
  if (!(n3 > 0)) return false;
  for (int i = 0; i < 3; ++i) e3[i] /= n3;
  double e2[3];
  cross3(e3, e1, e2);
  for (int i = 0; i < 3; ++i) { F[i * 3 + 0] = e1[i]; F[i * 3 + 1] = e2[i]; F[i * 3 + 2] = e3[i]; }
  return true;
}

// X[3][3] model points, f[3][3] unit bearing vectors.  Up to 4 poses (R row-major, t).
P3P_HD int p3p_grunert(const double* X, const double* f, double* Rs, double* ts) {
  auto d2 = [](const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
  };
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double a2 = d2(X + 3, X + 6), b2 = d2(X, X + 6), c2 = d2(X, X + 3);
  if (!(a2 > 0 && b2 > 0 && c2 > 0)) return 0;
  const double ca = dot(f + 3, f + 6), cb = dot(f, f + 6), cg = dot(f, f + 3);
  const double q = (a2 - c2) / b2, p = (a2 + c2) / b2;
  const double A4 = (q - 1) * (q - 1) - 4 * c2 / b2 * ca * ca;
  const double A3 = 4 * (q * (1 - q) * cb - (1 - p) * ca * cg + 2 * c2 / b2 * ca * ca * cb);
  const double A2 = 2 * (q * q - 1 + 2 * q * q * cb * cb + 2 * (b2 - c2) / b2 * ca * ca - 4 * p * ca * cb * cg + 2 * (b2 - a2) / b2 * cg * cg);
  const double A1 = 4 * (-q * (1 + q) * cb + 2 * a2 / b2 * cg * cg * cb - (1 - p) * ca * cg);
  const double A0 = (1 + q) * (1 + q) - 4 * a2 / b2 * cg * cg;
  double roots[4];
  const int nr = solve_quartic_real(A4, A3, A2, A1, A0, roots);
  double Fw[9];
  if (!frame3(X, X + 3, X + 6, Fw)) return 0;
  int ns = 0;
  for (int i = 0; i < nr; ++i) {
    const double v = roots[i];
    if (!(v > 0)) continue;
    const double den = 2 * (cg - v * ca);
    if (den == 0) continue;
    const double u = ((q - 1) * v * v - 2 * q * cb * v + 1 + q) / den;
    if (!(u > 0)) continue;
    const double s1sq = c2 / (1 + u * u - 2 * u * cg);
    if (!(s1sq > 0)) continue;
    const double s1 = sqrt(s1sq), s2 = u * s1, s3 = v * s1;
    double Y[9];
    for (int k = 0; k < 3; ++k) { Y[k] = s1 * f[k]; Y[3 + k] = s2 * f[3 + k]; Y[6 + k] = s3 * f[6 + k]; }
    double Fc[9];
    if (!frame3(Y, Y + 3, Y + 6, Fc)) continue;
    double* R = Rs + 9 * ns;
    double* t = ts + 3 * ns;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[r * 3 + c] = Fc[r * 3 + 0] * Fw[c * 3 + 0] + Fc[r * 3 + 1] * Fw[c * 3 + 1] + Fc[r * 3 + 2] * Fw[c * 3 + 2];
    for (int r = 0; r < 3; ++r) t[r] = Y[r] - (R[r * 3 + 0] * X[0] + R[r * 3 + 1] * X[1] + R[r * 3 + 2] * X[2]);
    ++ns;
  }
  return ns;
}

// One RANSAC hypothesis: minimal solver on sample points 0..2, sample point 3 picks the solution.
// Xs[4][3] model points, xs[4][2] normalised image points.  Returns false when there is no solution.
P3P_HD bool ransac_hypothesis(const double* Xs, const double* xs, double* R, double* t) {
  double f[9];
  for (int i = 0; i < 3; ++i) {
    const double nrm = sqrt(xs[2 * i] * xs[2 * i] + xs[2 * i + 1] * xs[2 * i + 1] + 1.0);
    f[3 * i] = xs[2 * i] / nrm; f[3 * i + 1] = xs[2 * i + 1] / nrm; f[3 * i + 2] = 1.0 / nrm;
  }
  double Rs[36], ts[12];
  const int ns = p3p_grunert(Xs, f, Rs, ts);
  int best = -1;
  double best_e = 0;
  for (int s = 0; s < ns; ++s) {
    const double* Rm = Rs + 9 * s;
    const double* X = Xs + 9;
    const double px = Rm[0] * X[0] + Rm[1] * X[1] + Rm[2] * X[2] + ts[3 * s];
    const double py = Rm[3] * X[0] + Rm[4] * X[1] + Rm[5] * X[2] + ts[3 * s + 1];
    const double pz = Rm[6] * X[0] + Rm[7] * X[1] + Rm[8] * X[2] + ts[3 * s + 2];
    const double ex = px / pz - xs[6], ey = py / pz - xs[7];
    const double e = ex * ex + ey * ey;
    if (!(e == e)) continue;
    if (best < 0 || e < best_e) { best = s; best_e = e; }
  }
  if (best < 0) return false;
  for (int i = 0; i < 9; ++i) R[i] = Rs[9 * best + i];
  for (int i = 0; i < 3; ++i) t[i] = ts[3 * best + i];
  return true;
}

// PnPRansacCallback::computeError + findInliers for one point: projection rounded to float (cv::projectPoints writes
// Point2f for float model points), difference in float, squared norm accumulated in double, stored as float.
P3P_HD bool is_inlier(const double* R, const double* t, const float* Xf, float ix, float iy, float thr2) {
  const double X = Xf[0], Y = Xf[1], Z = Xf[2];
  const double px = R[0] * X + R[1] * Y + R[2] * Z + t[0];
  const double py = R[3] * X + R[4] * Y + R[5] * Z + t[1];
  double pz = R[6] * X + R[7] * Y + R[8] * Z + t[2];
  pz = pz ? 1. / pz : 1.;
  const float u = (float)(px * pz), v = (float)(py * pz);
  const float dx = ix - u, dy = iy - v;
  const float err = (float)((double)dx * dx + (double)dy * dy);
  return err <= thr2;
}

// ---------------------------------------------------------------------------------------------------------------
// EPnP (modules/calib3d/src/epnp.cpp), identity camera matrix.  pw[n][3], us[n][2] are the inlier points in doubles.
struct Epnp {
  double cws[4][3];
  double cc_inv[9];
  double ut[12 * 12];   // rows = left singular vectors of M^T M
  double l_6x10[60];
  double rho[6];
};

P3P_HD void epnp_control_points(const double* pw, int n, Epnp& E) {
  for (int j = 0; j < 3; ++j) E.cws[0][j] = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) E.cws[0][j] += pw[3 * i + j];
  for (int j = 0; j < 3; ++j) E.cws[0][j] /= n;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const double d[3] = {pw[3 * i] - E.cws[0][0], pw[3 * i + 1] - E.cws[0][1], pw[3 * i + 2] - E.cws[0][2]};
    for (int r = 0; r < 3; ++r)
      for (int c = r; c < 3; ++c) C[r * 3 + c] += d[r] * d[c];
  }
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  double At[9], dc[3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) At[j * 3 + i] = C[i * 3 + j];
  jacobi_svd(At, 3, dc, nullptr, 0, 3, 3, 3);
  // NOTE cvSVD is asked for U^T (CV_SVD_U_T) without V: OpenCV then still runs the Jacobi with Vt for the left vectors;
  // rows of At are the left singular vectors either way.
  for (int i = 1; i < 4; ++i) {
    const double k = sqrt(dc[i - 1] / n);
    for (int j = 0; j < 3; ++j) E.cws[i][j] = E.cws[0][j] + k * At[3 * (i - 1) + j];
  }
  double cc[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = E.cws[j][i] - E.cws[0][i];
  svd_inv3(cc, E.cc_inv);
}

P3P_HD void epnp_alphas(const Epnp& E, const double* p, double* a) {
  const double d0 = p[0] - E.cws[0][0], d1 = p[1] - E.cws[0][1], d2 = p[2] - E.cws[0][2];
  for (int j = 0; j < 3; ++j) a[1 + j] = E.cc_inv[3 * j] * d0 + E.cc_inv[3 * j + 1] * d1 + E.cc_inv[3 * j + 2] * d2;
  a[0] = 1.0f - a[1] - a[2] - a[3];
}

// the two rows of M of one correspondence (fu = fv = 1, uc = vc = 0)
P3P_HD void epnp_rows(const double* a, double u, double v, double* M1, double* M2) {
  for (int i = 0; i < 4; ++i) {
    M1[3 * i] = a[i] * 1.0; M1[3 * i + 1] = 0.0; M1[3 * i + 2] = a[i] * (0.0 - u);
    M2[3 * i] = 0.0; M2[3 * i + 1] = a[i] * 1.0; M2[3 * i + 2] = a[i] * (0.0 - v);
  }
}

// accumulate the upper triangle (78 entries, row-major packed) of M^T M for one correspondence
P3P_HD void epnp_accumulate(const Epnp& E, const double* p, double u, double v, double* acc78) {
  double a[4], M1[12], M2[12];
  epnp_alphas(E, p, a);
  epnp_rows(a, u, v, M1, M2);
  int k = 0;
  for (int r = 0; r < 12; ++r)
    for (int c = r; c < 12; ++c, ++k) acc78[k] += M1[r] * M1[c] + M2[r] * M2[c];
}

P3P_HD double epnp_dist2(const double* a, const double* b) {
  return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}
P3P_HD double epnp_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// from the accumulated M^T M (upper triangle): null vectors, L_6x10, rho
P3P_HD void epnp_null_space(const double* acc78, Epnp& E) {
  double* At = E.ut;
  int k = 0;
  for (int r = 0; r < 12; ++r)
    for (int c = r; c < 12; ++c, ++k) { At[r * 12 + c] = acc78[k]; At[c * 12 + r] = acc78[k]; }
  double d[12];
  jacobi_svd(At, 12, d, nullptr, 0, 12, 12, 12);
  const double* v[4] = {At + 12 * 11, At + 12 * 10, At + 12 * 9, At + 12 * 8};
  double dv[4][6][3];
  for (int i = 0; i < 4; ++i) {
    int a = 0, b = 1;
    for (int j = 0; j < 6; ++j) {
      dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
      dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
      dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
      b++;
      if (b > 3) { a++; b = a + 1; }
    }
  }
  for (int i = 0; i < 6; ++i) {
    double* row = E.l_6x10 + 10 * i;
    row[0] = epnp_dot(dv[0][i], dv[0][i]);
    row[1] = 2.0f * epnp_dot(dv[0][i], dv[1][i]);
    row[2] = epnp_dot(dv[1][i], dv[1][i]);
    row[3] = 2.0f * epnp_dot(dv[0][i], dv[2][i]);
    row[4] = 2.0f * epnp_dot(dv[1][i], dv[2][i]);
    row[5] = epnp_dot(dv[2][i], dv[2][i]);
    row[6] = 2.0f * epnp_dot(dv[0][i], dv[3][i]);
    row[7] = 2.0f * epnp_dot(dv[1][i], dv[3][i]);
    row[8] = 2.0f * epnp_dot(dv[2][i], dv[3][i]);
    row[9] = epnp_dot(dv[3][i], dv[3][i]);
  }
  E.rho[0] = epnp_dist2(E.cws[0], E.cws[1]);
  E.rho[1] = epnp_dist2(E.cws[0], E.cws[2]);
  E.rho[2] = epnp_dist2(E.cws[0], E.cws[3]);
  E.rho[3] = epnp_dist2(E.cws[1], E.cws[2]);
  E.rho[4] = epnp_dist2(E.cws[1], E.cws[3]);
  E.rho[5] = epnp_dist2(E.cws[2], E.cws[3]);
}

// find_betas_approx_{1,2,3}
P3P_HD void epnp_betas_approx(const Epnp& E, int which, double* betas) {
  const double* L = E.l_6x10;
  betas[0] = betas[1] = betas[2] = betas[3] = 0;
  if (which == 1) {
    double l[24], b4[4];
    for (int i = 0; i < 6; ++i) { l[i * 4] = L[i * 10]; l[i * 4 + 1] = L[i * 10 + 1]; l[i * 4 + 2] = L[i * 10 + 3]; l[i * 4 + 3] = L[i * 10 + 6]; }
    svd_solve(l, 6, 4, E.rho, b4);
    if (b4[0] < 0) {
      betas[0] = sqrt(-b4[0]); betas[1] = -b4[1] / betas[0]; betas[2] = -b4[2] / betas[0]; betas[3] = -b4[3] / betas[0];
    } else {
      betas[0] = sqrt(b4[0]); betas[1] = b4[1] / betas[0]; betas[2] = b4[2] / betas[0]; betas[3] = b4[3] / betas[0];
    }
  } else if (which == 2) {
    double l[18], b3[3];
    for (int i = 0; i < 6; ++i) { l[i * 3] = L[i * 10]; l[i * 3 + 1] = L[i * 10 + 1]; l[i * 3 + 2] = L[i * 10 + 2]; }
    svd_solve(l, 6, 3, E.rho, b3);
    if (b3[0] < 0) {
      betas[0] = sqrt(-b3[0]); betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
    } else {
      betas[0] = sqrt(b3[0]); betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
    }
    if (b3[1] < 0) betas[0] = -betas[0];
    betas[2] = 0.0; betas[3] = 0.0;
  } else {
    double l[30], b5[5];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 5; ++j) l[i * 5 + j] = L[i * 10 + j];
    svd_solve(l, 6, 5, E.rho, b5);
    if (b5[0] < 0) {
      betas[0] = sqrt(-b5[0]); betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
    } else {
      betas[0] = sqrt(b5[0]); betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
    }
    if (b5[1] < 0) betas[0] = -betas[0];
    betas[2] = b5[3] / betas[0];
    betas[3] = 0.0;
  }
}

P3P_HD void epnp_gauss_newton(const Epnp& E, double* betas) {
  double x[4] = {0, 0, 0, 0};
  for (int k = 0; k < 5; ++k) {
    double A[24], b[6];
    for (int i = 0; i < 6; ++i) {
      const double* r = E.l_6x10 + i * 10;
      double* a = A + i * 4;
      a[0] = 2 * r[0] * betas[0] + r[1] * betas[1] + r[3] * betas[2] + r[6] * betas[3];
      a[1] = r[1] * betas[0] + 2 * r[2] * betas[1] + r[4] * betas[2] + r[7] * betas[3];
      a[2] = r[3] * betas[0] + r[4] * betas[1] + 2 * r[5] * betas[2] + r[8] * betas[3];
      a[3] = r[6] * betas[0] + r[7] * betas[1] + r[8] * betas[2] + 2 * r[9] * betas[3];
      b[i] = E.rho[i] - (r[0] * betas[0] * betas[0] + r[1] * betas[0] * betas[1] + r[2] * betas[1] * betas[1] +
                         r[3] * betas[0] * betas[2] + r[4] * betas[1] * betas[2] + r[5] * betas[2] * betas[2] +
                         r[6] * betas[0] * betas[3] + r[7] * betas[1] * betas[3] + r[8] * betas[2] * betas[3] +
                         r[9] * betas[3] * betas[3]);
    }
    qr_solve(A, b, x, 6, 4);   // on failure x keeps the previous step, as in OpenCV
    for (int i = 0; i < 4; ++i) betas[i] += x[i];
  }
}

// compute_ccs / compute_pcs / solve_for_sign / estimate_R_and_t / reprojection_error
P3P_HD double epnp_R_and_t(const Epnp& E, const double* betas, const double* pw, const double* us, int n, double* R, double* t) {
  double ccs[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < 4; ++i) {
    const double* v = E.ut + 12 * (11 - i);
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 3; ++k) ccs[j][k] += betas[i] * v[3 * j + k];
  }
  double a[4];
  epnp_alphas(E, pw, a);
  const double pc0z = a[0] * ccs[0][2] + a[1] * ccs[1][2] + a[2] * ccs[2][2] + a[3] * ccs[3][2];
  if (pc0z < 0.0)
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) ccs[i][j] = -ccs[i][j];
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) {
    epnp_alphas(E, pw + 3 * i, a);
    for (int j = 0; j < 3; ++j) {
      pc0[j] += a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
      pw0[j] += pw[3 * i + j];
    }
  }
  for (int j = 0; j < 3; ++j) { pc0[j] /= n; pw0[j] /= n; }
  double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    epnp_alphas(E, pw + 3 * i, a);
    for (int j = 0; j < 3; ++j) {
      const double pc = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
      abt[3 * j] += (pc - pc0[j]) * (pw[3 * i] - pw0[0]);
      abt[3 * j + 1] += (pc - pc0[j]) * (pw[3 * i + 1] - pw0[1]);
      abt[3 * j + 2] += (pc - pc0[j]) * (pw[3 * i + 2] - pw0[2]);
    }
  }
  double At[9], Vt[9], W[3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) At[j * 3 + i] = abt[i * 3 + j];
  jacobi_svd(At, 3, W, Vt, 3, 3, 3, 3);
  // R = U V^T ; U = At^T (rows of At are the left vectors), V^T = Vt
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = At[0 * 3 + i] * Vt[0 * 3 + j] + At[1 * 3 + i] * Vt[1 * 3 + j] + At[2 * 3 + i] * Vt[2 * 3 + j];
  const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
  if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
  for (int i = 0; i < 3; ++i) t[i] = pc0[i] - (R[3 * i] * pw0[0] + R[3 * i + 1] * pw0[1] + R[3 * i + 2] * pw0[2]);
  double sum2 = 0.0;
  for (int i = 0; i < n; ++i) {
    const double* p = pw + 3 * i;
    const double Xc = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0];
    const double Yc = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
    const double inv_Zc = 1.0 / (R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2]);
    const double ue = Xc * inv_Zc, ve = Yc * inv_Zc;
    const double u = us[2 * i], v = us[2 * i + 1];
    sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
  }
  return sum2 / n;
}

// everything after the null space: three candidates, best reprojection error
P3P_HD void epnp_solve(const Epnp& E, const double* pw, const double* us, int n, double* R, double* t) {
  double Rs[3][9], ts[3][3], err[3];
  for (int c = 0; c < 3; ++c) {
    double betas[4];
    epnp_betas_approx(E, c + 1, betas);
    epnp_gauss_newton(E, betas);
    err[c] = epnp_R_and_t(E, betas, pw, us, n, Rs[c], ts[c]);
  }
  int N = 0;
  if (err[1] < err[0]) N = 1;
  if (err[2] < err[N]) N = 2;
  for (int i = 0; i < 9; ++i) R[i] = Rs[N][i];
  for (int i = 0; i < 3; ++i) t[i] = ts[N][i];
}

// ---------------------------------------------------------------------------------------------------------------
// cv::Rodrigues, matrix -> vector (what solvePnP returns) and vector -> matrix (what the reference applies to it).
P3P_HD void rodrigues_to_vector(const double* Rin, double* r) {
  double At[9], Vt[9], W[3], R[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) At[j * 3 + i] = Rin[i * 3 + j];
  jacobi_svd(At, 3, W, Vt, 3, 3, 3, 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = At[0 * 3 + i] * Vt[0 * 3 + j] + At[1 * 3 + i] * Vt[1 * 3 + j] + At[2 * 3 + i] * Vt[2 * 3 + j];
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      rx = ry = rz = 0;
    } else {
      double t;
      t = (R[0] + 1) * 0.5; rx = sqrt(fmax(t, 0.));
      t = (R[4] + 1) * 0.5; ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5; rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= sqrt(rx * rx + ry * ry + rz * rz);
      rx *= theta; ry *= theta; rz *= theta;
    }
  } else {
    double vth = 1 / (2 * s);
    vth *= theta;
    rx *= vth; ry *= vth; rz *= vth;
  }
  r[0] = rx; r[1] = ry; r[2] = rz;
}

P3P_HD void rodrigues_to_matrix(const double* rv, double* R) {
  double rx = rv[0], ry = rv[1], rz = rv[2];
  const double theta = sqrt(rx * rx + ry * ry + rz * rz);
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = theta ? 1. / theta : 0.;
  rx *= itheta; ry *= itheta; rz *= itheta;
  const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
  const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
  for (int i = 0; i < 9; ++i) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i] + s * r_x[i];
}

// The reference's tail (ezcalibrator.cpp:616-625): Rodrigues(rvec) -> cv2eigen into FLOAT matrices -> cast to double ->
// Sophus::SO3d::fitToSO3 (closest rotation U diag(1,1,det(U V^T)) V^T) -> SE3d storage [qx qy qz qw tx ty tz]
// (Eigen's matrix -> quaternion rule; sign normalised to w >= 0 like the host mirror).
P3P_HD void reference_pose7(const double* rvec, const double* tvec, double* pose7) {
  double R[9], Rf[9];
  rodrigues_to_matrix(rvec, R);
  for (int i = 0; i < 9; ++i) Rf[i] = (double)(float)R[i];
  double At[9], Vt[9], W[3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) At[j * 3 + i] = Rf[i * 3 + j];
  jacobi_svd(At, 3, W, Vt, 3, 3, 3, 3);
  double M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = At[0 * 3 + i] * Vt[0 * 3 + j] + At[1 * 3 + i] * Vt[1 * 3 + j] + At[2 * 3 + i] * Vt[2 * 3 + j];
  const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
  if (det < 0)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[i * 3 + j] -= 2.0 * At[2 * 3 + i] * Vt[2 * 3 + j];
  double q[4];
  const double tr = M[0] + M[4] + M[8];
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    q[0] = (M[7] - M[5]) / s; q[1] = (M[2] - M[6]) / s; q[2] = (M[3] - M[1]) / s; q[3] = 0.25 * s;
  } else if (M[0] > M[4] && M[0] > M[8]) {
    const double s = sqrt(1.0 + M[0] - M[4] - M[8]) * 2;
    q[0] = 0.25 * s; q[1] = (M[1] + M[3]) / s; q[2] = (M[2] + M[6]) / s; q[3] = (M[7] - M[5]) / s;
  } else if (M[4] > M[8]) {
    const double s = sqrt(1.0 + M[4] - M[0] - M[8]) * 2;
    q[0] = (M[1] + M[3]) / s; q[1] = 0.25 * s; q[2] = (M[5] + M[7]) / s; q[3] = (M[2] - M[6]) / s;
  } else {
    const double s = sqrt(1.0 + M[8] - M[0] - M[4]) * 2;
    q[0] = (M[2] + M[6]) / s; q[1] = (M[5] + M[7]) / s; q[2] = 0.25 * s; q[3] = (M[3] - M[1]) / s;
  }
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double sg = q[3] < 0 ? -1.0 : 1.0;
  for (int i = 0; i < 4; ++i) pose7[i] = sg * q[i] / nq;
  for (int i = 0; i < 3; ++i) pose7[4 + i] = (double)(float)tvec[i];
}

// IntrinsicsParam::projImageToCam with Eigen's 3x3 cofactor inverse of K (calib_models.hpp:69-77, 233-240), then the
// .cast<float>() of ezcalibrator.cpp:580.
P3P_HD void normalise_pixel(double fx, double fy, double cx, double cy, float u, float v, float* x, float* y) {
  const double invdet = 1.0 / (fx * fy);
  const double a00 = fy * invdet, a02 = (0.0 * cy - cx * fy) * invdet;
  const double a11 = fx * invdet, a12 = (cx * 0.0 - fx * cy) * invdet;
  *x = (float)((a00 * (double)u + 0.0 * (double)v) + a02);
  *y = (float)((0.0 * (double)u + a11 * (double)v) + a12);
}

P3P_HD bool point_in_image(float u, float v, float wb, float hb) { return u > -0.5f && v > -0.5f && u < wb && v < hb; }

}  // namespace p3p
