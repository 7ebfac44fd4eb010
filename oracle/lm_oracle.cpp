// oracle/lm_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product
// path (ezcalib_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from this file.
//
// CPU restatement ("restated Ceres") of the Levenberg-Marquardt calibration solve that the
// reference drives through Ceres 2.2.0 + Sophus:
//   * problem structure / 3-stage schedule / in-image filter / per-frame RMSE:
//       /root/reference/src/ezcalibrator.cpp:105-291
//   * multi-camera extrinsic solve:        /root/reference/src/ezcalibrator.cpp:631-844
//   * residual functors (14 of them):      /root/reference/include/dist_models/*.hpp
//   * SE3 left-perturbation manifold:      /root/reference/include/ceres_params/se3_param.hpp:12-41
//
// PARITY: the residual DEFINITIONS and the manifold are PINNED -- tests/test_oracle_lm.py checks the 14 restated
// functors (value + Jacobian, <= 1e-14) and se3_plus (<= 1e-15) against the reference's OWN headers compiled
// unmodified into oracle/_ref/liblm_ref.so (oracle/lm_ref/: dist_models/*.hpp, ceres_params/se3_param.hpp against
// stand-in ceres::Jet / AutoDiffCostFunction / AutoDiffManifold, Sophus::SE3, Eigen::Matrix headers).
// The solver RULES remain a restatement ("parity unpinned" for the iterate sequence): Ceres 2.2.0 (pinned by
// /root/reference/build.sh:45), Sophus (unpinned master, build.sh:58) and Eigen are third-party dependencies that
// are NOT in /root/reference nor in this container, and the reference ships no tests / golden vectors.
// What is restated here is Ceres' published trust-region algorithm
// (TrustRegionMinimizer + LevenbergMarquardtStrategy + SchurEliminator + Corrector +
// AutoDiffCostFunction/AutoDiffManifold), see SURVEY.md section 8a row B8 for the rule list.
//
// Build: see oracle/Makefile  (g++ -O3 -fopenmp, no -ffast-math, matching CMakeLists.txt:11-14).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------------
// Forward-mode dual numbers (what ceres::Jet<double,N> is).
// ----------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // ceres jet.h: h = f/g ; dh = (df - h*dg)/g
  Jet<N> h; const double gi = 1.0 / g.a; h.a = f.a * gi;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi; return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { return f * (1.0 / s); }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; h.a = s * gi; const double m = -s * gi * gi;
  for (int i = 0; i < N; ++i) h.v[i] = m * g.v[i]; return h; }
template <int N> inline Jet<N> jsqrt(const Jet<N>& f) { Jet<N> h; h.a = std::sqrt(f.a); const double m = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
template <int N> inline Jet<N> jatan(const Jet<N>& f) { Jet<N> h; h.a = std::atan(f.a); const double m = 1.0 / (1.0 + f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
template <int N> inline Jet<N> jsin(const Jet<N>& f) { Jet<N> h; h.a = std::sin(f.a); const double m = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
template <int N> inline Jet<N> jcos(const Jet<N>& f) { Jet<N> h; h.a = std::cos(f.a); const double m = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jatan(double x) { return std::atan(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jcos(double x) { return std::cos(x); }
inline double jval(double x) { return x; }
template <int N> inline double jval(const Jet<N>& x) { return x.a; }

// ----------------------------------------------------------------------------------------
// Model table: /root/reference/src/camera.cpp:67-94
// ----------------------------------------------------------------------------------------
enum Model { RAD1 = 0, RAD2 = 1, RAD3 = 2, RADTAN4 = 3, RADTAN5 = 4, RADTAN8 = 5, KB4 = 6 };
const int kNumDist[7] = {1, 2, 3, 4, 5, 8, 4};

// Sophus SO3 * point (sophus/so3.hpp, "operator*(point)"): uv = q.vec x p; uv += uv;
// p + q.w*uv + q.vec x uv.  pose storage [qx qy qz qw tx ty tz] (Eigen::Map<const SE3<T>>).
template <class T>
inline void se3_act(const T* Tcw, const double* p, T* out) {
  const T& qx = Tcw[0]; const T& qy = Tcw[1]; const T& qz = Tcw[2]; const T& qw = Tcw[3];
  T uvx = qy * p[2] - qz * p[1];
  T uvy = qz * p[0] - qx * p[2];
  T uvz = qx * p[1] - qy * p[0];
  uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
  out[0] = p[0] + qw * uvx + (qy * uvz - qz * uvy) + Tcw[4];
  out[1] = p[1] + qw * uvy + (qz * uvx - qx * uvz) + Tcw[5];
  out[2] = p[2] + qw * uvz + (qx * uvy - qy * uvx) + Tcw[6];
}

// The 14 functors share this skeleton; e.g. dist_model_k1k2k3p1p2.hpp:111-160 (dual focal)
// and :174-224 (mono focal).  `model` selects the distortion body.
template <class T>
inline void reproj_functor(int model, bool mono, const T* focal, const T* pp, const T* dist,
                           const T* Tcw, const double* wpt, double u, double v, T* err) {
  const T fx = focal[0];
  const T fy = mono ? focal[0] : focal[1];
  const T cx = pp[0];
  const T cy = pp[1];
  T campt[3];
  se3_act(Tcw, wpt, campt);
  const T invz = 1.0 / campt[2];
  const T x = campt[0] * invz;
  const T y = campt[1] * invz;
  T xd, yd;
  switch (model) {
    case RAD1: {  // dist_model_k1.hpp:113-129
      const T x2 = x * x; const T y2 = y * y; const T r2 = x2 + y2;
      const T D = 1.0 + r2 * dist[0];
      xd = x * D; yd = y * D;
      // reference writes fx * x * D + cx (left-to-right); identical up to rounding order.
      err[0] = fx * x * D + cx - u; err[1] = fy * y * D + cy - v; return;
    }
    case RAD2: {  // dist_model_k1k2.hpp:116-132
      const T x2 = x * x; const T y2 = y * y; const T r2 = x2 + y2;
      const T D = 1.0 + r2 * (dist[0] + dist[1] * r2);
      err[0] = fx * x * D + cx - u; err[1] = fy * y * D + cy - v; return;
    }
    case RAD3: {  // dist_model_k1k2k3.hpp:119-136
      const T x2 = x * x; const T y2 = y * y; const T r2 = x2 + y2; const T r4 = r2 * r2;
      const T D = 1.0 + r2 * (dist[0] + dist[1] * r2 + dist[2] * r4);
      err[0] = fx * x * D + cx - u; err[1] = fy * y * D + cy - v; return;
    }
    case RADTAN4: {  // dist_model_k1k2p1p2.hpp:132-152
      const T x2 = x * x; const T y2 = y * y; const T xy_2 = 2.0 * x * y; const T r2 = x2 + y2;
      const T D = 1.0 + r2 * (dist[0] + dist[1] * r2);
      const T& p1 = dist[2]; const T& p2 = dist[3];
      xd = x * D + p1 * xy_2 + p2 * (r2 + 2.0 * x2);
      yd = y * D + p2 * xy_2 + p1 * (r2 + 2.0 * y2);
      break;
    }
    case RADTAN5: {  // dist_model_k1k2k3p1p2.hpp:135-156 ; order [k1,k2,k3,p1,p2]
      const T x2 = x * x; const T y2 = y * y; const T xy_2 = 2.0 * x * y; const T r2 = x2 + y2; const T r4 = r2 * r2;
      const T D = 1.0 + r2 * (dist[0] + dist[1] * r2 + dist[2] * r4);
      const T& p1 = dist[3]; const T& p2 = dist[4];
      xd = x * D + p1 * xy_2 + p2 * (r2 + 2.0 * x2);
      yd = y * D + p2 * xy_2 + p1 * (r2 + 2.0 * y2);
      break;
    }
    case RADTAN8: {  // dist_model_radtan_full.hpp:148-169 ; order [k1..k6,p1,p2]
      const T x2 = x * x; const T y2 = y * y; const T xy_2 = 2.0 * x * y; const T r2 = x2 + y2; const T r4 = r2 * r2;
      const T D = (1.0 + r2 * (dist[0] + dist[1] * r2 + dist[2] * r4)) /
                  (1.0 + r2 * (dist[3] + dist[4] * r2 + dist[5] * r4));
      const T& p1 = dist[6]; const T& p2 = dist[7];
      xd = x * D + p1 * xy_2 + p2 * (r2 + 2.0 * x2);
      yd = y * D + p2 * xy_2 + p1 * (r2 + 2.0 * y2);
      break;
    }
    case KB4: default: {  // dist_model_kb.hpp:138-160 ; r == 0 -> 0/0 NaN exactly like the reference
      const T r2 = x * x + y * y;
      const T r = jsqrt(r2);
      const T theta = jatan(r);
      const T theta2 = theta * theta; const T theta4 = theta2 * theta2; const T theta6 = theta4 * theta2;
      const T D = theta * (1.0 + theta2 * (dist[0] + theta2 * dist[1] + theta4 * dist[2] + theta6 * dist[3]));
      const T D_r = D / r;
      xd = D_r * x; yd = D_r * y;
      break;
    }
  }
  err[0] = fx * xd + cx - u;
  err[1] = fy * yd + cy - v;
}

// DistParam::distortCamPoint(double,double) + IntrinsicsParam::projCamToImage in plain doubles
// (used for the per-frame RMSE: /root/reference/src/ezcalibrator.cpp:236-274).
inline void project_double(int model, bool mono, const double* focal, const double* pp, const double* dist,
                           const double* Tcw, const double* wpt, double* px) {
  double e[2];
  reproj_functor<double>(model, mono, focal, pp, dist, Tcw, wpt, 0.0, 0.0, e);
  px[0] = e[0]; px[1] = e[1];
}

// ----------------------------------------------------------------------------------------
// Sophus SE3 pieces (sophus/so3.hpp expAndTheta, se3.hpp exp, group product with the
// normalising SO3 constructor) -- restated, generic in T so the manifold's PlusJacobian can
// be obtained with Jets exactly like ceres::AutoDiffManifold does.
// ----------------------------------------------------------------------------------------
const double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()

template <class T>
inline void quat_rotate(const T* q, const T* p, T* out) {  // q = [x y z w]
  T uvx = q[1] * p[2] - q[2] * p[1];
  T uvy = q[2] * p[0] - q[0] * p[2];
  T uvz = q[0] * p[1] - q[1] * p[0];
  uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
  out[0] = p[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
  out[1] = p[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
  out[2] = p[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}

template <class T>
inline void quat_to_matrix(const T* q, T R[3][3]) {  // Eigen::Quaternion::toRotationMatrix
  const T tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0][0] = 1.0 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1.0 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1.0 - (txx + tyy);
}

// exp(delta) * T ; delta = (upsilon, omega).  se3_param.hpp:14-25.
template <class T>
inline void se3_plus(const double* x, const T* delta, T* out) {
  const T* ups = delta; const T* om = delta + 3;
  // SO3::expAndTheta
  const T theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  T theta, imag, real;
  const bool small = jval(theta_sq) < kSophusEps * kSophusEps;
  if (small) {
    theta = T(0.0);
    const T theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    theta = jsqrt(theta_sq);
    const T half = 0.5 * theta;
    imag = jsin(half) / theta;
    real = jcos(half);
  }
  T qd[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  // SE3::exp : V
  T Om[3][3] = {{T(0.0), -om[2], om[1]}, {om[2], T(0.0), -om[0]}, {-om[1], om[0], T(0.0)}};
  T V[3][3];
  if (jval(theta) < kSophusEps) {
    quat_to_matrix(qd, V);
  } else {
    T Om2[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T s(0.0); for (int k = 0; k < 3; ++k) s = s + Om[i][k] * Om[k][j]; Om2[i][j] = s; }
    const T th2 = theta * theta;
    const T c1 = (1.0 - jcos(theta)) / th2;
    const T c2 = (theta - jsin(theta)) / (th2 * theta);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j ? 1.0 : 0.0) + c1 * Om[i][j] + c2 * Om2[i][j];
  }
  T td[3];
  for (int i = 0; i < 3; ++i) td[i] = V[i][0] * ups[0] + V[i][1] * ups[1] + V[i][2] * ups[2];
  // group product: q' = normalize(qd (x) q) ; t' = td + qd * t
  const double ax = x[0], ay = x[1], az = x[2], aw = x[3];
  T q[4];
  q[3] = qd[3] * aw - qd[0] * ax - qd[1] * ay - qd[2] * az;
  q[0] = qd[3] * ax + qd[0] * aw + qd[1] * az - qd[2] * ay;
  q[1] = qd[3] * ay + qd[1] * aw + qd[2] * ax - qd[0] * az;
  q[2] = qd[3] * az + qd[2] * aw + qd[0] * ay - qd[1] * ax;
  const T nrm = jsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) out[i] = q[i] / nrm;
  const T tt[3] = {T(x[4]), T(x[5]), T(x[6])};
  T rt[3];
  quat_rotate(qd, tt, rt);
  for (int i = 0; i < 3; ++i) out[4 + i] = td[i] + rt[i];
}

// ceres::AutoDiffManifold<.,7,6>::PlusJacobian: d Plus(x,delta) / d delta at delta = 0 (7x6 row-major).
inline void se3_plus_jacobian(const double* x, double* Jp) {
  Jet<6> d[6];
  for (int i = 0; i < 6; ++i) d[i] = Jet<6>(0.0, i);
  Jet<6> out[7];
  se3_plus<Jet<6>>(x, d, out);
  for (int r = 0; r < 7; ++r) for (int c = 0; c < 6; ++c) Jp[r * 6 + c] = out[r].v[c];
}

// ----------------------------------------------------------------------------------------
// One residual block: value + Jacobians w.r.t. (focal | pp | dist | pose tangent[6]).
// Columns of Jfull: [focal(nf) pp(2) dist(nd) pose(6)], raw (un-robustified).
// ----------------------------------------------------------------------------------------
template <int NF, int ND>
inline void eval_block_t(int model, const double* focal, const double* pp, const double* dist, const double* pose,
                         const double* plusJ, const double* wpt, double u, double v, double* r, double* Jfull) {
  constexpr int N = NF + 2 + ND + 7;
  typedef Jet<N> J;
  J jf[2], jpp[2], jd[8], jT[7];
  int k = 0;
  for (int i = 0; i < NF; ++i) jf[i] = J(focal[i], k++);
  for (int i = 0; i < 2; ++i) jpp[i] = J(pp[i], k++);
  for (int i = 0; i < ND; ++i) jd[i] = J(dist[i], k++);
  for (int i = 0; i < 7; ++i) jT[i] = J(pose[i], k++);
  J e[2];
  reproj_functor<J>(model, NF == 1, jf, jpp, jd, jT, wpt, u, v, e);
  constexpr int KS = NF + 2 + ND;
  constexpr int D = KS + 6;
  for (int rr = 0; rr < 2; ++rr) {
    r[rr] = e[rr].a;
    for (int c = 0; c < KS; ++c) Jfull[rr * D + c] = e[rr].v[c];
    // ambient 2x7 times PlusJacobian 7x6
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
      for (int a = 0; a < 7; ++a) s += e[rr].v[KS + a] * plusJ[a * 6 + c];
      Jfull[rr * D + KS + c] = s;
    }
  }
}

typedef void (*EvalBlockFn)(int, const double*, const double*, const double*, const double*, const double*,
                            const double*, double, double, double*, double*);

EvalBlockFn pick_eval(int model, bool mono) {
  switch (kNumDist[model]) {
    case 1: return mono ? eval_block_t<1, 1> : eval_block_t<2, 1>;
    case 2: return mono ? eval_block_t<1, 2> : eval_block_t<2, 2>;
    case 3: return mono ? eval_block_t<1, 3> : eval_block_t<2, 3>;
    case 4: return mono ? eval_block_t<1, 4> : eval_block_t<2, 4>;
    case 5: return mono ? eval_block_t<1, 5> : eval_block_t<2, 5>;
    default: return mono ? eval_block_t<1, 8> : eval_block_t<2, 8>;
  }
}

// ceres::CauchyLoss(1.0) + Corrector (rho'' < 0 => alpha = 0 => scale r and J by sqrt(rho')).
inline void cauchy(double s, double* rho0, double* sqrt_rho1) {
  const double sum = 1.0 + s;
  const double inv = 1.0 / sum;
  *rho0 = std::log(sum);
  *sqrt_rho1 = std::sqrt(std::max(std::numeric_limits<double>::min(), inv));
}

// Dense helpers ---------------------------------------------------------------------------
// In-place Cholesky (lower) of n x n row-major SPD matrix. Returns false if not PD.
bool cholesky(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; }
}

}  // namespace

// ============================================================================================
// Public C interface of the oracle
// ============================================================================================
extern "C" {

struct OracleLmProblem {
  int32_t model;        // 0..6
  int32_t mono;         // use_mono_focal
  int32_t n_blocks;     // number of SE3 blocks ("views"; extrinsics in the multi-camera solve)
  int32_t obs_per_block;  // observations per block (n_pts for mono)
  int32_t pts_period;   // 3-D point for local obs index l is pts[(l % pts_period)]
  int32_t shared_per_block;  // 0: one shared (focal,pp,dist) set; 1: one set per block (multi-cam; all constant)
  double img_w, img_h;  // isPointInImage bounds (calib_models.hpp:107-115, float overload)
  const double* pts;    // [pts_period][3]
  const float* obs;     // [n_blocks][obs_per_block][2]
  int32_t num_threads;
};

struct OracleLmOptions {
  int32_t opt_focal, opt_pp, opt_dist;  // which shared blocks are variable
  int32_t max_iterations;               // 1000
  int32_t trace_capacity;               // entries available in the trace arrays
};

struct OracleLmSummary {
  int32_t iterations;        // number of iteration summaries (incl. iteration 0)
  int32_t num_successful;    // incl. iteration 0, like Ceres
  int32_t termination;       // 0 convergence(function tol) 1 convergence(parameter tol) 2 gradient tol
                             // 3 no convergence (max iter) 4 min radius 5 failure
  int32_t n_residual_blocks;
  double initial_cost, final_cost;
};

static inline bool in_image(float u, float v, float wb, float hb) {
  return (u > -0.5f && v > -0.5f && u < wb && v < hb);
}

// Evaluate a single residual block (test hook): raw residual + raw Jacobian
// columns [focal(nf) pp(2) dist(nd) pose(6)].
void oracle_residual_jacobian(int model, int mono, const double* focal, const double* pp, const double* dist,
                              const double* pose, const double* wpt, double u, double v, double* r, double* J) {
  double plusJ[42];
  se3_plus_jacobian(pose, plusJ);
  pick_eval(model, mono != 0)(model, focal, pp, dist, pose, plusJ, wpt, u, v, r, J);
}

void oracle_se3_plus(const double* x, const double* delta, double* out) { se3_plus<double>(x, delta, out); }

void oracle_project(int model, int mono, const double* focal, const double* pp, const double* dist,
                    const double* pose, const double* wpt, double* px) {
  project_double(model, mono != 0, focal, pp, dist, pose, wpt, px);
}

namespace {

struct Evaluator {
  const OracleLmProblem& P;
  int nf, nd, kfull, D;
  int nthreads;
  std::vector<int> obs_valid;  // compact list of valid obs (block-major)
  std::vector<int> block_start;  // per block: range in obs_valid
  EvalBlockFn fn;

  explicit Evaluator(const OracleLmProblem& p) : P(p) {
    nf = p.mono ? 1 : 2; nd = kNumDist[p.model]; kfull = nf + 2 + nd; D = kfull + 6;
    fn = pick_eval(p.model, p.mono != 0);
    nthreads = p.num_threads > 0 ? p.num_threads : 1;
    const float wb = static_cast<float>(p.img_w) - 0.5f, hb = static_cast<float>(p.img_h) - 0.5f;
    block_start.assign(p.n_blocks + 1, 0);
    for (int b = 0; b < p.n_blocks; ++b) {
      block_start[b] = static_cast<int>(obs_valid.size());
      for (int l = 0; l < p.obs_per_block; ++l) {
        const float* o = p.obs + (static_cast<size_t>(b) * p.obs_per_block + l) * 2;
        if (in_image(o[0], o[1], wb, hb)) obs_valid.push_back(l);
      }
    }
    block_start[p.n_blocks] = static_cast<int>(obs_valid.size());
  }
  size_t num_obs() const { return obs_valid.size(); }
  const double* shared(const double* base, int n, int b) const { return P.shared_per_block ? base + static_cast<size_t>(b) * n : base; }

  // cost only (candidate evaluation). Returns false when a residual is non-finite.
  bool cost(const double* focal, const double* pp, const double* dist, const double* poses, double* out) const {
    double total = 0.0; bool ok = true;
    std::vector<double> part(P.n_blocks, 0.0);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int b = 0; b < P.n_blocks; ++b) {
      double c = 0.0;
      const double* pose = poses + static_cast<size_t>(b) * 7;
      for (int i = block_start[b]; i < block_start[b + 1]; ++i) {
        const int l = obs_valid[i];
        const float* o = P.obs + (static_cast<size_t>(b) * P.obs_per_block + l) * 2;
        double e[2];
        reproj_functor<double>(P.model, P.mono != 0, shared(focal, nf, b), shared(pp, 2, b), shared(dist, nd, b), pose,
                               P.pts + static_cast<size_t>(l % P.pts_period) * 3, o[0], o[1], e);
        const double s = e[0] * e[0] + e[1] * e[1];
        c += 0.5 * std::log(1.0 + s);
      }
      part[b] = c;
    }
    for (int b = 0; b < P.n_blocks; ++b) total += part[b];
    if (!std::isfinite(total)) ok = false;
    *out = total;
    return ok;
  }

  // Full evaluation: cost, robustified residuals f (2 per obs) and Jacobian rows Jrows (2 x D per obs,
  // columns [focal pp dist | pose6]).
  bool full(const double* focal, const double* pp, const double* dist, const double* poses, double* cost_out,
            std::vector<double>& f, std::vector<double>& Jrows) const {
    const size_t n = num_obs();
    f.resize(n * 2); Jrows.resize(n * 2 * D);
    std::vector<double> part(P.n_blocks, 0.0);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int b = 0; b < P.n_blocks; ++b) {
      const double* pose = poses + static_cast<size_t>(b) * 7;
      double plusJ[42];
      se3_plus_jacobian(pose, plusJ);
      double c = 0.0;
      for (int i = block_start[b]; i < block_start[b + 1]; ++i) {
        const int l = obs_valid[i];
        const float* o = P.obs + (static_cast<size_t>(b) * P.obs_per_block + l) * 2;
        double r[2]; double* J = &Jrows[static_cast<size_t>(i) * 2 * D];
        fn(P.model, shared(focal, nf, b), shared(pp, 2, b), shared(dist, nd, b), pose, plusJ,
           P.pts + static_cast<size_t>(l % P.pts_period) * 3, o[0], o[1], r, J);
        const double s = r[0] * r[0] + r[1] * r[1];
        double rho0, w;
        cauchy(s, &rho0, &w);
        c += 0.5 * rho0;
        f[2 * i] = w * r[0]; f[2 * i + 1] = w * r[1];
        for (int k = 0; k < 2 * D; ++k) J[k] *= w;
      }
      part[b] = c;
    }
    double total = 0.0;
    for (int b = 0; b < P.n_blocks; ++b) total += part[b];
    *cost_out = total;
    return std::isfinite(total);
  }
};

}  // namespace

// One ceres::Solve() on the problem with the given variable/constant pattern.
// focal/pp/dist/poses are updated in place (like Ceres writing through the user's pointers).
// trace_* (optional, may be null): per iteration summary cost / trust-region radius / step flag
// (0 = iteration zero, 1 successful, 2 unsuccessful, 3 invalid).
int oracle_lm_solve(const OracleLmProblem* prob, const OracleLmOptions* opt, double* focal, double* pp, double* dist,
                    double* poses, OracleLmSummary* summary, double* trace_cost, double* trace_radius,
                    int32_t* trace_flag) {
  const OracleLmProblem& P = *prob;
  Evaluator ev(P);
  const int nf = ev.nf, nd = ev.nd, D = ev.D, kfull = ev.kfull;
  const int nb = P.n_blocks;
  // active shared columns (ceres removes constant blocks from the reduced program)
  std::vector<int> cols;
  if (!P.shared_per_block) {
    if (opt->opt_focal) for (int i = 0; i < nf; ++i) cols.push_back(i);
    if (opt->opt_pp) for (int i = 0; i < 2; ++i) cols.push_back(nf + i);
    if (opt->opt_dist) for (int i = 0; i < nd; ++i) cols.push_back(nf + 2 + i);
  }
  const int k = static_cast<int>(cols.size());
  // Blocks without residuals are removed from the program (Program::RemoveFixedBlocks).
  std::vector<char> used(nb, 0);
  for (int b = 0; b < nb; ++b) used[b] = ev.block_start[b + 1] > ev.block_start[b];

  summary->n_residual_blocks = static_cast<int>(ev.num_obs());
  const int ncols = k + 6 * nb;  // column layout: [shared k | pose blocks 6 each]

  std::vector<double> f, Jrows;
  std::vector<double> scale(ncols, 1.0), diagonal(ncols, 0.0), lm_diag(ncols, 0.0);
  std::vector<double> step(ncols), delta(ncols), gradient(ncols);
  std::vector<double> cand_focal(focal, focal + (P.shared_per_block ? nb : 1) * nf);
  std::vector<double> cand_pp(pp, pp + (P.shared_per_block ? nb : 1) * 2);
  std::vector<double> cand_dist(dist, dist + (P.shared_per_block ? nb : 1) * nd);
  std::vector<double> cand_poses(poses, poses + static_cast<size_t>(nb) * 7);

  auto shared_ptr = [&](int c) -> double* {  // active shared column -> parameter address
    const int j = cols[c];
    if (j < nf) return focal + j;
    if (j < nf + 2) return pp + (j - nf);
    return dist + (j - nf - 2);
  };
  auto x_norm_fn = [&]() {
    double s = 0.0;
    for (int c = 0; c < k; ++c) { const double v = *shared_ptr(c); s += v * v; }
    for (int b = 0; b < nb; ++b) if (used[b]) for (int i = 0; i < 7; ++i) { const double v = poses[b * 7 + i]; s += v * v; }
    return std::sqrt(s);
  };

  double x_cost = 0.0, x_norm = 0.0;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  const double min_diag = 1e-6, max_diag = 1e32, max_radius = 1e16, min_radius = 1e-32;
  const double function_tol = 1e-6, gradient_tol = 1e-10, parameter_tol = 1e-8, min_rel_decrease = 1e-3;
  int num_consecutive_invalid = 0;
  int iter = 0, n_success = 0;
  int termination = 3;
  double grad_max_norm = 0.0;

  auto col_of = [&](int j) -> int {  // full shared column j -> active index or -1
    for (int c = 0; c < k; ++c) if (cols[c] == j) return c;
    return -1;
  };
  std::vector<int> colmap(kfull, -1);
  for (int j = 0; j < kfull; ++j) colmap[j] = col_of(j);

  // Evaluate gradient + Jacobian at the current x (TrustRegionMinimizer::EvaluateGradientAndJacobian)
  auto eval_grad_jac = [&](bool first) -> bool {
    if (!ev.full(focal, pp, dist, poses, &x_cost, f, Jrows)) return false;
    // gradient of the *unscaled* Jacobian: g = J^T f
    std::fill(gradient.begin(), gradient.end(), 0.0);
    std::vector<double> colsq(ncols, 0.0);
    for (int b = 0; b < nb; ++b) {
      for (int i = ev.block_start[b]; i < ev.block_start[b + 1]; ++i) {
        for (int rr = 0; rr < 2; ++rr) {
          const double* J = &Jrows[(static_cast<size_t>(i) * 2 + rr) * D];
          const double fr = f[2 * i + rr];
          for (int j = 0; j < kfull; ++j) { const int c = colmap[j]; if (c >= 0) { gradient[c] += J[j] * fr; colsq[c] += J[j] * J[j]; } }
          for (int j = 0; j < 6; ++j) { gradient[k + 6 * b + j] += J[kfull + j] * fr; colsq[k + 6 * b + j] += J[kfull + j] * J[kfull + j]; }
        }
      }
    }
    if (first) for (int c = 0; c < ncols; ++c) scale[c] = 1.0 / (1.0 + std::sqrt(colsq[c]));
    // jacobian_->ScaleColumns(scale) : applied lazily below through `scale`.
    // gradient_max_norm = | x - Plus(x, -g) |_inf  (ambient)
    double gmax = 0.0;
    for (int c = 0; c < k; ++c) gmax = std::max(gmax, std::fabs(gradient[c]));
    for (int b = 0; b < nb; ++b) {
      if (!used[b]) continue;
      double ng[6], out[7];
      for (int j = 0; j < 6; ++j) ng[j] = -gradient[k + 6 * b + j];
      se3_plus<double>(poses + b * 7, ng, out);
      for (int j = 0; j < 7; ++j) gmax = std::max(gmax, std::fabs(poses[b * 7 + j] - out[j]));
    }
    grad_max_norm = gmax;
    return true;
  };

  auto record = [&](double cost, int flag) {
    if (iter < opt->trace_capacity) {
      if (trace_cost) trace_cost[iter] = cost;
      if (trace_radius) trace_radius[iter] = radius;
      if (trace_flag) trace_flag[iter] = flag;
    }
  };

  // ---- iteration zero
  x_norm = x_norm_fn();
  if (!eval_grad_jac(true)) {
    summary->iterations = 0; summary->num_successful = 0; summary->termination = 5;
    summary->initial_cost = summary->final_cost = x_cost;
    return 1;
  }
  summary->initial_cost = x_cost;
  record(x_cost, 0);
  n_success = 1;
  bool last_successful = false;  // Ceres: IterationZero sets step_is_successful = false

  std::vector<double> S(static_cast<size_t>(k) * k), rhs(k), yc(k);
  std::vector<double> Vinv(static_cast<size_t>(nb) * 36), Wb(static_cast<size_t>(nb) * 6 * std::max(k, 1)), gpb(static_cast<size_t>(nb) * 6);

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iter >= opt->max_iterations) { termination = 3; break; }
    if (last_successful && grad_max_norm <= gradient_tol) { termination = 2; break; }
    if (radius <= min_radius) { termination = 4; break; }
    ++iter;

    // ---- LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      std::fill(diagonal.begin(), diagonal.end(), 0.0);
      for (int b = 0; b < nb; ++b)
        for (int i = ev.block_start[b]; i < ev.block_start[b + 1]; ++i)
          for (int rr = 0; rr < 2; ++rr) {
            const double* J = &Jrows[(static_cast<size_t>(i) * 2 + rr) * D];
            for (int j = 0; j < kfull; ++j) { const int c = colmap[j]; if (c >= 0) { const double v = J[j] * scale[c]; diagonal[c] += v * v; } }
            for (int j = 0; j < 6; ++j) { const int c = k + 6 * b + j; const double v = J[kfull + j] * scale[c]; diagonal[c] += v * v; }
          }
      for (int c = 0; c < ncols; ++c) diagonal[c] = std::min(std::max(diagonal[c], min_diag), max_diag);
    }
    for (int c = 0; c < ncols; ++c) lm_diag[c] = std::sqrt(diagonal[c] / radius);

    // ---- SchurEliminator::Eliminate (poses are the e-blocks) on the scaled Jacobian
    std::fill(S.begin(), S.end(), 0.0); std::fill(rhs.begin(), rhs.end(), 0.0);
    for (int c = 0; c < k; ++c) S[c * k + c] = lm_diag[c] * lm_diag[c];
    bool solver_ok = true;
    {
      const int nt = ev.nthreads;
      std::vector<std::vector<double>> Sp(nt, std::vector<double>(static_cast<size_t>(k) * k + k, 0.0));
      std::vector<char> okp(nt, 1);
#pragma omp parallel for schedule(static) num_threads(nt)
      for (int b = 0; b < nb; ++b) {
        if (!used[b]) continue;
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* Sl = Sp[tid].data(); double* rl = Sl + static_cast<size_t>(k) * k;
        double ete[36] = {0}, etb[6] = {0};
        std::vector<double> etf(6 * std::max(k, 1), 0.0);
        for (int i = ev.block_start[b]; i < ev.block_start[b + 1]; ++i)
          for (int rr = 0; rr < 2; ++rr) {
            const double* J = &Jrows[(static_cast<size_t>(i) * 2 + rr) * D];
            const double fr = f[2 * i + rr];
            double e[6], fc[12];
            for (int j = 0; j < 6; ++j) e[j] = J[kfull + j] * scale[k + 6 * b + j];
            for (int c = 0; c < k; ++c) fc[c] = J[cols[c]] * scale[c];
            for (int a = 0; a < 6; ++a) { for (int c2 = 0; c2 < 6; ++c2) ete[a * 6 + c2] += e[a] * e[c2]; etb[a] += e[a] * fr; for (int c = 0; c < k; ++c) etf[a * k + c] += e[a] * fc[c]; }
            for (int c = 0; c < k; ++c) { rl[c] += fc[c] * fr; for (int c2 = 0; c2 < k; ++c2) Sl[c * k + c2] += fc[c] * fc[c2]; }
          }
        for (int a = 0; a < 6; ++a) ete[a * 6 + a] += lm_diag[k + 6 * b + a] * lm_diag[k + 6 * b + a];
        // inverse_ete = InvertPSDMatrix(ete) via LLT.solve(Identity)
        double L[36]; std::memcpy(L, ete, sizeof(L));
        if (!cholesky(L, 6)) { okp[tid] = 0; continue; }
        double* inv = &Vinv[static_cast<size_t>(b) * 36];
        for (int c = 0; c < 6; ++c) { double col[6] = {0}; col[c] = 1.0; chol_solve(L, 6, col); for (int a = 0; a < 6; ++a) inv[a * 6 + c] = col[a]; }
        // rhs -= F^T E inv g ; lhs -= (E^T F)^T inv (E^T F)
        double ig[6];
        for (int a = 0; a < 6; ++a) { double s = 0; for (int c = 0; c < 6; ++c) s += inv[a * 6 + c] * etb[c]; ig[a] = s; }
        std::vector<double> iW(6 * std::max(k, 1));
        for (int a = 0; a < 6; ++a) for (int c = 0; c < k; ++c) { double s = 0; for (int m = 0; m < 6; ++m) s += inv[a * 6 + m] * etf[m * k + c]; iW[a * k + c] = s; }
        for (int c = 0; c < k; ++c) {
          double s = 0; for (int a = 0; a < 6; ++a) s += etf[a * k + c] * ig[a];
          rl[c] -= s;
          for (int c2 = 0; c2 < k; ++c2) { double t = 0; for (int a = 0; a < 6; ++a) t += etf[a * k + c] * iW[a * k + c2]; Sl[c * k + c2] -= t; }
        }
        for (int a = 0; a < 6; ++a) { gpb[b * 6 + a] = etb[a]; for (int c = 0; c < k; ++c) Wb[(static_cast<size_t>(b) * 6 + a) * k + c] = etf[a * k + c]; }
      }
      for (int t = 0; t < nt; ++t) {
        if (!okp[t]) solver_ok = false;
        for (int i = 0; i < k * k; ++i) S[i] += Sp[t][i];
        for (int c = 0; c < k; ++c) rhs[c] += Sp[t][static_cast<size_t>(k) * k + c];
      }
    }
    // reduced system
    if (solver_ok && k > 0) {
      std::vector<double> Ls(S);
      if (!cholesky(Ls.data(), k)) solver_ok = false;
      else { yc = rhs; chol_solve(Ls.data(), k, yc.data()); }
    }
    // back substitution; Ceres solves J y = f then step = -y
    if (solver_ok) {
      for (int c = 0; c < k; ++c) step[c] = -yc[c];
      for (int b = 0; b < nb; ++b) {
        if (!used[b]) { for (int a = 0; a < 6; ++a) step[k + 6 * b + a] = 0.0; continue; }
        double t[6];
        for (int a = 0; a < 6; ++a) { double s = gpb[b * 6 + a]; for (int c = 0; c < k; ++c) s -= Wb[(static_cast<size_t>(b) * 6 + a) * k + c] * yc[c]; t[a] = s; }
        const double* inv = &Vinv[static_cast<size_t>(b) * 36];
        for (int a = 0; a < 6; ++a) { double s = 0; for (int c = 0; c < 6; ++c) s += inv[a * 6 + c] * t[c]; step[k + 6 * b + a] = -s; }
      }
      for (int c = 0; c < ncols; ++c) if (!std::isfinite(step[c])) solver_ok = false;
    }
    reuse_diagonal = true;

    bool step_valid = false;
    double model_cost_change = 0.0;
    if (solver_ok) {
      // model_cost_change = -(J step)^T (f + J step / 2) with the scaled Jacobian
      double acc = 0.0;
      for (int b = 0; b < nb; ++b)
        for (int i = ev.block_start[b]; i < ev.block_start[b + 1]; ++i)
          for (int rr = 0; rr < 2; ++rr) {
            const double* J = &Jrows[(static_cast<size_t>(i) * 2 + rr) * D];
            double m = 0.0;
            for (int c = 0; c < k; ++c) m += J[cols[c]] * scale[c] * step[c];
            for (int j = 0; j < 6; ++j) { const int c = k + 6 * b + j; m += J[kfull + j] * scale[c] * step[c]; }
            acc += m * (f[2 * i + rr] + m / 2.0);
          }
      model_cost_change = -acc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      // HandleInvalidStep
      ++num_consecutive_invalid;
      if (num_consecutive_invalid >= 5) { record(x_cost, 3); termination = 5; break; }
      // LevenbergMarquardtStrategy::StepIsInvalid() == StepRejected(0.0)
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      record(x_cost, 3);
      last_successful = false;
      continue;
    }
    num_consecutive_invalid = 0;
    for (int c = 0; c < ncols; ++c) delta[c] = step[c] * scale[c];

    // candidate = Plus(x, delta)
    std::copy(focal, focal + (P.shared_per_block ? nb : 1) * nf, cand_focal.begin());
    std::copy(pp, pp + (P.shared_per_block ? nb : 1) * 2, cand_pp.begin());
    std::copy(dist, dist + (P.shared_per_block ? nb : 1) * nd, cand_dist.begin());
    for (int c = 0; c < k; ++c) {
      const int j = cols[c];
      if (j < nf) cand_focal[j] = focal[j] + delta[c];
      else if (j < nf + 2) cand_pp[j - nf] = pp[j - nf] + delta[c];
      else cand_dist[j - nf - 2] = dist[j - nf - 2] + delta[c];
    }
    for (int b = 0; b < nb; ++b) {
      if (used[b]) se3_plus<double>(poses + b * 7, &delta[k + 6 * b], &cand_poses[b * 7]);
      else std::copy(poses + b * 7, poses + b * 7 + 7, &cand_poses[b * 7]);
    }
    double cand_cost;
    if (!ev.cost(cand_focal.data(), cand_pp.data(), cand_dist.data(), cand_poses.data(), &cand_cost))
      cand_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached (ambient step norm)
    double sn = 0.0;
    for (int c = 0; c < k; ++c) sn += delta[c] * delta[c];
    for (int b = 0; b < nb; ++b) if (used[b]) for (int i = 0; i < 7; ++i) { const double d = poses[b * 7 + i] - cand_poses[b * 7 + i]; sn += d * d; }
    const double step_norm = std::sqrt(sn);
    if (step_norm <= parameter_tol * (x_norm + parameter_tol)) { record(x_cost, 2); termination = 1; break; }
    // FunctionToleranceReached
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= function_tol * x_cost) { record(x_cost, 2); termination = 0; break; }

    double rel_decrease;
    if (cand_cost >= std::numeric_limits<double>::max()) rel_decrease = std::numeric_limits<double>::lowest();
    else rel_decrease = cost_change / model_cost_change;
    if (rel_decrease > min_rel_decrease) {
      // HandleSuccessfulStep
      std::copy(cand_focal.begin(), cand_focal.end(), focal);
      std::copy(cand_pp.begin(), cand_pp.end(), pp);
      std::copy(cand_dist.begin(), cand_dist.end(), dist);
      std::copy(cand_poses.begin(), cand_poses.end(), poses);
      x_norm = x_norm_fn();
      if (!eval_grad_jac(false)) { termination = 5; break; }
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel_decrease - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      ++n_success;
      last_successful = true;
      record(x_cost, 1);
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      last_successful = false;
      record(x_cost, 2);
    }
  }
  summary->iterations = iter + 1;
  summary->num_successful = n_success;
  summary->termination = termination;
  summary->final_cost = x_cost;
  return 0;
}

// Per-block RMSE of the plain (non-robust) reprojection error over in-image observations:
// /root/reference/src/ezcalibrator.cpp:236-274.  Blocks with no observation get DBL_MAX.
void oracle_frame_rmse(const OracleLmProblem* prob, const double* focal, const double* pp, const double* dist,
                       const double* poses, double* rmse) {
  const OracleLmProblem& P = *prob;
  const int nf = P.mono ? 1 : 2, nd = kNumDist[P.model];
  const float wb = static_cast<float>(P.img_w) - 0.5f, hb = static_cast<float>(P.img_h) - 0.5f;
  for (int b = 0; b < P.n_blocks; ++b) {
    double acc = 0.0; int n = 0;
    for (int l = 0; l < P.obs_per_block; ++l) {
      const float* o = P.obs + (static_cast<size_t>(b) * P.obs_per_block + l) * 2;
      if (!in_image(o[0], o[1], wb, hb)) continue;
      double px[2];
      project_double(P.model, P.mono != 0, P.shared_per_block ? focal + b * nf : focal, P.shared_per_block ? pp + b * 2 : pp,
                     P.shared_per_block ? dist + b * nd : dist, poses + static_cast<size_t>(b) * 7,
                     P.pts + static_cast<size_t>(l % P.pts_period) * 3, px);
      const double dx = static_cast<double>(o[0]) - px[0], dy = static_cast<double>(o[1]) - px[1];
      const double nrm = std::sqrt(dx * dx + dy * dy);  // (px_obs - px_proj).norm()
      acc += nrm * nrm; ++n;
    }
    rmse[b] = n > 0 ? std::sqrt(acc / static_cast<double>(n)) : std::numeric_limits<double>::max();
  }
}

// The reference's 3-stage schedule (/root/reference/src/ezcalibrator.cpp:192-228):
//   1. focal + poses   2. + distortion   3. + principal point ; then per-frame RMSE.
int oracle_lm_mono(const OracleLmProblem* prob, double* focal, double* pp, double* dist, double* poses,
                   double* frame_rmse, OracleLmSummary* summaries /*[3]*/, int32_t max_iterations) {
  OracleLmOptions o; o.max_iterations = max_iterations; o.trace_capacity = 0;
  o.opt_focal = 1; o.opt_pp = 0; o.opt_dist = 0;
  int rc = oracle_lm_solve(prob, &o, focal, pp, dist, poses, &summaries[0], nullptr, nullptr, nullptr);
  o.opt_dist = 1;
  rc |= oracle_lm_solve(prob, &o, focal, pp, dist, poses, &summaries[1], nullptr, nullptr, nullptr);
  o.opt_pp = 1;
  rc |= oracle_lm_solve(prob, &o, focal, pp, dist, poses, &summaries[2], nullptr, nullptr, nullptr);
  if (frame_rmse) oracle_frame_rmse(prob, focal, pp, dist, poses, frame_rmse);
  return rc;
}

int oracle_num_dist(int model) { return (model >= 0 && model < 7) ? kNumDist[model] : -1; }

}  // extern "C"
