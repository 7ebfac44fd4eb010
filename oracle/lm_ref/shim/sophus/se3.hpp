// Stand-in for <sophus/se3.hpp> -- TEST INFRASTRUCTURE ONLY (oracle/lm_ref).
// Sophus (unpinned master, /root/reference/build.sh:57-65) is neither in the reference tree nor in
// this image.  Restated from Sophus' published formulas, generic in the scalar so that ceres-style
// Jets differentiate through it: SO3::expAndTheta (Taylor branch below epsilon), SE3::exp (V matrix),
// the group product with the normalising SO3 constructor, SE3 * point (uv = q.vec x p; uv += uv;
// p + w uv + q.vec x uv -- Sophus spells this out itself so that Jets and doubles mix), inverse, log.
// Storage of Eigen::Map<SE3<T>>: [qx qy qz qw tx ty tz] (SE3::data()).
#pragma once
#include <cmath>
#include <Eigen/Core>

namespace Sophus {

template <class T> struct Constants { static double epsilon() { return 1e-10; } };

namespace detail {
inline double val(double x) { return x; }
template <class J> inline auto val(const J& x) -> decltype(x.a) { return x.a; }
using std::sqrt; using std::sin; using std::cos; using std::atan2;
}  // namespace detail

template <class T>
class SE3 {
 public:
  using Tangent = Eigen::Matrix<T, 6, 1>;
  using Point = Eigen::Matrix<T, 3, 1>;
  T q[4];  // x y z w
  T t[3];
  SE3() { q[0] = q[1] = q[2] = T(0.0); q[3] = T(1.0); t[0] = t[1] = t[2] = T(0.0); }
  explicit SE3(const T* p) { for (int i = 0; i < 4; ++i) q[i] = p[i]; for (int i = 0; i < 3; ++i) t[i] = p[4 + i]; }
  const T* data() const { return q; }

  template <class S>
  static void rotate(const T* qq, const S* p, T* out) {
    T uvx = qq[1] * p[2] - qq[2] * p[1];
    T uvy = qq[2] * p[0] - qq[0] * p[2];
    T uvz = qq[0] * p[1] - qq[1] * p[0];
    uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
    out[0] = p[0] + qq[3] * uvx + (qq[1] * uvz - qq[2] * uvy);
    out[1] = p[1] + qq[3] * uvy + (qq[2] * uvx - qq[0] * uvz);
    out[2] = p[2] + qq[3] * uvz + (qq[0] * uvy - qq[1] * uvx);
  }

  static void to_matrix(const T* qq, T R[3][3]) {  // Eigen::Quaternion::toRotationMatrix
    const T tx = 2.0 * qq[0], ty = 2.0 * qq[1], tz = 2.0 * qq[2];
    const T twx = tx * qq[3], twy = ty * qq[3], twz = tz * qq[3];
    const T txx = tx * qq[0], txy = ty * qq[0], txz = tz * qq[0];
    const T tyy = ty * qq[1], tyz = tz * qq[1], tzz = tz * qq[2];
    R[0][0] = 1.0 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
    R[1][0] = txy + twz; R[1][1] = 1.0 - (txx + tzz); R[1][2] = tyz - twx;
    R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1.0 - (txx + tyy);
  }

  // SE3::exp(a), a = (upsilon, omega)
  template <class Vec6>
  static SE3 exp(const Vec6& a) {
    using namespace detail;
    const T ups[3] = {a[0], a[1], a[2]};
    const T om[3] = {a[3], a[4], a[5]};
    const T theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    T theta, imag, real;
    if (val(theta_sq) < Constants<T>::epsilon() * Constants<T>::epsilon()) {
      theta = T(0.0);
      const T theta_po4 = theta_sq * theta_sq;
      imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
      real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
      theta = sqrt(theta_sq);
      const T half = 0.5 * theta;
      imag = sin(half) / theta;
      real = cos(half);
    }
    SE3 r;
    r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2]; r.q[3] = real;
    T Om[3][3] = {{T(0.0), -om[2], om[1]}, {om[2], T(0.0), -om[0]}, {-om[1], om[0], T(0.0)}};
    T V[3][3];
    if (std::fabs(val(theta)) < Constants<T>::epsilon()) {
      to_matrix(r.q, V);
    } else {
      T Om2[3][3];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T s(0.0); for (int k = 0; k < 3; ++k) s = s + Om[i][k] * Om[k][j]; Om2[i][j] = s; }
      const T th2 = theta * theta;
      const T c1 = (1.0 - cos(theta)) / th2;
      const T c2 = (theta - sin(theta)) / (th2 * theta);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j ? 1.0 : 0.0) + c1 * Om[i][j] + c2 * Om2[i][j];
    }
    for (int i = 0; i < 3; ++i) r.t[i] = V[i][0] * ups[0] + V[i][1] * ups[1] + V[i][2] * ups[2];
    return r;
  }

  // group product; the SO3 constructor from a quaternion normalises
  SE3 operator*(const SE3& o) const {
    using namespace detail;
    SE3 r;
    const T* a = q; const T* b = o.q;
    T w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    T x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    T y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    T z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    const T nrm = sqrt(x * x + y * y + z * z + w * w);
    r.q[0] = x / nrm; r.q[1] = y / nrm; r.q[2] = z / nrm; r.q[3] = w / nrm;
    T rt[3];
    rotate(q, o.t, rt);
    for (int i = 0; i < 3; ++i) r.t[i] = t[i] + rt[i];
    return r;
  }

  template <class S>
  Point operator*(const Eigen::Matrix<S, 3, 1>& p) const {
    T out[3];
    rotate(q, p.d, out);
    return Point(out[0] + t[0], out[1] + t[1], out[2] + t[2]);
  }

  SE3 inverse() const {
    SE3 r;
    r.q[0] = -q[0]; r.q[1] = -q[1]; r.q[2] = -q[2]; r.q[3] = q[3];
    T rt[3];
    rotate(r.q, t, rt);
    for (int i = 0; i < 3; ++i) r.t[i] = -rt[i];
    return r;
  }

  Tangent log() const {
    using namespace detail;
    const T sq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const T w = q[3];
    T two_atan_nbyw_by_n;
    if (val(sq) < Constants<T>::epsilon() * Constants<T>::epsilon()) {
      const T w2 = w * w;
      two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * sq / (w * w2);
    } else {
      const T n = sqrt(sq);
      const T at = (val(w) < 0.0) ? atan2(-n, -w) : atan2(n, w);
      two_atan_nbyw_by_n = 2.0 * at / n;
    }
    const T om[3] = {two_atan_nbyw_by_n * q[0], two_atan_nbyw_by_n * q[1], two_atan_nbyw_by_n * q[2]};
    const T theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    T Om[3][3] = {{T(0.0), -om[2], om[1]}, {om[2], T(0.0), -om[0]}, {-om[1], om[0], T(0.0)}};
    T Om2[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T s(0.0); for (int k = 0; k < 3; ++k) s = s + Om[i][k] * Om[k][j]; Om2[i][j] = s; }
    T c;
    if (val(theta_sq) < Constants<T>::epsilon() * Constants<T>::epsilon()) {
      c = T(1.0 / 12.0);
    } else {
      const T theta = sqrt(theta_sq);
      const T half = 0.5 * theta;
      c = (1.0 - theta * cos(half) / (2.0 * sin(half))) / theta_sq;
    }
    Tangent r;
    for (int i = 0; i < 3; ++i) {
      T s(0.0);
      for (int j = 0; j < 3; ++j) s = s + ((i == j ? 1.0 : 0.0) - 0.5 * Om[i][j] + c * Om2[i][j]) * t[j];
      r[i] = s;
      r[3 + i] = om[i];
    }
    return r;
  }
};

using SE3d = SE3<double>;

}  // namespace Sophus

namespace Eigen {

template <class T>
class Map<const Sophus::SE3<T>> {
 public:
  explicit Map(const T* p) : p_(p) {}
  operator Sophus::SE3<T>() const { return Sophus::SE3<T>(p_); }
  Sophus::SE3<T> inverse() const { return Sophus::SE3<T>(p_).inverse(); }
  Sophus::SE3<T> operator*(const Sophus::SE3<T>& o) const { return Sophus::SE3<T>(p_) * o; }
  template <class S>
  Matrix<T, 3, 1> operator*(const Matrix<S, 3, 1>& p) const { return Sophus::SE3<T>(p_) * p; }
 private:
  const T* p_;
};

template <class T>
class Map<Sophus::SE3<T>> {
 public:
  explicit Map(T* p) : p_(p) {}
  Map& operator=(const Sophus::SE3<T>& o) { for (int i = 0; i < 4; ++i) p_[i] = o.q[i]; for (int i = 0; i < 3; ++i) p_[4 + i] = o.t[i]; return *this; }
 private:
  T* p_;
};

}  // namespace Eigen
