// ezcalib_b200.hpp -- C++ host-side mirror of the reference's interfaces for the accelerated path, above the C ABI of
// ezcalib_b200.h.  Same class names, constructor arguments, method meaning and corner ordering as
//   CalibDetector / ChessboardCalibDetector / AprilTagCalibDetector   (/root/reference/include/target_detectors/*.hpp)
//   DistParam + Rad1/Rad2/Rad3/RadTan4/RadTan5/RadTan8/KB4DistParam    (/root/reference/include/dist_models/*.hpp)
//   IntrinsicsParam                                                   (/root/reference/include/calib_models.hpp:15-264)
//   Camera / CalibFrame                                               (/root/reference/include/camera.hpp:14-69, src/camera.cpp)
//   EZCalibrator::computeP3PRansac / computeCalibration / runMultiCameraCalib
//                                                                     (/root/reference/src/ezcalibrator.cpp:560-628, 105-339, 631-844)
// so a call site written against the reference reads the same.  OpenCV / Eigen / Sophus are not dependencies: the POD
// types below are layout-compatible with cv::Point2f, Eigen::Vector3d and Sophus::SE3d's storage
// (qx qy qz qw tx ty tz), so reference containers can be passed with a reinterpret_cast.
//
// The one structural difference is batching: detectTarget(image) of the reference is a pure function of the image, so
// the detectors here take the whole image set once (detectBatch) and detectTarget(i, ...) is a lookup -- the image loop
// of EZCalibrator::runCalibration (ezcalibrator.cpp:40-59, with its skip rule) is replayed by the caller unchanged.
// Errors of the C ABI become ezcalib_b200::Error (the reference calls exit(-1) on bad input; a library must not).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ezcalib_b200.h"

namespace ezcalib_b200 {

struct Point2f { float x, y; };            // == cv::Point2f
struct Vector3d { double x, y, z; };       // == Eigen::Vector3d
struct SE3d { double q[4]; double t[3]; }; // == Sophus::SE3d data(): unit quaternion (x y z w), translation

class Error : public std::runtime_error {
 public:
  Error(const std::string& what, ezc_status st) : std::runtime_error(what + ": " + ezc_last_error()), status(st) {}
  ezc_status status;
};
inline void check(ezc_status st, const char* what) {
  if (st != EZC_OK) throw Error(what, st);
}

class Context {
 public:
  explicit Context(int device = 0) { check(ezc_create(device, &h_), "ezc_create"); }
  ~Context() { ezc_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  ezc_context* handle() const { return h_; }
  uint64_t kernelLaunches() const { return ezc_kernel_launches(h_); }

 private:
  ezc_context* h_ = nullptr;
};

// A set of 8-bit gray images of one size resident in HBM (what cv::imread(..., IMREAD_GRAYSCALE) hands to
// EZCalibrator::processImage, one by one, in the reference).
class ImageBatch {
 public:
  // host pixels [n][height][width] (row_stride / image_stride in bytes, 0 = packed); async = copies overlap detection
  ImageBatch(Context& ctx, const uint8_t* pixels, int n, int width, int height, int64_t row_stride = 0, int64_t image_stride = 0,
             bool async = false)
      : n_(n), width_(width), height_(height) {
    ezc_image_batch b{pixels, n, width, height, row_stride, image_stride};
    check(async ? ezc_images_upload_async(ctx.handle(), &b, &h_) : ezc_images_upload(ctx.handle(), &b, &h_), "ezc_images_upload");
  }
  ~ImageBatch() { ezc_images_destroy(h_); }
  ImageBatch(const ImageBatch&) = delete;
  ImageBatch& operator=(const ImageBatch&) = delete;
  ezc_images* handle() const { return h_; }
  int size() const { return n_; }
  int width() const { return width_; }
  int height() const { return height_; }

 private:
  ezc_images* h_ = nullptr;
  int n_, width_, height_;
};

// base_detector.hpp:12-44
class CalibDetector {
 public:
  virtual ~CalibDetector() = default;
  // Runs the detector on every image of the batch (one call into the CUDA library).
  virtual void detectBatch(Context& ctx, const ImageBatch& imgs) = 0;
  // CalibDetector::detectTarget for image i of the last batch: same return value, same corner vector.
  virtual bool detectTarget(size_t i, std::vector<Point2f>& v_corners_pts) const = 0;
  virtual const std::vector<Vector3d>& getTargetCoords() const = 0;

 protected:
  std::vector<Vector3d> m_vtgt_coords;
  std::vector<float> m_corners;     // [n][n_pts][2]
  std::vector<uint8_t> m_success;   // [n]
  size_t m_n_pts = 0;
  bool lookup(size_t i, std::vector<Point2f>& out) const {
    if (i >= m_success.size()) throw std::out_of_range("detectTarget: image index outside the last batch");
    const Point2f* p = reinterpret_cast<const Point2f*>(&m_corners[i * m_n_pts * 2]);
    out.assign(p, p + m_n_pts);
    return m_success[i] != 0;
  }
};

// chessboard_detector.hpp:8-105.  nb_rows / nb_cols count SQUARES; the OpenCV pattern is (nb_cols-1, nb_rows-1).
class ChessboardCalibDetector : public CalibDetector {
 public:
  ChessboardCalibDetector(int nb_rows, int nb_cols, double square_size) : m_nb_rows(nb_rows), m_nb_cols(nb_cols), m_square_size(square_size) {
    if (!(nb_rows > 0 && nb_cols > 0 && square_size > 0.)) throw std::invalid_argument("ChessboardCalibDetector: bad target description");
    setupCalibTarget();
  }
  void detectBatch(Context& ctx, const ImageBatch& imgs) override {
    m_n_pts = m_vtgt_coords.size();
    m_corners.assign((size_t)imgs.size() * m_n_pts * 2, 0.f);
    m_success.assign(imgs.size(), 0);
    check(ezc_detect_chessboard_batch(ctx.handle(), imgs.handle(), m_nb_rows, m_nb_cols, m_corners.data(), m_success.data(), 0),
          "ezc_detect_chessboard_batch");
  }
  // On failure the reference leaves whatever partial corners findChessboardCorners produced; here the vector is zero-filled.
  bool detectTarget(size_t i, std::vector<Point2f>& v_corners_pts) const override { return lookup(i, v_corners_pts); }
  const std::vector<Vector3d>& getTargetCoords() const override { return m_vtgt_coords; }

 private:
  void setupCalibTarget() {  // chessboard_detector.hpp:71-84: OpenCV corner order, z = 1
    m_vtgt_coords.reserve((size_t)(m_nb_rows - 1) * (m_nb_cols - 1));
    for (int r = 1; r < m_nb_rows; ++r)
      for (int c = m_nb_cols - 1; c > 0; --c) m_vtgt_coords.push_back(Vector3d{-c * m_square_size, r * m_square_size, 1.});
  }
  int m_nb_rows, m_nb_cols;
  double m_square_size;
};

// aprilgrid_detector.hpp:8-150
class AprilTagCalibDetector : public CalibDetector {
 public:
  AprilTagCalibDetector(int nb_rows, int nb_cols, double tag_size, double tag_space)
      : m_nb_rows(nb_rows), m_nb_cols(nb_cols), m_nb_tags(nb_rows * nb_cols), m_tag_size(tag_size), m_tag_space(tag_space) {
    if (!(nb_rows > 0 && nb_cols > 0 && tag_size > 0. && tag_space > 0.)) throw std::invalid_argument("AprilTagCalibDetector: bad target description");
    setupCalibTarget();
  }
  void detectBatch(Context& ctx, const ImageBatch& imgs) override {
    m_n_pts = m_vtgt_coords.size();
    m_corners.assign((size_t)imgs.size() * m_n_pts * 2, 0.f);
    m_success.assign(imgs.size(), 0);
    m_n_detections.assign(imgs.size(), 0);
    check(ezc_detect_aprilgrid_batch(ctx.handle(), imgs.handle(), m_nb_rows, m_nb_cols, m_corners.data(), m_n_detections.data(), m_success.data(), 0),
          "ezc_detect_aprilgrid_batch");
  }
  // corners not detected stay (-1,-1), exactly like aprilgrid_detector.hpp:35
  bool detectTarget(size_t i, std::vector<Point2f>& v_corners_pts) const override { return lookup(i, v_corners_pts); }
  const std::vector<Vector3d>& getTargetCoords() const override { return m_vtgt_coords; }
  int nbDetections(size_t i) const { return m_n_detections.at(i); }

 private:
  void setupCalibTarget() {  // aprilgrid_detector.hpp:93-119
    m_vtgt_coords.reserve(4 * (size_t)m_nb_tags);
    const double tag_spacing = m_tag_size * m_tag_space;
    const double tags_offset = m_tag_size + tag_spacing;
    const Vector3d c1{0., 0., 1.}, c2{m_tag_size, 0., 1.}, c3{m_tag_size, -m_tag_size, 1.}, c4{0., -m_tag_size, 1.};
    for (int r = 0; r < m_nb_rows; ++r)
      for (int c = 0; c < m_nb_cols; ++c) {
        const double ox = c * tags_offset, oy = -r * tags_offset;
        for (const Vector3d& k : {c1, c2, c3, c4}) m_vtgt_coords.push_back(Vector3d{k.x + ox, k.y + oy, k.z});
      }
  }
  int m_nb_rows, m_nb_cols, m_nb_tags;
  double m_tag_size, m_tag_space;
  std::vector<int32_t> m_n_detections;
};

// Distortion models in the order of camera.cpp:67-94
enum class DistModel : int32_t { rad1 = EZC_RAD1, rad2 = EZC_RAD2, rad3 = EZC_RAD3, radtan4 = EZC_RADTAN4, radtan5 = EZC_RADTAN5,
                                 radtan8 = EZC_RADTAN8, kannalabrandt = EZC_KB4 };

// ---------------------------------------------------------------------------------------------------------------------
// Small SE3 helpers on the host (Sophus formulas: SO3::expAndTheta / SE3::exp / log / group product / inverse), needed by
// runMultiCameraCalib's initialisation (ezcalibrator.cpp:678-713).  Storage = Sophus::SE3d::data().
namespace se3 {
inline void rotate(const double* q, const double* p, double* out) {   // Sophus SO3 * point
  double uv[3] = {q[1] * p[2] - q[2] * p[1], q[2] * p[0] - q[0] * p[2], q[0] * p[1] - q[1] * p[0]};
  for (double& v : uv) v += v;
  out[0] = p[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  out[1] = p[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  out[2] = p[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}
inline Vector3d act(const SE3d& T, const Vector3d& p) {
  double o[3];
  rotate(T.q, &p.x, o);
  return Vector3d{o[0] + T.t[0], o[1] + T.t[1], o[2] + T.t[2]};
}
inline SE3d inverse(const SE3d& T) {
  SE3d r;
  r.q[0] = -T.q[0]; r.q[1] = -T.q[1]; r.q[2] = -T.q[2]; r.q[3] = T.q[3];
  double rt[3];
  rotate(r.q, T.t, rt);
  for (int i = 0; i < 3; ++i) r.t[i] = -rt[i];
  return r;
}
inline SE3d mul(const SE3d& A, const SE3d& B) {
  SE3d r;
  const double* a = A.q; const double* b = B.q;
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  const double n = std::sqrt(x * x + y * y + z * z + w * w);
  r.q[0] = x / n; r.q[1] = y / n; r.q[2] = z / n; r.q[3] = w / n;
  double rt[3];
  rotate(A.q, B.t, rt);
  for (int i = 0; i < 3; ++i) r.t[i] = A.t[i] + rt[i];
  return r;
}
inline void hat2(const double* om, double Om[3][3], double Om2[3][3]) {
  const double O[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Om[i][j] = O[i][j];
      Om2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
    }
}
inline std::array<double, 6> log(const SE3d& T) {   // (upsilon, omega)
  const double eps = 1e-10;
  const double* q = T.q;
  const double sq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], w = q[3];
  double two_atan;
  if (sq < eps * eps) {
    two_atan = 2.0 / w - (2.0 / 3.0) * sq / (w * w * w);
  } else {
    const double n = std::sqrt(sq);
    const double at = w < 0 ? std::atan2(-n, -w) : std::atan2(n, w);
    two_atan = 2.0 * at / n;
  }
  const double om[3] = {two_atan * q[0], two_atan * q[1], two_atan * q[2]};
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double Om[3][3], Om2[3][3];
  hat2(om, Om, Om2);
  double c;
  if (th2 < eps * eps) c = 1.0 / 12.0;
  else { const double th = std::sqrt(th2), h = 0.5 * th; c = (1.0 - th * std::cos(h) / (2.0 * std::sin(h))) / th2; }
  std::array<double, 6> r{};
  for (int i = 0; i < 3; ++i) {
    double sum = 0;
    for (int j = 0; j < 3; ++j) sum += ((i == j ? 1.0 : 0.0) - 0.5 * Om[i][j] + c * Om2[i][j]) * T.t[j];
    r[i] = sum;
    r[3 + i] = om[i];
  }
  return r;
}
inline SE3d exp(const std::array<double, 6>& a) {
  const double eps = 1e-10;
  const double* ups = &a[0]; const double* om = &a[3];
  const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double th, imag, real;
  if (th2 < eps * eps) { th = 0; const double t4 = th2 * th2; imag = 0.5 - th2 / 48.0 + t4 / 3840.0; real = 1.0 - th2 / 8.0 + t4 / 384.0; }
  else { th = std::sqrt(th2); imag = std::sin(0.5 * th) / th; real = std::cos(0.5 * th); }
  SE3d r;
  r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2]; r.q[3] = real;
  double Om[3][3], Om2[3][3];
  hat2(om, Om, Om2);
  double V[3][3];
  if (std::fabs(th) < eps) {
    const double* q = r.q;
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    const double M[3][3] = {{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = M[i][j];
  } else {
    const double c1 = (1.0 - std::cos(th)) / th2, c2 = (th - std::sin(th)) / (th2 * th);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j ? 1.0 : 0.0) + c1 * Om[i][j] + c2 * Om2[i][j];
  }
  for (int i = 0; i < 3; ++i) r.t[i] = V[i][0] * ups[0] + V[i][1] * ups[1] + V[i][2] * ups[2];
  return r;
}
}  // namespace se3

// ---------------------------------------------------------------------------------------------------------------------
// dist_model_base.hpp:8-108 and the seven concrete models (dist_model_k1.hpp ... dist_model_kb.hpp).
// createCeresCostFunction(u, v, wpt) is what the CUDA solver replaces (the residual + analytic Jacobian of each model live
// in csrc/lm_models.cuh); the mirror exposes model() so that the solver entry points can be addressed instead.
class DistParam {
 public:
  virtual ~DistParam() = default;
  virtual DistModel model() const = 0;
  virtual std::array<double, 2> distortCamPoint(double x, double y) const = 0;
  std::array<double, 2> distortCamPoint(const std::array<double, 2>& cam_pt) const { return distortCamPoint(cam_pt[0], cam_pt[1]); }
  std::array<double, 2> distortCamPoint(const Vector3d& cam_pt) const {
    const double inv_z = 1. / cam_pt.z;
    return distortCamPoint(cam_pt.x * inv_z, cam_pt.y * inv_z);
  }
  // Newton iteration with central differences, dist_model_base.hpp:33-71 (same step rule, same stopping rule)
  std::array<double, 2> undistortCamPoint(double xd, double yd) const {
    const double kMaxStepNorm = 1e-10, kRelStepSize = 1e-6, eps = std::numeric_limits<double>::epsilon();
    double x[2] = {xd, yd};
    for (int it = 0; it < 100; ++it) {
      const double s0 = std::max(eps, std::abs(kRelStepSize * x[0])), s1 = std::max(eps, std::abs(kRelStepSize * x[1]));
      auto d = [&](double a, double b) { const auto r = distortCamPoint(a, b); return std::array<double, 2>{r[0] - x[0], r[1] - x[1]}; };
      const auto dx = d(x[0], x[1]), dx0b = d(x[0] - s0, x[1] - s0), dx0f = d(x[0] + s0, x[1] + s0);
      const auto dx1b = d(x[0] - s1, x[1] - s1), dx1f = d(x[0] + s1, x[1] + s1);
      const double J00 = 1. + (dx0f[0] - dx0b[0]) / (2. * s0), J01 = (dx1f[0] - dx1b[0]) / (2. * s1);
      const double J10 = (dx0f[1] - dx0b[1]) / (2. * s0), J11 = 1. + (dx1f[1] - dx1b[1]) / (2. * s1);
      const double inv = 1.0 / (J00 * J11 - J01 * J10);
      const double r0 = x[0] + dx[0] - xd, r1 = x[1] + dx[1] - yd;
      const double st0 = (J11 * inv) * r0 + (-J01 * inv) * r1, st1 = (-J10 * inv) * r0 + (J00 * inv) * r1;
      x[0] -= st0; x[1] -= st1;
      if (st0 * st0 + st1 * st1 < kMaxStepNorm) break;
    }
    return {x[0], x[1]};
  }
  virtual void resetParameters(const std::vector<double>& v_dist_coefs) = 0;
  virtual std::vector<double> getDistParameters() const = 0;
  int getNumberOfParameters() const { return m_nb_params; }
  std::vector<double> getDistParamsStd() const { return m_vdist_params_std; }
  // the reference push_back()s onto a vector that already holds nb_params zeros (dist_model_base.hpp:90-102): kept, so the
  // YAML written from this mirror carries the same 2 * nd values
  void setDistParamsStd(const std::vector<double>& cov_dist) {
    for (int i = 0; i < m_nb_params; ++i) m_vdist_params_std.push_back(std::sqrt(cov_dist[(size_t)m_nb_params * i + i]));
  }

 protected:
  DistParam(int nb_params, bool use_mono_focal) : m_nb_params(nb_params), m_use_mono_focal(use_mono_focal), m_vdist_params_std(nb_params, 0.) {}
  int m_nb_params = -1;
  bool m_use_mono_focal;
  std::vector<double> m_vdist_params_std;
};

namespace detail {
template <int ND, DistModel M>
class DistParamN : public DistParam {
 public:
  explicit DistParamN(bool use_mono_focal) : DistParam(ND, use_mono_focal) { m_k.fill(0.); }
  DistModel model() const override { return M; }
  void resetParameters(const std::vector<double>& v) override {
    if ((int)v.size() != ND) throw std::invalid_argument("DistParam::resetParameters: wrong number of coefficients");
    std::copy(v.begin(), v.end(), m_k.begin());
  }
  std::vector<double> getDistParameters() const override { return std::vector<double>(m_k.begin(), m_k.end()); }
  using DistParam::distortCamPoint;
  std::array<double, 2> distortCamPoint(double x, double y) const override {
    const auto& k = m_k;
    if (M == DistModel::kannalabrandt) {   // dist_model_kb.hpp:43-56 (r == 0 gives 0/0 like the reference)
      const double r = std::sqrt(x * x + y * y), th = std::atan(r), th2 = th * th, th4 = th2 * th2, th6 = th4 * th2;
      const double D = th * (1. + th2 * (k[0] + th2 * k[1] + th4 * k[2] + th6 * k[3]));
      const double D_r = D / r;
      return {D_r * x, D_r * y};
    }
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r4 = r2 * r2;
    double D = 1.;
    switch (M) {
      case DistModel::rad1: D = 1. + k[0] * r2; break;
      case DistModel::rad2: case DistModel::radtan4: D = 1. + r2 * (k[0] + k[1] * r2); break;
      case DistModel::rad3: case DistModel::radtan5: D = 1. + r2 * (k[0] + k[1] * r2 + k[2] * r2 * r2); break;
      case DistModel::radtan8: D = (1. + r2 * (k[0] + k[1] * r2 + k[2] * r4)) / (1. + r2 * (k[3] + k[4] * r2 + k[5] * r4)); break;
      default: break;
    }
    if (M == DistModel::rad1 || M == DistModel::rad2 || M == DistModel::rad3) return {x * D, y * D};
    const double p1 = k[ND - 2], p2 = k[ND - 1], xy_2 = 2. * x * y;   // [.., p1, p2] in every radial-tangential model
    return {x * D + p1 * xy_2 + p2 * (r2 + 2. * x2), y * D + p2 * xy_2 + p1 * (r2 + 2. * y2)};
  }

 private:
  std::array<double, ND> m_k;
};
}  // namespace detail
using Rad1DistParam = detail::DistParamN<1, DistModel::rad1>;
using Rad2DistParam = detail::DistParamN<2, DistModel::rad2>;
using Rad3DistParam = detail::DistParamN<3, DistModel::rad3>;
using RadTan4DistParam = detail::DistParamN<4, DistModel::radtan4>;
using RadTan5DistParam = detail::DistParamN<5, DistModel::radtan5>;
using RadTan8DistParam = detail::DistParamN<8, DistModel::radtan8>;
using KB4DistParam = detail::DistParamN<4, DistModel::kannalabrandt>;

// Camera::setupInitialDistortion (camera.cpp:64-100): the config strings; unknown -> exception (the reference exit(-1)s)
inline std::unique_ptr<DistParam> makeDistParam(const std::string& dist_model, bool use_mono_focal) {
  if (dist_model == "rad1") return std::make_unique<Rad1DistParam>(use_mono_focal);
  if (dist_model == "rad2") return std::make_unique<Rad2DistParam>(use_mono_focal);
  if (dist_model == "rad3") return std::make_unique<Rad3DistParam>(use_mono_focal);
  if (dist_model == "radtan4") return std::make_unique<RadTan4DistParam>(use_mono_focal);
  if (dist_model == "radtan5") return std::make_unique<RadTan5DistParam>(use_mono_focal);
  if (dist_model == "radtan8") return std::make_unique<RadTan8DistParam>(use_mono_focal);
  if (dist_model == "kannalabrandt") return std::make_unique<KB4DistParam>(use_mono_focal);
  throw std::invalid_argument("The provided distortion model: " + dist_model + " is not implemented!");
}

// calib_models.hpp:15-264
class IntrinsicsParam {
 public:
  IntrinsicsParam(double img_width, double img_height, double fx, double fy, double cx, double cy)
      : m_img_width(img_width), m_img_height(img_height), m_fx(fx), m_fy(fy), m_cx(cx), m_cy(cy), m_use_mono_focal(false) {}
  IntrinsicsParam(double img_width, double img_height, double f, double cx, double cy, bool use_mono_focal = true)
      : m_img_width(img_width), m_img_height(img_height), m_fx(f), m_fy(f), m_cx(cx), m_cy(cy), m_use_mono_focal(use_mono_focal) {}
  // K^-1 (u, v, 1) with Eigen's cofactor inverse of K, exactly what ezc_p3p_ransac_batch applies (calib_models.hpp:69-77)
  Vector3d projImageToCam(double u, double v) const {
    const double invdet = 1.0 / (m_fx * m_fy);
    return Vector3d{(m_fy * invdet * u + 0.0 * v) + (0.0 * m_cy - m_cx * m_fy) * invdet, (0.0 * u + m_fx * invdet * v) + (m_cx * 0.0 - m_fx * m_cy) * invdet, 1.0};
  }
  std::array<double, 2> projCamToImage(const std::array<double, 2>& n) const { return {m_fx * n[0] + m_cx, m_fy * n[1] + m_cy}; }
  std::array<double, 2> projCamToImage(const Vector3d& p) const { const double iz = 1. / p.z; return projCamToImage(std::array<double, 2>{p.x * iz, p.y * iz}); }
  // The reference caches W - 0.5 / H - 0.5 in function-local statics initialised from the FIRST instance ever used
  // (calib_models.hpp:101-102, 111-112).  Per-instance borders here: identical whenever all cameras share
  // one resolution (every BASELINE config), and what the solver entry points take (ezc_lm_problem::img_w / img_h).
  bool isPointInImage(double u, double v) const { return u > -0.5 && v > -0.5 && u < m_img_width - 0.5 && v < m_img_height - 0.5; }
  bool isPointInImage(float u, float v) const { return u > -0.5f && v > -0.5f && u < (float)m_img_width - 0.5f && v < (float)m_img_height - 0.5f; }
  void resetParameters(const std::vector<double>& vfocal, const std::vector<double>& vpp) {
    if (m_use_mono_focal) { m_fx = vfocal.at(0); m_fy = vfocal.at(0); } else { m_fx = vfocal.at(0); m_fy = vfocal.at(1); }
    m_cx = vpp.at(0); m_cy = vpp.at(1);
  }
  void resetParameters(double fx, double fy, double cx, double cy) { m_fx = fx; m_fy = fy; m_cx = cx; m_cy = cy; }
  void resetParameters(double f, double cx, double cy) { m_fx = f; m_fy = f; m_cx = cx; m_cy = cy; }
  std::vector<double> getFocal() const { return m_use_mono_focal ? std::vector<double>{m_fx} : std::vector<double>{m_fx, m_fy}; }
  std::vector<double> getPrincipalPoint() const { return {m_cx, m_cy}; }
  double getMeanFocal() const { return m_use_mono_focal ? m_fx : (m_fx + m_fy) * 0.5; }
  void setFocalStd(const std::vector<double>& cov) {   // calib_models.hpp:201-211: covariance in, standard deviations stored
    if (m_use_mono_focal) { m_fx_std = std::sqrt(cov.at(0)); m_fy_std = m_fx_std; } else { m_fx_std = std::sqrt(cov.at(0)); m_fy_std = std::sqrt(cov.at(3)); }
  }
  void setPPStd(const std::vector<double>& cov) { m_cx_std = std::sqrt(cov.at(0)); m_cy_std = std::sqrt(cov.at(3)); }
  double getHFOV() const { return 2. * std::atan2(m_img_width / 2., m_fx) * (180. / 3.14159265358979323846); }
  double getVFOV() const { return 2. * std::atan2(m_img_height / 2., m_fy) * (180. / 3.14159265358979323846); }

  double m_img_width, m_img_height;
  double m_fx, m_fy, m_cx, m_cy;
  double m_fx_std = 0., m_fy_std = 0., m_cx_std = 0., m_cy_std = 0.;
  bool m_use_mono_focal;
};

// camera.hpp:14-33
struct CalibFrame {
  CalibFrame() = default;
  CalibFrame(const std::string& img_name, double timestamp, const std::vector<Point2f>& v_corners_px, const SE3d& T_world_2_cam)
      : m_img_name(img_name), m_timestamp(timestamp), m_v_corners_px(v_corners_px), m_T_world_2_cam(T_world_2_cam) {}
  std::string m_img_name = "";
  double m_timestamp = -1.;
  std::vector<Point2f> m_v_corners_px;   // one entry per target point, (-1,-1) = not observed
  SE3d m_T_world_2_cam{{0, 0, 0, 1}, {0, 0, 0}};
  double rmse_err = -1.;
};

// camera.hpp:35-69 / camera.cpp.  No file-system checks here (directory scanning and imread stay with the application).
class Camera {
 public:
  Camera(const std::string& input_images_path, bool use_mono_focal, const std::string& dist_model, double prior_fov)
      : m_input_images_path(input_images_path), m_use_mono_focal(use_mono_focal), m_dist_model(dist_model), m_prior_fov_deg(prior_fov) {
    if (m_prior_fov_deg < 1.) throw std::invalid_argument("Provided prior field of view should be positive!");
    setupInitialDistortion();
  }
  void setupInitialDistortion() { m_pdist_params = makeDistParam(m_dist_model, m_use_mono_focal); }
  void setupInitialCalibration(int img_width_px, int img_height_px) {   // camera.cpp:102-120
    const double img_width = img_width_px, img_height = img_height_px;
    const double prior_cx = img_width / 2., prior_cy = img_height / 2.;
    const double PI = 2. * std::asin(1.), DEG2RAD = PI / 180.;
    const double inv_tan_half_fov = 1. / std::tan(m_prior_fov_deg / 2. * DEG2RAD);
    const double prior_focal = std::max(prior_cx * inv_tan_half_fov, prior_cy * inv_tan_half_fov);
    m_pcalib_params = std::make_unique<IntrinsicsParam>(img_width, img_height, prior_focal, prior_cx, prior_cy, m_use_mono_focal);
  }
  std::array<double, 2> undistortPx(const Point2f& corner_pt) const {   // camera.cpp:123-132
    const Vector3d c = m_pcalib_params->projImageToCam(corner_pt.x, corner_pt.y);
    return m_pcalib_params->projCamToImage(m_pdist_params->undistortCamPoint(c.x / c.z, c.y / c.z));
  }
  int m_id = -1;
  std::string m_input_images_path;
  std::unique_ptr<IntrinsicsParam> m_pcalib_params;
  std::unique_ptr<DistParam> m_pdist_params;
  std::vector<CalibFrame> m_v_calib_frames;
  SE3d m_T_cam0_2_cam{{0, 0, 0, 1}, {0, 0, 0}};
  bool m_use_mono_focal = true;
  std::string m_dist_model = "";
  double m_prior_fov_deg = -1.;
};

struct CalibResult {
  std::vector<double> focal, pp, dist;            // getFocal() / getPrincipalPoint() / getDistParameters()
  std::vector<double> focal_std, pp_std, dist_std; // setFocalStd / setPPStd / setDistParamsStd (ezcalibrator.cpp:293-331); empty if the
                                                   // covariance estimation failed (the reference prints and carries on, :327-331)
  bool covariance_ok = false;
  ezc_lm_summary stage[3];                         // the three ceres::Solve calls (focal | + dist | + pp)
};

// EZCalibrator::computeP3PRansac (ezcalibrator.cpp:560-628) for a batch of views: corners[v] holds one cv::Point2f per target
// point.  Returns the success flag per view; poses[v] is T_world_2_cam (after Rodrigues -> float -> fitToSO3), n_inliers[v] the
// size of the reference's _v_inliers.
inline std::vector<uint8_t> computeP3PRansac(Context& ctx, const std::vector<std::vector<Point2f>>& corners, const std::vector<Vector3d>& target,
                                             const IntrinsicsParam& calib_params, std::vector<SE3d>& poses, std::vector<int32_t>* n_inliers = nullptr) {
  const size_t n = corners.size(), m = target.size();
  std::vector<float> px;
  px.reserve(n * m * 2);
  for (const auto& v : corners) {
    if (v.size() != m) throw std::invalid_argument("computeP3PRansac: a view does not hold one corner per target point");
    for (const Point2f& c : v) { px.push_back(c.x); px.push_back(c.y); }
  }
  std::vector<uint8_t> ok(n, 0);
  std::vector<double> p7(n * 7, 0.);
  std::vector<int32_t> ninl(n, 0);
  const std::vector<double> focal = calib_params.getFocal(), pp = calib_params.getPrincipalPoint();
  if (n)
    check(ezc_p3p_ransac_batch(ctx.handle(), px.data(), &target[0].x, (int32_t)n, (int32_t)m, (int32_t)m, focal.data(), (int32_t)focal.size(), pp.data(),
                               (int32_t)calib_params.m_img_width, (int32_t)calib_params.m_img_height, 1000, 4.0 / calib_params.getMeanFocal(), 0.999, 32,
                               ok.data(), p7.data(), ninl.data(), nullptr, nullptr, nullptr),
          "ezc_p3p_ransac_batch");
  poses.resize(n);
  for (size_t i = 0; i < n; ++i) {
    for (int j = 0; j < 4; ++j) poses[i].q[j] = p7[7 * i + j];
    for (int j = 0; j < 3; ++j) poses[i].t[j] = p7[7 * i + 4 + j];
  }
  if (n_inliers) *n_inliers = ninl;
  return ok;
}

// EZCalibrator::computeCalibration (ezcalibrator.cpp:105-339): three-stage solve, per-frame RMSE, covariance.
// focal (1 value if use_mono_focal else 2), pp (2) and dist (model dependent) are the initial values from
// Camera::setupInitialCalibration / setupInitialDistortion; frames carry the P3P poses and get poses + rmse_err back.
inline CalibResult computeCalibration(Context& ctx, DistModel model, bool use_mono_focal, std::vector<CalibFrame>& frames,
                                      const std::vector<Vector3d>& target, double img_width, double img_height,
                                      std::vector<double> focal, std::vector<double> pp, std::vector<double> dist) {
  const int nd = ezc_num_dist((int32_t)model);
  const int nf = use_mono_focal ? 1 : 2;
  if ((int)focal.size() < nf || pp.size() < 2 || (int)dist.size() < nd) throw std::invalid_argument("computeCalibration: parameter vectors too short");
  const size_t n = frames.size(), m = target.size();
  std::vector<float> obs;
  obs.reserve(n * m * 2);
  std::vector<double> poses;
  poses.reserve(n * 7);
  for (const CalibFrame& f : frames) {
    if (f.m_v_corners_px.size() != m) throw std::invalid_argument("computeCalibration: a frame does not hold one corner per target point");
    for (const Point2f& c : f.m_v_corners_px) { obs.push_back(c.x); obs.push_back(c.y); }
    poses.insert(poses.end(), f.m_T_world_2_cam.q, f.m_T_world_2_cam.q + 4);
    poses.insert(poses.end(), f.m_T_world_2_cam.t, f.m_T_world_2_cam.t + 3);
  }
  ezc_lm_problem p{};
  p.model = (int32_t)model;
  p.use_mono_focal = use_mono_focal ? 1 : 0;
  p.n_blocks = p.n_blocks_global = (int32_t)n;
  p.obs_per_block = p.pts_period = (int32_t)m;
  p.img_w = img_width; p.img_h = img_height;
  p.pts = &target[0].x;
  p.obs = obs.data();
  ezc_lm_solver* s = nullptr;
  check(ezc_lm_create(ctx.handle(), &p, &s), "ezc_lm_create");
  CalibResult r;
  struct Guard { ezc_lm_solver* s; ~Guard() { ezc_lm_destroy(s); } } guard{s};
  check(ezc_lm_set_params(s, focal.data(), pp.data(), dist.data(), poses.data()), "ezc_lm_set_params");
  ezc_lm_options o;
  ezc_lm_default_options(&o);
  for (int stage = 0; stage < 3; ++stage) {   // ezcalibrator.cpp:186-229: focal + poses, then + distortion, then + principal point
    o.opt_focal = 1; o.opt_dist = stage >= 1; o.opt_pp = stage >= 2;
    check(ezc_lm_solve(s, &o, &r.stage[stage]), "ezc_lm_solve");
  }
  std::vector<double> rmse(n);
  check(ezc_lm_frame_rmse(s, rmse.data()), "ezc_lm_frame_rmse");
  focal.resize(2); pp.resize(2); dist.resize(8);
  check(ezc_lm_get_params(s, focal.data(), pp.data(), dist.data(), poses.data()), "ezc_lm_get_params");
  const int k = nf + 2 + nd;
  std::vector<double> cov((size_t)k * k, 0.);
  const ezc_status cst = ezc_lm_covariance(s, cov.data());
  if (cst != EZC_OK && cst != EZC_ERR_NUMERIC) check(cst, "ezc_lm_covariance");
  r.covariance_ok = cst == EZC_OK;   // a rank-deficient system is not fatal: "Covariance Estimation Failed", calibration kept
  r.focal.assign(focal.begin(), focal.begin() + nf);
  r.pp.assign(pp.begin(), pp.begin() + 2);
  r.dist.assign(dist.begin(), dist.begin() + nd);
  if (r.covariance_ok) {
    auto sd = [&](int i) { return std::sqrt(cov[(size_t)i * k + i]); };
    for (int i = 0; i < nf; ++i) r.focal_std.push_back(sd(i));
    for (int i = 0; i < 2; ++i) r.pp_std.push_back(sd(nf + i));
    for (int i = 0; i < nd; ++i) r.dist_std.push_back(sd(nf + 2 + i));
  }
  for (size_t i = 0; i < n; ++i) {
    for (int j = 0; j < 4; ++j) frames[i].m_T_world_2_cam.q[j] = poses[7 * i + j];
    for (int j = 0; j < 3; ++j) frames[i].m_T_world_2_cam.t[j] = poses[7 * i + 4 + j];
    frames[i].rmse_err = rmse[i];
  }
  return r;
}

// The same on a Camera, writing the result back like ezcalibrator.cpp:199-200, 213-214, 227-228, 314-325.
inline CalibResult computeCalibration(Context& ctx, Camera& cam, const std::vector<Vector3d>& target) {
  if (!cam.m_pcalib_params || !cam.m_pdist_params) throw std::invalid_argument("computeCalibration: camera not initialised");
  IntrinsicsParam& K = *cam.m_pcalib_params;
  CalibResult r = computeCalibration(ctx, cam.m_pdist_params->model(), cam.m_use_mono_focal, cam.m_v_calib_frames, target, K.m_img_width, K.m_img_height,
                                     K.getFocal(), K.getPrincipalPoint(), cam.m_pdist_params->getDistParameters());
  K.resetParameters(r.focal, r.pp);
  cam.m_pdist_params->resetParameters(r.dist);
  if (r.covariance_ok) {
    K.m_fx_std = r.focal_std.front(); K.m_fy_std = r.focal_std.back();
    K.m_cx_std = r.pp_std[0]; K.m_cy_std = r.pp_std[1];
    const int nd = cam.m_pdist_params->getNumberOfParameters();
    std::vector<double> cov((size_t)nd * nd, 0.);
    for (int i = 0; i < nd; ++i) cov[(size_t)nd * i + i] = r.dist_std[i] * r.dist_std[i];
    cam.m_pdist_params->setDistParamsStd(cov);
  }
  return r;
}

// EZCalibrator::refineCalibration (ezcalibrator.cpp:342-557; dormant in the reference -- its call at :63 is commented out):
// one solve with every block variable over the frames whose rmse_err <= 1; their poses / rmse_err are updated, the other
// frames are left as they are; parameters and standard deviations written back like :428-429, :523-537.
inline ezc_lm_summary refineCalibration(Context& ctx, Camera& cam, const std::vector<Vector3d>& target) {
  if (!cam.m_pcalib_params || !cam.m_pdist_params) throw std::invalid_argument("refineCalibration: camera not initialised");
  IntrinsicsParam& K = *cam.m_pcalib_params;
  const DistModel model = cam.m_pdist_params->model();
  const int nf = cam.m_use_mono_focal ? 1 : 2, nd = cam.m_pdist_params->getNumberOfParameters();
  std::vector<size_t> good;
  for (size_t i = 0; i < cam.m_v_calib_frames.size(); ++i)
    if (!(cam.m_v_calib_frames[i].rmse_err > 1.)) good.push_back(i);
  if (good.empty()) throw std::invalid_argument("refineCalibration: no frame with rmse_err <= 1");
  const size_t n = good.size(), m = target.size();
  std::vector<float> obs;
  std::vector<double> poses;
  for (size_t i : good) {
    const CalibFrame& f = cam.m_v_calib_frames[i];
    for (const Point2f& c : f.m_v_corners_px) { obs.push_back(c.x); obs.push_back(c.y); }
    poses.insert(poses.end(), f.m_T_world_2_cam.q, f.m_T_world_2_cam.q + 4);
    poses.insert(poses.end(), f.m_T_world_2_cam.t, f.m_T_world_2_cam.t + 3);
  }
  ezc_lm_problem p{};
  p.model = (int32_t)model; p.use_mono_focal = cam.m_use_mono_focal ? 1 : 0;
  p.n_blocks = p.n_blocks_global = (int32_t)n;
  p.obs_per_block = p.pts_period = (int32_t)m;
  p.img_w = K.m_img_width; p.img_h = K.m_img_height;
  p.pts = &target[0].x; p.obs = obs.data();
  ezc_lm_solver* s = nullptr;
  check(ezc_lm_create(ctx.handle(), &p, &s), "ezc_lm_create");
  struct Guard { ezc_lm_solver* s; ~Guard() { ezc_lm_destroy(s); } } guard{s};
  std::vector<double> focal = K.getFocal(), pp = K.getPrincipalPoint(), dist = cam.m_pdist_params->getDistParameters();
  focal.resize(2); dist.resize(8);
  check(ezc_lm_set_params(s, focal.data(), pp.data(), dist.data(), poses.data()), "ezc_lm_set_params");
  ezc_lm_options o;
  ezc_lm_default_options(&o);
  ezc_lm_summary summary{};
  check(ezc_lm_solve(s, &o, &summary), "ezc_lm_solve");
  std::vector<double> rmse(n);
  check(ezc_lm_frame_rmse(s, rmse.data()), "ezc_lm_frame_rmse");
  check(ezc_lm_get_params(s, focal.data(), pp.data(), dist.data(), poses.data()), "ezc_lm_get_params");
  K.resetParameters(std::vector<double>(focal.begin(), focal.begin() + nf), pp);
  cam.m_pdist_params->resetParameters(std::vector<double>(dist.begin(), dist.begin() + nd));
  for (size_t a = 0; a < n; ++a) {
    CalibFrame& f = cam.m_v_calib_frames[good[a]];
    for (int j = 0; j < 4; ++j) f.m_T_world_2_cam.q[j] = poses[7 * a + j];
    for (int j = 0; j < 3; ++j) f.m_T_world_2_cam.t[j] = poses[7 * a + 4 + j];
    f.rmse_err = rmse[a];
  }
  const int k = nf + 2 + nd;
  std::vector<double> cov((size_t)k * k, 0.);
  const ezc_status cst = ezc_lm_covariance(s, cov.data());
  if (cst != EZC_OK && cst != EZC_ERR_NUMERIC) check(cst, "ezc_lm_covariance");
  if (cst == EZC_OK) {
    K.m_fx_std = std::sqrt(cov[0]); K.m_fy_std = std::sqrt(cov[(size_t)(nf - 1) * k + (nf - 1)]);
    K.m_cx_std = std::sqrt(cov[(size_t)nf * k + nf]); K.m_cy_std = std::sqrt(cov[(size_t)(nf + 1) * k + nf + 1]);
    std::vector<double> cd((size_t)nd * nd, 0.);
    for (int i = 0; i < nd; ++i) cd[(size_t)nd * i + i] = cov[(size_t)(nf + 2 + i) * k + nf + 2 + i];
    cam.m_pdist_params->setDistParamsStd(cd);
  }
  return summary;
}

struct MultiCamResult {
  ezc_lm_summary summary{};
  size_t nb_good_pairs = 0;
  std::vector<SE3d> init;   // v_opt_Tcic0 before the solve
};

// EZCalibrator::runMultiCameraCalib (ezcalibrator.cpp:631-844): intrinsics of cameras 1..N-1 frozen, one SE3 block T_ci_c0
// per camera initialised from the per-coefficient median of log(T_ci_w * T_c0_w^-1) -- with the reference's quirk Q3 (cam0's
// frame is indexed with the CAMERA index, :680) -- one joint solve, results written to m_T_cam0_2_cam.
// All cameras 1..N-1 must share dist_model / use_mono_focal (true for every BASELINE config).
inline MultiCamResult runMultiCameraCalib(Context& ctx, std::vector<Camera>& v_cameras, const std::vector<Vector3d>& v_tgt_coords) {
  MultiCamResult res;
  const size_t nb_cams = v_cameras.size();
  if (nb_cams < 2) return res;
  const Camera& cam0 = v_cameras[0];
  const size_t n0 = cam0.m_v_calib_frames.size(), m = v_tgt_coords.size();
  const DistModel model = v_cameras[1].m_pdist_params->model();
  const bool mono = v_cameras[1].m_use_mono_focal;
  const int nf = mono ? 1 : 2, nd = ezc_num_dist((int32_t)model);
  std::vector<double> focal, pp, dist, ext;
  for (size_t i = 1; i < nb_cams; ++i) {
    const Camera& camj = v_cameras[i];
    if (camj.m_pdist_params->model() != model || camj.m_use_mono_focal != mono)
      throw std::invalid_argument("runMultiCameraCalib: mixed distortion models across cameras are not supported");
    const auto f = camj.m_pcalib_params->getFocal(), c = camj.m_pcalib_params->getPrincipalPoint(), d = camj.m_pdist_params->getDistParameters();
    focal.insert(focal.end(), f.begin(), f.end()); pp.insert(pp.end(), c.begin(), c.end()); dist.insert(dist.end(), d.begin(), d.end());
    std::array<std::vector<double>, 6> tab;
    for (size_t j = 0; j < n0; ++j) {
      if (i >= n0) throw std::out_of_range("runMultiCameraCalib: the reference indexes cam0.m_v_calib_frames[camera index] out of range here (quirk Q3)");
      const CalibFrame& f0 = cam0.m_v_calib_frames[i];
      if (f0.rmse_err > 1.) continue;
      for (const CalibFrame& fj : camj.m_v_calib_frames) {
        if (f0.m_img_name == fj.m_img_name) {
          ++res.nb_good_pairs;
          const auto lg = se3::log(se3::mul(fj.m_T_world_2_cam, se3::inverse(f0.m_T_world_2_cam)));
          for (int l = 0; l < 6; ++l) tab[l].push_back(lg[l]);
          break;
        } else if (fj.m_img_name > f0.m_img_name) {
          break;
        }
      }
    }
    if (tab[0].empty()) throw std::out_of_range("runMultiCameraCalib: no usable frame pair for the extrinsic initialisation (the reference reads out of bounds, :707-713)");
    const size_t med_idx = tab[0].size() / 2;
    std::array<double, 6> med{};
    for (int j = 0; j < 6; ++j) { std::sort(tab[j].begin(), tab[j].end()); med[j] = tab[j][med_idx]; }
    const SE3d T0 = se3::exp(med);
    res.init.push_back(T0);
    ext.insert(ext.end(), T0.q, T0.q + 4); ext.insert(ext.end(), T0.t, T0.t + 3);
  }
  // residual blocks (:727-785): points = T_c0_w * p_w of every cam0 frame; observations of camera j where names match
  std::vector<double> pts(n0 * m * 3);
  std::vector<float> obs((nb_cams - 1) * n0 * m * 2, -1.f);
  for (size_t a = 0; a < n0; ++a) {
    const CalibFrame& f0 = cam0.m_v_calib_frames[a];
    for (size_t k = 0; k < m; ++k) {
      const Vector3d c = se3::act(f0.m_T_world_2_cam, v_tgt_coords[k]);
      pts[(a * m + k) * 3] = c.x; pts[(a * m + k) * 3 + 1] = c.y; pts[(a * m + k) * 3 + 2] = c.z;
    }
    if (f0.rmse_err > 1.) continue;
    for (size_t j = 1; j < nb_cams; ++j)
      for (const CalibFrame& fj : v_cameras[j].m_v_calib_frames) {
        if (fj.rmse_err > 1.) continue;
        if (f0.m_img_name == fj.m_img_name) {
          float* o = &obs[((j - 1) * n0 + a) * m * 2];
          for (size_t k = 0; k < m; ++k) { o[2 * k] = fj.m_v_corners_px[k].x; o[2 * k + 1] = fj.m_v_corners_px[k].y; }
          break;
        } else if (fj.m_img_name > f0.m_img_name) {
          break;
        }
      }
  }
  ezc_lm_problem p{};
  p.model = (int32_t)model; p.use_mono_focal = mono ? 1 : 0;
  p.n_blocks = p.n_blocks_global = (int32_t)(nb_cams - 1);
  p.obs_per_block = p.pts_period = (int32_t)(n0 * m);
  p.cams_per_block = 1;
  p.img_w = v_cameras[1].m_pcalib_params->m_img_width; p.img_h = v_cameras[1].m_pcalib_params->m_img_height;
  p.pts = pts.data(); p.obs = obs.data();
  (void)nf; (void)nd;
  check(ezc_lm_multicam(ctx.handle(), &p, focal.data(), pp.data(), dist.data(), ext.data(), &res.summary), "ezc_lm_multicam");
  for (size_t i = 1; i < nb_cams; ++i) {
    for (int j = 0; j < 4; ++j) v_cameras[i].m_T_cam0_2_cam.q[j] = ext[7 * (i - 1) + j];
    for (int j = 0; j < 3; ++j) v_cameras[i].m_T_cam0_2_cam.t[j] = ext[7 * (i - 1) + 4 + j];
  }
  return res;
}

}  // namespace ezcalib_b200
