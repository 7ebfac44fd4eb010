// ezcalib_b200.hpp -- C++ host-side mirror of the reference's interfaces for the accelerated path, above the C ABI of
// ezcalib_b200.h.  Same class names, constructor arguments, method meaning and corner ordering as
//   CalibDetector / ChessboardCalibDetector / AprilTagCalibDetector   (/root/reference/include/target_detectors/*.hpp)
//   EZCalibrator::computeCalibration                                  (/root/reference/src/ezcalibrator.cpp:105-339)
// so a call site written against the reference reads the same.  OpenCV / Eigen / Sophus are not dependencies: the POD
// types below are layout-compatible with cv::Point2f, Eigen::Vector3d and Sophus::SE3d's storage
// (qx qy qz qw tx ty tz), so reference containers can be passed with a reinterpret_cast.
//
// The one structural difference is batching: detectTarget(image) of the reference is a pure function of the image, so
// the detectors here take the whole image set once (detectBatch) and detectTarget(i, ...) is a lookup -- the image loop
// of EZCalibrator::runCalibration (ezcalibrator.cpp:40-59, with its skip rule) is replayed by the caller unchanged.
// Errors of the C ABI become ezcalib_b200::Error (the reference calls exit(-1) on bad input; a library must not).
#pragma once

#include <cmath>
#include <cstdint>
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

// camera.hpp CalibFrame: corners of one view, its pose and (after the solve) its RMS reprojection error
struct CalibFrame {
  std::vector<Point2f> m_v_corners_px;   // one entry per target point, (-1,-1) = not observed
  SE3d m_T_world_2_cam;
  double rmse_err = 0.;
};

// Distortion models in the order of camera.cpp:67-94
enum class DistModel : int32_t { rad1 = EZC_RAD1, rad2 = EZC_RAD2, rad3 = EZC_RAD3, radtan4 = EZC_RADTAN4, radtan5 = EZC_RADTAN5,
                                 radtan8 = EZC_RADTAN8, kannalabrandt = EZC_KB4 };

struct CalibResult {
  std::vector<double> focal, pp, dist;            // getFocal() / getPrincipalPoint() / getDistParameters()
  std::vector<double> focal_std, pp_std, dist_std; // setFocalStd / setPPStd / setDistParamsStd (ezcalibrator.cpp:293-331)
  ezc_lm_summary stage[3];                         // the three ceres::Solve calls (focal | + dist | + pp)
};

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
  check(ezc_lm_covariance(s, cov.data()), "ezc_lm_covariance");
  r.focal.assign(focal.begin(), focal.begin() + nf);
  r.pp.assign(pp.begin(), pp.begin() + 2);
  r.dist.assign(dist.begin(), dist.begin() + nd);
  auto sd = [&](int i) { return std::sqrt(cov[(size_t)i * k + i]); };
  for (int i = 0; i < nf; ++i) r.focal_std.push_back(sd(i));
  for (int i = 0; i < 2; ++i) r.pp_std.push_back(sd(nf + i));
  for (int i = 0; i < nd; ++i) r.dist_std.push_back(sd(nf + 2 + i));
  for (size_t i = 0; i < n; ++i) {
    for (int j = 0; j < 4; ++j) frames[i].m_T_world_2_cam.q[j] = poses[7 * i + j];
    for (int j = 0; j < 3; ++j) frames[i].m_T_world_2_cam.t[j] = poses[7 * i + 4 + j];
    frames[i].rmse_err = rmse[i];
  }
  return r;
}

}  // namespace ezcalib_b200
