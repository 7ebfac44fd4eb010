// Exercises include/ezcalib_b200.hpp (the C++ mirror of the reference's detector / calibrator interfaces) end to end:
//   argv[1] = binary file written by tests/test_cpp_mirror_gpu.py:
//     int32 n_cb, w, h ; uint8 cb_images[n_cb][h][w]         (chessboard views, 10 x 7 squares of 0.05 m)
//     int32 n_ag, w2, h2 ; uint8 ag_images[n_ag][h2][w2]     (AprilGrid views, 6 x 6 tags)
//     int32 model, mono, n_views, n_pts ; double img_w, img_h ; double focal[2], pp[2], dist[8]
//     double target[n_pts][3] ; float obs[n_views][n_pts][2] ; double poses[n_views][7]
// Prints one line per result in a fixed text format that the Python test compares with the ctypes API's output.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "ezcalib_b200.hpp"

using namespace ezcalib_b200;

template <class T> static void rd(std::ifstream& f, T* p, size_t n) {
  f.read(reinterpret_cast<char*>(p), sizeof(T) * n);
  if (!f) { std::fprintf(stderr, "short read\n"); std::exit(2); }
}

// CPU-only: the DistParam / IntrinsicsParam / Camera mirrors (no CUDA context).  One line per quantity; the Python test
// compares them with the reference's own headers compiled into oracle/_ref/liblm_ref.so.
static int cpu_mirror() {
  const char* names[7] = {"rad1", "rad2", "rad3", "radtan4", "radtan5", "radtan8", "kannalabrandt"};
  const double coefs[8] = {-0.21, 0.06, -0.012, 0.004, 0.0011, -0.0007, 0.0009, -0.0005};
  for (int m = 0; m < 7; ++m) {
    auto d = makeDistParam(names[m], m % 2 == 0);
    const int nd = d->getNumberOfParameters();
    if ((int)d->model() != m || nd != ezc_num_dist(m)) { std::printf("model-table FAILED\n"); return 1; }
    std::vector<double> v(coefs, coefs + nd);
    d->resetParameters(v);
    const auto back = d->getDistParameters();
    const auto o = d->distortCamPoint(0.31, -0.17);
    const auto o3 = d->distortCamPoint(Vector3d{0.62, -0.34, 2.0});
    const auto u = d->undistortCamPoint(o[0], o[1]);
    std::vector<double> cov((size_t)nd * nd, 0.);
    for (int i = 0; i < nd; ++i) cov[(size_t)nd * i + i] = (i + 1.0) * (i + 1.0);
    d->setDistParamsStd(cov);
    const auto sd = d->getDistParamsStd();
    std::printf("dist %s %d %.17g %.17g %.17g %.17g %.17g %.17g %zu", names[m], nd, o[0], o[1], o3[0], o3[1], u[0], u[1], sd.size());
    for (double x : sd) std::printf(" %.17g", x);
    for (double x : back) std::printf(" %.17g", x);
    std::printf("\n");
  }
  try { makeDistParam("fisheye62", true); std::printf("unknown-model FAILED\n"); } catch (const std::invalid_argument&) { std::printf("unknown-model ok\n"); }
  Camera cam("/data/cam0/", true, "radtan5", 70.0);
  cam.setupInitialCalibration(1280, 1024);
  const IntrinsicsParam& K = *cam.m_pcalib_params;
  const Vector3d c = K.projImageToCam(100.25, 700.5);
  const auto px = K.projCamToImage(Vector3d{0.2, -0.1, 2.0});
  std::printf("camera %.17g %.17g %.17g %zu %.17g %.17g %.17g %.17g %d %d %d\n", K.getFocal()[0], K.getPrincipalPoint()[0], K.getPrincipalPoint()[1],
              K.getFocal().size(), c.x, c.y, px[0], px[1], (int)K.isPointInImage(1279.4f, 10.f), (int)K.isPointInImage(1279.5f, 10.f),
              (int)K.isPointInImage(-0.5, 3.0));
  const SE3d T = se3::exp({0.1, -0.2, 0.3, 0.2, -0.1, 0.4});
  const auto lg = se3::log(se3::mul(T, se3::inverse(se3::exp({0.0, 0.0, 0.0, 0.0, 0.0, 1e-12}))));
  std::printf("se3 %.17g %.17g %.17g %.17g %.17g %.17g %.17g", T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]);
  for (double x : lg) std::printf(" %.17g", x);
  std::printf("\n");
  return 0;
}

// A 2-camera calibration through the mirror classes: Camera -> setupInitialCalibration -> computeP3PRansac ->
// computeCalibration(cam) per camera -> runMultiCameraCalib.  argv[2] = binary file:
//   int32 n_cams, n_frames, n_pts, width, height, mono ; double prior_fov ; char model[32]
//   double target[n_pts][3] ; float corners[n_cams][n_frames][n_pts][2]
static int rig(const char* path) {
  std::ifstream f(path, std::ios::binary);
  int32_t h[6];
  rd(f, h, 6);
  double fov;
  rd(f, &fov, 1);
  char model[32];
  rd(f, model, 32);
  const int n_cams = h[0], n_frames = h[1], n_pts = h[2];
  std::vector<Vector3d> target(n_pts);
  rd(f, &target[0].x, (size_t)n_pts * 3);
  Context ctx(0);
  std::vector<Camera> cams;
  for (int c = 0; c < n_cams; ++c) {
    cams.emplace_back("/data/cam" + std::to_string(c) + "/", h[5] != 0, std::string(model), fov);
    Camera& cam = cams.back();
    cam.m_id = c;
    cam.setupInitialCalibration(h[3], h[4]);
    std::vector<std::vector<Point2f>> corners(n_frames, std::vector<Point2f>(n_pts));
    for (auto& v : corners) rd(f, &v[0].x, (size_t)n_pts * 2);
    std::vector<SE3d> poses;
    std::vector<int32_t> ninl;
    const std::vector<uint8_t> ok = computeP3PRansac(ctx, corners, target, *cam.m_pcalib_params, poses, &ninl);
    for (int i = 0; i < n_frames; ++i)
      if (ok[i]) {
        char name[32];
        std::snprintf(name, sizeof(name), "%06d.png", i);
        cam.m_v_calib_frames.emplace_back(name, (double)i, corners[i], poses[i]);
      }
    const CalibResult r = computeCalibration(ctx, cam, target);
    std::printf("rig cam %d frames %zu focal %.17g pp %.17g %.17g dist", c, cam.m_v_calib_frames.size(), cam.m_pcalib_params->getFocal()[0],
                cam.m_pcalib_params->getPrincipalPoint()[0], cam.m_pcalib_params->getPrincipalPoint()[1]);
    for (double v : cam.m_pdist_params->getDistParameters()) std::printf(" %.17g", v);
    std::printf(" cov %d nstd %zu\n", (int)r.covariance_ok, cam.m_pdist_params->getDistParamsStd().size());
  }
  const MultiCamResult mr = runMultiCameraCalib(ctx, cams, target);
  std::printf("rig pairs %zu iterations %d", mr.nb_good_pairs, mr.summary.iterations);
  for (int c = 1; c < n_cams; ++c) {
    const SE3d& T = cams[c].m_T_cam0_2_cam;
    std::printf(" ext %.17g %.17g %.17g %.17g %.17g %.17g %.17g", T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]);
  }
  std::printf("\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  if (std::string(argv[1]) == "--cpu-mirror") return cpu_mirror();
  if (std::string(argv[1]) == "--rig") {
    if (argc < 3) return 2;
    try { return rig(argv[2]); } catch (const std::exception& e) { std::fprintf(stderr, "exception: %s\n", e.what()); return 1; }
  }
  std::ifstream f(argv[1], std::ios::binary);
  try {
    Context ctx(0);
    int32_t hdr[3];
    rd(f, hdr, 3);
    std::vector<uint8_t> cb((size_t)hdr[0] * hdr[1] * hdr[2]);
    rd(f, cb.data(), cb.size());
    {
      ImageBatch imgs(ctx, cb.data(), hdr[0], hdr[1], hdr[2]);
      ChessboardCalibDetector det(7, 10, 0.05);
      det.detectBatch(ctx, imgs);
      std::printf("chessboard target %zu %.6f %.6f %.6f\n", det.getTargetCoords().size(), det.getTargetCoords()[0].x, det.getTargetCoords()[0].y,
                  det.getTargetCoords()[0].z);
      for (int i = 0; i < imgs.size(); ++i) {
        std::vector<Point2f> pts;
        const bool ok = det.detectTarget(i, pts);
        std::printf("chessboard %d %d", i, ok ? 1 : 0);
        for (const Point2f& p : pts) std::printf(" %.9g %.9g", p.x, p.y);
        std::printf("\n");
      }
    }
    rd(f, hdr, 3);
    std::vector<uint8_t> ag((size_t)hdr[0] * hdr[1] * hdr[2]);
    rd(f, ag.data(), ag.size());
    {
      ImageBatch imgs(ctx, ag.data(), hdr[0], hdr[1], hdr[2], 0, 0, /*async=*/true);
      AprilTagCalibDetector det(6, 6, 0.06, 0.3);
      det.detectBatch(ctx, imgs);
      for (int i = 0; i < imgs.size(); ++i) {
        std::vector<Point2f> pts;
        const bool ok = det.detectTarget(i, pts);
        std::printf("aprilgrid %d %d %d", i, ok ? 1 : 0, det.nbDetections(i));
        for (const Point2f& p : pts) std::printf(" %.9g %.9g", p.x, p.y);
        std::printf("\n");
      }
    }
    int32_t lh[4];
    rd(f, lh, 4);
    double wh[2], focal[2], pp[2], dist[8];
    rd(f, wh, 2); rd(f, focal, 2); rd(f, pp, 2); rd(f, dist, 8);
    const int n_views = lh[2], n_pts = lh[3];
    std::vector<Vector3d> target(n_pts);
    rd(f, &target[0].x, (size_t)n_pts * 3);
    std::vector<float> obs((size_t)n_views * n_pts * 2);
    rd(f, obs.data(), obs.size());
    std::vector<double> poses((size_t)n_views * 7);
    rd(f, poses.data(), poses.size());
    std::vector<CalibFrame> frames(n_views);
    for (int i = 0; i < n_views; ++i) {
      frames[i].m_v_corners_px.assign(reinterpret_cast<const Point2f*>(&obs[(size_t)i * n_pts * 2]),
                                      reinterpret_cast<const Point2f*>(&obs[(size_t)(i + 1) * n_pts * 2]));
      for (int j = 0; j < 4; ++j) frames[i].m_T_world_2_cam.q[j] = poses[7 * i + j];
      for (int j = 0; j < 3; ++j) frames[i].m_T_world_2_cam.t[j] = poses[7 * i + 4 + j];
    }
    const int nd = ezc_num_dist(lh[0]);
    CalibResult r = computeCalibration(ctx, (DistModel)lh[0], lh[1] != 0, frames, target, wh[0], wh[1], std::vector<double>(focal, focal + 2),
                                       std::vector<double>(pp, pp + 2), std::vector<double>(dist, dist + nd));
    std::printf("calib focal");
    for (double v : r.focal) std::printf(" %.17g", v);
    std::printf(" pp %.17g %.17g dist", r.pp[0], r.pp[1]);
    for (double v : r.dist) std::printf(" %.17g", v);
    std::printf("\ncalib iterations %d %d %d final_cost %.17g\n", r.stage[0].iterations, r.stage[1].iterations, r.stage[2].iterations, r.stage[2].final_cost);
    std::printf("calib std");
    for (double v : r.focal_std) std::printf(" %.17g", v);
    for (double v : r.pp_std) std::printf(" %.17g", v);
    for (double v : r.dist_std) std::printf(" %.17g", v);
    std::printf("\ncalib rmse0 %.17g pose0", frames[0].rmse_err);
    for (int j = 0; j < 4; ++j) std::printf(" %.17g", frames[0].m_T_world_2_cam.q[j]);
    for (int j = 0; j < 3; ++j) std::printf(" %.17g", frames[0].m_T_world_2_cam.t[j]);
    std::printf("\n");
    // error behaviour: a bad argument throws, it does not exit
    try {
      ChessboardCalibDetector bad(0, 10, 0.05);
      std::printf("error-check FAILED\n");
    } catch (const std::invalid_argument&) {
      std::printf("error-check ok\n");
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
