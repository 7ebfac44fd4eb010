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

#include "ezcalib_b200.hpp"

using namespace ezcalib_b200;

template <class T> static void rd(std::ifstream& f, T* p, size_t n) {
  f.read(reinterpret_cast<char*>(p), sizeof(T) * n);
  if (!f) { std::fprintf(stderr, "short read\n"); std::exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
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
