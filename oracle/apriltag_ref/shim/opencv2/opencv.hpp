// Minimal stand-in for the parts of OpenCV that the reference's AprilTag sources use ON the
// extractTags path: cv::Mat{data,cols,rows,at<double>}, cv::Point2f/Point3f, cv::findHomography for
// exactly 4 correspondences (TEST INFRASTRUCTURE; OpenCV C++ headers are not in this image).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace cv {

struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Point3f { float x, y, z; Point3f() : x(0), y(0), z(0) {} Point3f(float a, float b, float c) : x(a), y(b), z(c) {} };

struct Mat {
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  std::vector<double> store;  // for the 3x3 double result of findHomography
  Mat() {}
  Mat(int r, int c, unsigned char* d) : rows(r), cols(c), data(d) {}
  template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(store.data())[r * cols + c]; }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(store.data())[r * cols + c]; }
};

// Exact homography through 4 point pairs (h33 = 1): 8x8 Gaussian elimination with partial pivoting.
// In the reference the matrix is only used to pick the quad corner nearest to H*(-1,-1)
// (TagDetector.cc:512-524), a choice that is insensitive to the solver's rounding.
inline Mat findHomography(const std::vector<Point2f>& s, const std::vector<Point2f>& d) {
  double A[8][9];
  for (int i = 0; i < 4; ++i) {
    const double x = s[i].x, y = s[i].y, u = d[i].x, v = d[i].y;
    double r0[9] = {x, y, 1, 0, 0, 0, -u * x, -u * y, u};
    double r1[9] = {0, 0, 0, x, y, 1, -v * x, -v * y, v};
    std::memcpy(A[2 * i], r0, sizeof(r0));
    std::memcpy(A[2 * i + 1], r1, sizeof(r1));
  }
  for (int c = 0; c < 8; ++c) {
    int p = c;
    for (int r = c + 1; r < 8; ++r) if (std::fabs(A[r][c]) > std::fabs(A[p][c])) p = r;
    if (p != c) for (int k = 0; k < 9; ++k) std::swap(A[p][k], A[c][k]);
    const double piv = A[c][c];
    for (int r = 0; r < 8; ++r) {
      if (r == c || piv == 0.0) continue;
      const double f = A[r][c] / piv;
      for (int k = c; k < 9; ++k) A[r][k] -= f * A[c][k];
    }
  }
  Mat H;
  H.rows = H.cols = 3;
  H.store.assign(9, 0.0);
  for (int i = 0; i < 8; ++i) H.store[i] = A[i][i] != 0.0 ? A[i][8] / A[i][i] : 0.0;
  H.store[8] = 1.0;
  return H;
}

}  // namespace cv
