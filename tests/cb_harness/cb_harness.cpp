// Host harness around ezcalib_b200/csrc/cb_logic.h (the quad/grid logic the CUDA chessboard detector runs with one
// thread per image), so that the SAME code can be checked against the real OpenCV on a CPU (tests/test_cb_logic.py).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../ezcalib_b200/csrc/cb_logic.h"

extern "C" {

int cbh_approx_poly_dp(const int32_t* pts, int n, double eps, int32_t* out) {
  std::vector<cb::Pt> src(n), dst(n + 4);
  std::vector<int> stack(4 * n + 16);
  for (int i = 0; i < n; ++i) { src[i].x = pts[2 * i]; src[i].y = pts[2 * i + 1]; }
  const int m = cb::approx_poly_dp_closed(src.data(), n, dst.data(), eps, stack.data());
  for (int i = 0; i < m; ++i) { out[2 * i] = dst[i].x; out[2 * i + 1] = dst[i].y; }
  return m;
}

// contours: hole contours in the order generateQuads visits them (descending cv::findContours index),
// flat int32 (x,y) with lengths[] and parent[] per contour.  Returns found (0/1); corners_out [pw*ph][2].
int cbh_detect_from_contours(const int32_t* pts, const int32_t* lengths, const int32_t* parent, int n_contours, int pw, int ph, int dilations,
                             float* corners_out, int32_t* n_corners_out, int32_t* prev_sqr_size, int32_t* n_quads_out) {
  static cb::Work w;  // large: keep off the stack
  std::vector<cb::Pt> quads_pt;
  std::vector<int> quads_parent;
  std::vector<int> counter;
  int boardIdx = -1;
  size_t off = 0;
  std::vector<cb::Pt> c, ta, tb;
  std::vector<int> stack;
  for (int i = 0; i < n_contours; ++i) {
    const int n = lengths[i];
    c.resize(n); ta.resize(n + 4); tb.resize(n + 4); stack.resize(4 * n + 16);
    for (int k = 0; k < n; ++k) { c[k].x = pts[2 * (off + k)]; c[k].y = pts[2 * (off + k) + 1]; }
    off += n;
    cb::Pt q[4];
    if (!cb::contour_to_quad(c.data(), n, ta.data(), tb.data(), stack.data(), q)) continue;
    const int p = parent[i];
    if ((int)counter.size() <= p) counter.resize(p + 1, 0);
    counter[p]++;
    if (boardIdx != p && (boardIdx < 0 || counter[boardIdx] < counter[p])) boardIdx = p;
    for (int k = 0; k < 4; ++k) quads_pt.push_back(q[k]);
    quads_parent.push_back(p);
  }
  const int total = (int)quads_parent.size();
  w.n_quads = 0;
  w.max_quads = total * 3 > 2 ? total * 3 : 2;
  if (w.max_quads > cb::kMaxQuads) w.max_quads = cb::kMaxQuads;
  w.pw = pw; w.ph = ph;
  for (int i = 0; i < total; ++i) {
    if (quads_parent[i] != boardIdx) continue;
    if (w.n_quads >= w.max_quads) break;
    cb::add_quad(w, &quads_pt[4 * i], dilations);
  }
  *n_quads_out = w.n_quads;
  std::vector<cb::Ptf> out(cb::kMaxQuads * 4);
  std::vector<int> scratch(cb::kMaxQuads * 4);
  int n_out = 0;
  const bool found = cb::process_quads(w, out.data(), &n_out, prev_sqr_size, scratch.data());
  *n_corners_out = n_out;
  for (int i = 0; i < n_out && i < pw * ph; ++i) { corners_out[2 * i] = out[i].x; corners_out[2 * i + 1] = out[i].y; }
  return found ? 1 : 0;
}

int cbh_check_monotony(const float* corners, int pw, int ph) {
  return cb::check_board_monotony(reinterpret_cast<const cb::Ptf*>(corners), pw, ph) ? 1 : 0;
}

}  // extern "C"
