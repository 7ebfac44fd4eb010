// cb_logic.h -- the quad / grid logic of cv::findChessboardCorners, written for one device thread per image
// (plain arrays and indices instead of pointers, fixed capacities).  It is shared, unchanged, by the CUDA detector
// (csrc/chessboard.cu) and by a host harness (tests) so that it can be checked against the real OpenCV on a CPU.
//
// Replaces what the reference obtains from cv::findChessboardCorners at
// /root/reference/include/target_detectors/chessboard_detector.hpp:30-35.  The algorithm lives in OpenCV (un-vendored
// dependency): modules/calib3d/src/calibinit.cpp -- ChessBoardDetector::generateQuads / findQuadNeighbors /
// findConnectedQuads / orderFoundConnectedQuads / cleanFoundConnectedQuads / checkQuadGroup / checkBoardMonotony -- and
// modules/imgproc/src/approx.cpp (approxPolyDP).  Restated from the published algorithm; pinned by tests against cv2.
#pragma once
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define CB_HD __host__ __device__
#else
#define CB_HD
#endif

namespace cb {

constexpr int kMaxQuads = 1536;        // quad buffer (OpenCV: max(2, 3 * quads found))
constexpr int kMaxContourPts = 4096;   // vertices of one CHAIN_APPROX_SIMPLE contour handed to approxPolyDP
constexpr int kMaxApproxLevel = 7;     // MAX_CONTOUR_APPROX

struct Pt { int x, y; };
struct Ptf { float x, y; };

struct Corner {
  Ptf pt;
  int row;
  int count;
  int nb[4];
};

struct Quad {
  int count;
  int group_idx;
  int row, col;
  int ordered;
  float edge_len;
  int corners[4];   // indices into Work::corners
  int nb[4];        // indices into Work::quads, -1 = none
};

struct Work {
  Quad quads[kMaxQuads];
  Corner corners[kMaxQuads * 4];
  int n_quads;          // all_quads_count
  int max_quads;        // all_quads.size()
  int group[kMaxQuads];
  int n_group;
  int corner_group[kMaxQuads * 4];
  int stack[kMaxQuads];
  Ptf centers[kMaxQuads];
  Ptf hull_in[kMaxQuads];
  Ptf hull[kMaxQuads];
  int pw, ph;           // pattern_size.width / height (inner corners)
  int scan_pos;         // findConnectedQuads resumes its seed search here (quads before it are labelled or isolated for good)
};

CB_HD inline float normL2Sqrf(float dx, float dy) { return dx * dx + dy * dy; }

// ------------------------------------------------------------------------------------------ approxPolyDP (closed)
// cv::approxPolyDP(contour, out, eps, true) for integer points; src != dst; returns the new count.
// `stack` must hold at least 2*count ints pairs; we use a caller-provided int buffer of size >= 4*count + 16.
CB_HD inline int approx_poly_dp_closed(const Pt* src, int count, Pt* dst, double eps, int* stack_buf) {
  if (count == 0) return 0;
  int top = 0;
  auto push = [&](int s, int e) { stack_buf[2 * top] = s; stack_buf[2 * top + 1] = e; ++top; };
  const int init_iters = 3;
  int slice_s = 0, slice_e = 0, right_s = 0, right_e = 0;
  Pt start_pt = {-1000000, -1000000}, end_pt = {0, 0}, pt = {0, 0};
  int pos = 0, new_count = 0;
  bool le_eps = false;
  eps *= eps;
  // 1. two approximately farthest points
  right_s = 0;
  for (int i = 0; i < init_iters; i++) {
    double max_dist = 0;
    pos = (pos + right_s) % count;
    start_pt = src[pos]; if (++pos >= count) pos = 0;
    for (int j = 1; j < count; j++) {
      pt = src[pos]; if (++pos >= count) pos = 0;
      const double dx = pt.x - start_pt.x, dy = pt.y - start_pt.y;
      const double dist = dx * dx + dy * dy;
      if (dist > max_dist) { max_dist = dist; right_s = j; }
    }
    le_eps = max_dist <= eps;
  }
  // 2. initialise the stack
  if (!le_eps) {
    right_e = slice_s = pos % count;
    slice_e = right_s = (right_s + slice_s) % count;
    push(right_s, right_e);
    push(slice_s, slice_e);
  } else {
    dst[new_count++] = start_pt;
  }
  // 3. recursive subdivision
  while (top > 0) {
    --top;
    slice_s = stack_buf[2 * top]; slice_e = stack_buf[2 * top + 1];
    end_pt = src[slice_e];
    pos = slice_s;
    start_pt = src[pos]; if (++pos >= count) pos = 0;
    if (pos != slice_e) {
      // OpenCV >= 4.9 measures the distance to the SEGMENT [start,end] (projection clamped to the end points),
      // not to the infinite line; verified against cv2 4.13 on 15k random contours (0 mismatches).
      double max_dist = 0;
      const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
      const double L2 = dx * dx + dy * dy;
      while (pos != slice_e) {
        pt = src[pos]; if (++pos >= count) pos = 0;
        const double px = pt.x - start_pt.x, py = pt.y - start_pt.y;
        const double t = px * dx + py * dy;
        double dist;
        if (t < 0) dist = px * px + py * py;
        else if (t > L2) { const double ex = pt.x - end_pt.x, ey = pt.y - end_pt.y; dist = ex * ex + ey * ey; }
        else { const double cr = py * dx - px * dy; dist = cr * cr / L2; }
        if (dist > max_dist) { max_dist = dist; right_s = (pos + count - 1) % count; }
      }
      le_eps = max_dist <= eps;
    } else {
      le_eps = true;
      start_pt = src[slice_s];
    }
    if (le_eps) {
      dst[new_count++] = start_pt;
    } else {
      right_e = slice_e;
      slice_e = right_s;
      push(right_s, right_e);
      push(slice_s, slice_e);
    }
  }
  // last stage: remove extra points on [almost] straight lines
  count = new_count;
  pos = count - 1;
  start_pt = dst[pos]; if (++pos >= count) pos = 0;
  int wpos = pos;
  pt = dst[pos]; if (++pos >= count) pos = 0;
  for (int i = 0; i < count && new_count > 2; i++) {
    end_pt = dst[pos]; if (++pos >= count) pos = 0;
    const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
    const double dist = fabs((pt.x - start_pt.x) * dy - (pt.y - start_pt.y) * dx);
    const double sip = (double)(pt.x - start_pt.x) * (end_pt.x - pt.x) + (double)(pt.y - start_pt.y) * (end_pt.y - pt.y);
    if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
      new_count--;
      dst[wpos] = start_pt = end_pt;
      if (++wpos >= count) wpos = 0;
      pt = dst[pos]; if (++pos >= count) pos = 0;
      i++;
      continue;
    }
    dst[wpos] = start_pt = pt;
    if (++wpos >= count) wpos = 0;
    pt = end_pt;
  }
  return new_count;
}

// cv::isContourConvex for 4 integer points
CB_HD inline bool is_convex4(const Pt* p) {
  Pt prev = p[3], cur = p[0];
  int dx0 = cur.x - prev.x, dy0 = cur.y - prev.y;
  int orientation = 0;
  for (int i = 0; i < 4; i++) {
    prev = cur;
    cur = p[(i + 1) & 3];
    const int dx = cur.x - prev.x, dy = cur.y - prev.y;
    const long long dxdy0 = (long long)dx * dy0, dydx0 = (long long)dy * dx0;
    orientation |= (dydx0 > dxdy0) ? 1 : ((dydx0 < dxdy0) ? 2 : 3);
    if (orientation == 3) return false;
    dx0 = dx; dy0 = dy;
  }
  return true;
}

// ------------------------------------------------------------------------------------------ generateQuads (per contour)
// Returns true and fills q[4] if the hole contour yields an acceptable quadrangle (approxPolyDP levels 1..7 applied
// twice, convexity, CALIB_CB_FILTER_QUADS size / aspect tests).
// Convexity + CALIB_CB_FILTER_QUADS tests on a 4-vertex approximation; fills q on success.
CB_HD inline bool quad_from_4(const Pt* tmp_a, Pt q[4]) {
  if (!is_convex4(tmp_a)) return false;
  for (int i = 0; i < 4; ++i) q[i] = tmp_a[i];
  // CALIB_CB_FILTER_QUADS
  double p = 0;  // cv::arcLength(closed): float accumulation of per-edge lengths in OpenCV; double is used here
  {
    float perim = 0;
    Pt prev = q[3];
    for (int i = 0; i < 4; ++i) {
      const float dx = (float)q[i].x - (float)prev.x, dy = (float)q[i].y - (float)prev.y;
      perim += sqrtf(dx * dx + dy * dy);
      prev = q[i];
    }
    p = perim;
  }
  double a00 = 0;  // cv::contourArea(oriented=false)
  {
    Ptf prev = {(float)q[3].x, (float)q[3].y};
    for (int i = 0; i < 4; ++i) {
      const Ptf cur = {(float)q[i].x, (float)q[i].y};
      a00 += (double)prev.x * cur.y - (double)prev.y * cur.x;
      prev = cur;
    }
    a00 = fabs(a00 * 0.5);
  }
  const double area = a00;
  auto d = [&](int i, int j) { const double dx = q[i].x - q[j].x, dy = q[i].y - q[j].y; return sqrt(dx * dx + dy * dy); };
  const double d1 = d(0, 2), d2 = d(1, 3), d3 = d(0, 1), d4 = d(1, 2);
  if (!(d3 * 4 > d4 && d4 * 4 > d3 && d3 * d4 < area * 1.5 && area > 25 /* min_area */ && d1 >= 0.15 * p && d2 >= 0.15 * p)) return false;
  return true;
}

CB_HD inline bool contour_to_quad(const Pt* contour, int n, Pt* tmp_a, Pt* tmp_b, int* stack_buf, Pt q[4]) {
  // bounding rect area >= 25
  int minx = contour[0].x, maxx = minx, miny = contour[0].y, maxy = miny;
  for (int i = 1; i < n; ++i) {
    minx = contour[i].x < minx ? contour[i].x : minx; maxx = contour[i].x > maxx ? contour[i].x : maxx;
    miny = contour[i].y < miny ? contour[i].y : miny; maxy = contour[i].y > maxy ? contour[i].y : maxy;
  }
  const int min_area = 25;
  if ((maxx - minx + 1) * (maxy - miny + 1) < min_area) return false;
  // OpenCV 4.13 behaviour, pinned on refined corners that are bit-identical to cv2's only this way (tests/test_cb_logic.py):
  // every level approximates twice -- the contour, then its own output -- and a 4-gon from either call ends the search;
  // the NEXT level starts from the output of this level's FIRST call (not from the original contour, and not from the
  // second call's output).  Two buffers suffice: the second call may overwrite the first call's input.
  const Pt* src = contour;
  int nsrc = n;
  Pt* A = tmp_a;
  Pt* B = tmp_b;
  for (int level = 1; level <= kMaxApproxLevel; level++) {
    const int nf = approx_poly_dp_closed(src, nsrc, A, (double)(float)level, stack_buf);
    if (nf == 4) return quad_from_4(A, q);
    const int ns = approx_poly_dp_closed(A, nf, B, (double)(float)level, stack_buf);
    if (ns == 4) return quad_from_4(B, q);
    src = A; nsrc = nf;
    Pt* t = A; A = B; B = t;
  }
  return false;
}

// add a quad (generateQuads, second loop)
// quad `qi` of the work set from an approximated contour (generateQuads' per-quad part); independent per quad
CB_HD inline void add_quad_at(Work& w, int qi, const Pt q[4], int dilations) {
  Quad& Q = w.quads[qi];
  Q.count = 0; Q.group_idx = -1; Q.row = 0; Q.col = 0; Q.ordered = 0;
  for (int i = 0; i < 4; ++i) {
    Corner& c = w.corners[qi * 4 + i];
    c.pt.x = (float)q[i].x; c.pt.y = (float)q[i].y;
    c.row = 0; c.count = 0;
    c.nb[0] = c.nb[1] = c.nb[2] = c.nb[3] = -1;
    Q.corners[i] = qi * 4 + i;
    Q.nb[i] = -1;
  }
  Q.edge_len = FLT_MAX;
  for (int i = 0; i < 4; ++i) {
    const Ptf a = w.corners[Q.corners[i]].pt, b = w.corners[Q.corners[(i + 1) & 3]].pt;
    const float d2 = normL2Sqrf(a.x - b.x, a.y - b.y);
#ifdef CB_EDGE_MAX
    if (i == 0) Q.edge_len = d2; else Q.edge_len = d2 > Q.edge_len ? d2 : Q.edge_len;
#else
    Q.edge_len = d2 < Q.edge_len ? d2 : Q.edge_len;
#endif
  }
  // OpenCV >= 4.6: the (squared) edge length is compensated for the shrinkage caused by the dilations of this attempt
  const int comp = 2 * dilations;
  Q.edge_len += 2 * sqrtf(Q.edge_len) * comp + comp * comp;
}
CB_HD inline void add_quad(Work& w, const Pt q[4], int dilations) { add_quad_at(w, w.n_quads++, q, dilations); }

// ------------------------------------------------------------------------------------------ findQuadNeighbors
// `Par` lets a warp share the two O(n_quads) scans of every (quad, corner) step without changing the sequential
// semantics: candidates are strided over the lanes and reduced with the tie-break of the sequential loop (first in
// (k, j) order among equal distances); the state is written by lane 0.  SerialPar (host, tests) is the plain loop.
struct SerialPar {
  CB_HD int lane() const { return 0; }
  CB_HD int lanes() const { return 1; }
  CB_HD void sync() const {}
  CB_HD void reduce_min(float&, int&) const {}
  CB_HD bool any(bool v) const { return v; }
};

// Linking rule of findQuadNeighbors as cv2 >= 4.9 behaves (measured on cv2 4.13): a corner of another quad
// is a candidate if its squared distance is within kNbSqrScale x the squared shortest edge of both quads AND it is a
// DIAGONAL neighbour: the candidate corner and the candidate quad's opposite corner both lie in the same quadrant of this
// quad (w.r.t. the lines through the mid points of its opposite edges) as the corner being matched.  The same test
// restricts which third corners can veto a link.
constexpr float kNbSqrScale = 2.f;
CB_HD inline Ptf mid_pt(const Ptf& a, const Ptf& b) { return Ptf{(a.x + b.x) / 2, (a.y + b.y) / 2}; }
CB_HD inline bool same_side_of_line(const Ptf& l1, const Ptf& l2, const Ptf& p1, const Ptf& p2) {
  const float dx = l2.x - l1.x, dy = l2.y - l1.y;
  const float a = dx * (p1.y - l1.y) - dy * (p1.x - l1.x);
  const float b = dx * (p2.y - l1.y) - dy * (p2.x - l1.x);
  return a * b > 0;
}
CB_HD inline bool diagonal_neighbour(const Work& w, const Quad& q, int i, const Ptf& other, const Ptf& other_diag) {
  const Ptf c0 = w.corners[q.corners[i]].pt, c1 = w.corners[q.corners[(i + 1) & 3]].pt, c2 = w.corners[q.corners[(i + 2) & 3]].pt,
            c3 = w.corners[q.corners[(i + 3) & 3]].pt;
  const Ptf m1 = mid_pt(c0, c1), m2 = mid_pt(c2, c3), m3 = mid_pt(c1, c2), m4 = mid_pt(c3, c0);
  return same_side_of_line(m1, m2, c0, other) && same_side_of_line(m3, m4, c0, other) && same_side_of_line(m1, m2, c0, other_diag) &&
         same_side_of_line(m3, m4, c0, other_diag);
}

template <class Par>
CB_HD inline void find_quad_neighbors(Work& w, const Par& par) {
  const int nq = w.n_quads;
  for (int idx = 0; idx < nq; idx++) {
    Quad& cur = w.quads[idx];
    for (int i = 0; i < 4; i++) {
      if (cur.nb[i] >= 0) continue;
      float min_dist = FLT_MAX;
      int code = 0x7fffffff;  // k * 4 + j of the closest admissible corner
      const Ptf pt = w.corners[cur.corners[i]].pt;
      const float cur_edge = cur.edge_len;
      for (int k = par.lane(); k < nq; k += par.lanes()) {
        if (k == idx) continue;
        const Quad& qk = w.quads[k];
        for (int j = 0; j < 4; j++) {
          if (qk.nb[j] >= 0) continue;
          const Ptf o = w.corners[qk.corners[j]].pt;
          const float dist = normL2Sqrf(pt.x - o.x, pt.y - o.y);
          if (dist < min_dist && dist <= cur_edge * kNbSqrScale && dist <= qk.edge_len * kNbSqrScale) {
            // edges that differ by more than 1:4 in length (16 in squared length) are incompatible
            if (qk.edge_len > 16 * cur_edge || cur_edge > 16 * qk.edge_len) continue;
            if (!diagonal_neighbour(w, cur, i, o, w.corners[qk.corners[(j + 2) & 3]].pt)) continue;
            // ... and also from the candidate quad's side unless the candidate is much smaller than this quad (squared edge
            // ratio below 1 : kNbSqrScale).  Empirical, like everything around it (OpenCV's source is not in the image):
            //   * a 0.3-scale board needs the candidate-side test (two squares of one row 9 px apart pass the one-sided
            //     test and scramble the corner order);
            //   * noisy views at 5-7 dilations need it NOT to apply when a large blob looks at its small side neighbour (cv2
            //     links them and finds no board);
            //   * round 2's randomised degradations (tests/cb_fuzz.py) showed that "candidate at least as large" was too
            //     narrow once the fallback passes the dilation count to generateQuads: candidates up to sqrt(2) smaller are
            //     tested too -- 26 of 35 residual views then agree with cv2 (15 before), every older fixture still does.
            if (qk.edge_len * kNbSqrScale >= cur_edge && !diagonal_neighbour(w, qk, j, pt, w.corners[cur.corners[(i + 2) & 3]].pt)) continue;
            code = k * 4 + j; min_dist = dist;
          }
        }
      }
      par.reduce_min(min_dist, code);
      if (code != 0x7fffffff && min_dist < FLT_MAX) {
        const int closest_quad = code >> 2, closest_corner_idx = code & 3;
        Quad& cq = w.quads[closest_quad];
#ifdef CB_DEBUG2
#define CBD(msg) fprintf(stderr, "   q%d c%d (%.1f,%.1f) -> q%d c%d dist %.1f: %s\n", idx, i, pt.x, pt.y, closest_quad, closest_corner_idx, min_dist, msg)
#else
#define CBD(msg)
#endif
        if (cur.count >= 4 || cq.count >= 4) { CBD("count>=4"); continue; }
        Corner& cc = w.corners[cq.corners[closest_corner_idx]];
        int j = 0;
        for (; j < 4; j++) {
          if (cur.nb[j] == closest_quad) break;
          const Ptf o = w.corners[cur.corners[j]].pt;
          if (normL2Sqrf(cc.pt.x - o.x, cc.pt.y - o.y) < min_dist) break;
        }
        if (j < 4) { CBD("own corner closer / already neighbour"); continue; }
        for (j = 0; j < 4; j++) if (cq.nb[j] == idx) break;
        if (j < 4) { CBD("already linked back"); continue; }
        // is some free corner of a third quad closer to closest_corner than this one?
        bool blocked = false;
        for (int q3 = par.lane(); q3 < nq && !blocked; q3 += par.lanes()) {
          if (q3 == idx || q3 == closest_quad) continue;
          const Quad& q = w.quads[q3];
          for (int k = 0; k < 4; k++) {
            if (q.nb[k] < 0) {
              const Ptf o = w.corners[q.corners[k]].pt;
              if (normL2Sqrf(cc.pt.x - o.x, cc.pt.y - o.y) < min_dist &&
                  diagonal_neighbour(w, cq, closest_corner_idx, o, w.corners[q.corners[(k + 2) & 3]].pt)) { blocked = true; break; }
            }
          }
        }
        if (par.any(blocked)) { CBD("third corner closer"); continue; }
        CBD("LINK");
        if (par.lane() == 0) {
          cc.pt.x = (pt.x + cc.pt.x) * 0.5f;
          cc.pt.y = (pt.y + cc.pt.y) * 0.5f;
          cur.count++;
          cur.nb[i] = closest_quad;
          cur.corners[i] = cq.corners[closest_corner_idx];
          cq.count++;
          cq.nb[closest_corner_idx] = idx;
        }
        par.sync();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ findConnectedQuads
CB_HD inline void find_connected_quads(Work& w, int group_idx) {
  w.n_group = 0;
  int top = 0;
  // OpenCV rescans from 0 for every group; a quad passed over once (already labelled, or without neighbours) is passed
  // over forever -- labels are never removed and only quads of the group being processed gain neighbours -- so the scan
  // resumes where the previous seed was found (O(n) instead of O(n * groups) on noise with hundreds of groups).
  for (int i = w.scan_pos; i < w.n_quads; i++) {
    Quad* q = &w.quads[i];
    if (q->count <= 0 || q->group_idx >= 0) continue;
    w.scan_pos = i + 1;
    w.stack[top++] = i;
    w.group[w.n_group++] = i;
    q->group_idx = group_idx;
    q->ordered = 0;
    while (top > 0) {
      const int qi = w.stack[--top];
      for (int k = 0; k < 4; k++) {
        const int ni = w.quads[qi].nb[k];
        if (ni >= 0 && w.quads[ni].count > 0 && w.quads[ni].group_idx < 0) {
          w.stack[top++] = ni;
          w.group[w.n_group++] = ni;
          w.quads[ni].group_idx = group_idx;
          w.quads[ni].ordered = 0;
        }
      }
    }
    break;
  }
}

CB_HD inline void order_quad(Work& w, Quad& quad, const Corner& corner, int common) {
  int tc = 0;
  for (; tc < 4; ++tc) {
    const Ptf p = w.corners[quad.corners[tc]].pt;
    if (p.x == corner.pt.x && p.y == corner.pt.y) break;
  }
  while (tc != common) {
    const int tempc = quad.corners[3], tempq = quad.nb[3];
    for (int i = 3; i > 0; --i) { quad.corners[i] = quad.corners[i - 1]; quad.nb[i] = quad.nb[i - 1]; }
    quad.corners[0] = tempc; quad.nb[0] = tempq;
    tc = (tc + 1) & 3;
  }
}

CB_HD inline void remove_quad_from_group(Work& w, int q0) {
  const int count = w.n_group;
  for (int i = 0; i < count; ++i) {
    Quad& q = w.quads[w.group[i]];
    for (int j = 0; j < 4; j++) {
      if (q.nb[j] == q0) {
        q.nb[j] = -1;
        q.count--;
        for (int k = 0; k < 4; ++k) {
          if (w.quads[q0].nb[k] == w.group[i]) { w.quads[q0].nb[k] = -1; w.quads[q0].count--; break; }
        }
        break;
      }
    }
  }
  // remove the quad
  for (int i = 0; i < count; ++i) {
    if (w.group[i] == q0) { w.group[i] = w.group[count - 1]; break; }
  }
  w.n_group = count - 1;
}

CB_HD inline int add_outer_quad(Work& w, int quad_idx) {
  int added = 0;
  Quad& quad = w.quads[quad_idx];
  for (int i = 0; i < 4 && w.n_quads < w.max_quads; i++) {
    if (quad.nb[i] < 0) {
      const int j = (i + 2) & 3;
      const int q_index = w.n_quads++;
      Quad& q = w.quads[q_index];
      q.count = 0; q.group_idx = -1; q.row = 0; q.col = 0; q.ordered = 0; q.edge_len = 0;
      for (int k = 0; k < 4; ++k) { q.nb[k] = -1; q.corners[k] = -1; }
      added++;
      w.group[w.n_group++] = q_index;
      quad.nb[i] = q_index;
      quad.count += 1;
      q.nb[j] = quad_idx;
      q.group_idx = quad.group_idx;
      q.count = 1;
      q.ordered = 0;
      q.edge_len = quad.edge_len;
      const Ptf pi = w.corners[quad.corners[i]].pt, pj = w.corners[quad.corners[j]].pt;
      const Ptf off = {pi.x - pj.x, pi.y - pj.y};
      for (int k = 0; k < 4; k++) {
        Corner& c = w.corners[q_index * 4 + k];
        const Ptf p = w.corners[quad.corners[k]].pt;
        c.pt = p; c.row = 0; c.count = 0; c.nb[0] = c.nb[1] = c.nb[2] = c.nb[3] = -1;
        q.corners[k] = q_index * 4 + k;
        c.pt.x += off.x; c.pt.y += off.y;
      }
      q.corners[j] = quad.corners[i];
      // the second neighbour of the new quad, on either side (k = 1: the side OpenCV has always tried; k = 3: the mirror
      // case -- without it a missing edge square next to an unordered corner square gets TWO replacement quads)
      for (int k = 1; k <= 3; k += 2) {
        const int next_i = (i + k) & 3, prev_i = (i + k + 2) & 3;
        const int qp = quad.nb[prev_i];
        if (qp >= 0 && w.quads[qp].ordered && w.quads[qp].nb[i] >= 0 && w.quads[w.quads[qp].nb[i]].ordered) {
          const int qn = w.quads[qp].nb[i];
          q.count += 1;
          q.nb[prev_i] = qn;
          w.quads[qn].nb[next_i] = q_index;
          w.quads[qn].count += 1;
          q.corners[prev_i] = w.quads[qn].corners[next_i];
        }
      }
    }
  }
  return added;
}

// ------------------------------------------------------------------------------------------ orderFoundConnectedQuads
CB_HD inline int order_found_connected_quads(Work& w) {
  int quad_count = w.n_group;
  int start = -1;
  for (int i = 0; i < quad_count; i++) if (w.quads[w.group[i]].count == 4) { start = w.group[i]; break; }
  if (start < 0) return 0;
  int row_min = 0, col_min = 0, row_max = 0, col_max = 0;
  int top = 0;
  w.stack[top++] = start;
  w.quads[start].row = 0; w.quads[start].col = 0; w.quads[start].ordered = 1;
  while (top > 0) {
    const int qi = w.stack[--top];
    int col = w.quads[qi].col, row = w.quads[qi].row;
    if (row > row_max) row_max = row;
    if (row < row_min) row_min = row;
    if (col > col_max) col_max = col;
    if (col < col_min) col_min = col;
    for (int i = 0; i < 4; i++) {
      const int ni = w.quads[qi].nb[i];
      switch (i) {
        case 0: row--; col--; break;
        case 1: col += 2; break;
        case 2: row += 2; break;
        case 3: col -= 2; break;
      }
      if (ni >= 0 && !w.quads[ni].ordered && w.quads[ni].count == 4) {
        order_quad(w, w.quads[ni], w.corners[w.quads[qi].corners[i]], (i + 2) & 3);
        w.quads[ni].ordered = 1;
        w.quads[ni].row = row;
        w.quads[ni].col = col;
        w.stack[top++] = ni;
      }
    }
  }
  int wd = w.pw - 1, h = w.ph - 1;
  const int drow = row_max - row_min + 1, dcol = col_max - col_min + 1;
  if ((wd > h && dcol < drow) || (wd < h && drow < dcol)) { h = w.pw - 1; wd = w.ph - 1; }
  if (dcol < wd || drow < h) return 0;
  // check edges of inner quads; if there is an outer quad missing, fill it in
  int found = 0;
  for (int i = 0; i < quad_count; ++i) {
    Quad& q = w.quads[w.group[i]];
    if (q.count != 4) continue;
    int col = q.col, row = q.row;
    for (int j = 0; j < 4; j++) {
      switch (j) {
        case 0: row--; col--; break;
        case 1: col += 2; break;
        case 2: row += 2; break;
        case 3: col -= 2; break;
      }
      const int ni = q.nb[j];
      if (ni >= 0 && !w.quads[ni].ordered && col <= col_max && col >= col_min && row <= row_max && row >= row_min) {
        found++;
        order_quad(w, w.quads[ni], w.corners[q.corners[j]], (j + 2) & 3);
        w.quads[ni].ordered = 1;
        w.quads[ni].row = row;
        w.quads[ni].col = col;
      }
    }
  }
  if (found > 0) {
    for (int i = 0; i < quad_count && w.n_quads < w.max_quads; i++) {
      const int qi = w.group[i];
      if (w.quads[qi].count < 4 && w.quads[qi].ordered) {
        const int added = add_outer_quad(w, qi);
        quad_count += added;
      }
    }
    if (w.n_quads >= w.max_quads) return 0;
  }
  if (dcol == wd && drow == h) {
    for (int i = quad_count - 1; i >= 0; i--) {
      if (i >= w.n_group) continue;
      const int qi = w.group[i];
      if (!w.quads[qi].ordered) {
        bool outer = false;
        for (int j = 0; j < 4; j++) if (w.quads[qi].nb[j] >= 0 && w.quads[w.quads[qi].nb[j]].ordered) outer = true;
        if (!outer) remove_quad_from_group(w, qi);
      }
    }
    return w.n_group;
  }
  return 0;
}

// convex hull area of n float points (monotone chain; the area does not depend on the hull algorithm)
CB_HD inline double hull_area(Work& w, const Ptf* pts, int n) {
  // insertion sort by (x, y)
  Ptf* s = w.hull_in;
  for (int i = 0; i < n; ++i) {
    const Ptf v = pts[i];
    int j = i - 1;
    while (j >= 0 && (s[j].x > v.x || (s[j].x == v.x && s[j].y > v.y))) { s[j + 1] = s[j]; --j; }
    s[j + 1] = v;
  }
  Ptf* h = w.hull;
  int k = 0;
  auto cross = [](const Ptf& o, const Ptf& a, const Ptf& b) { return (double)(a.x - o.x) * (b.y - o.y) - (double)(a.y - o.y) * (b.x - o.x); };
  for (int i = 0; i < n; ++i) { while (k >= 2 && cross(h[k - 2], h[k - 1], s[i]) <= 0) k--; h[k++] = s[i]; }
  for (int i = n - 2, t = k + 1; i >= 0; --i) { while (k >= t && cross(h[k - 2], h[k - 1], s[i]) <= 0) k--; h[k++] = s[i]; }
  if (k > 1) k--;
  double a = 0;
  for (int i = 0; i < k; ++i) { const Ptf p = h[i], q = h[(i + 1) % k]; a += (double)p.x * q.y - (double)p.y * q.x; }
  return fabs(a * 0.5);
}

// ------------------------------------------------------------------------------------------ cleanFoundConnectedQuads
CB_HD inline int clean_found_connected_quads(Work& w) {
  const int count = ((w.pw + 1) * (w.ph + 1) + 1) / 2;
  int quad_count = w.n_group;
  if (quad_count <= count) return quad_count;
  Ptf center = {0, 0};
  for (int i = 0; i < quad_count; ++i) {
    const Quad& q = w.quads[w.group[i]];
    Ptf ci = {0, 0};
    for (int k = 0; k < 4; ++k) { ci.x += w.corners[q.corners[k]].pt.x; ci.y += w.corners[q.corners[k]].pt.y; }
    ci.x *= 0.25f; ci.y *= 0.25f;
    w.centers[i] = ci;
    center.x += ci.x; center.y += ci.y;
  }
  center.x *= (1.0f / quad_count); center.y *= (1.0f / quad_count);
  for (; quad_count > count; quad_count--) {
    double min_box_area = DBL_MAX;
    int min_idx = -1;
    for (int skip = 0; skip < quad_count; ++skip) {
      const Ptf temp = w.centers[skip];
      w.centers[skip] = center;
      const double a = hull_area(w, w.centers, quad_count);
      w.centers[skip] = temp;
      if (a < min_box_area) { min_box_area = a; min_idx = skip; }
    }
    const int q0 = w.group[min_idx];
    for (int i = 0; i < quad_count; ++i) {
      Quad& q = w.quads[w.group[i]];
      for (int j = 0; j < 4; ++j) {
        if (q.nb[j] == q0) {
          q.nb[j] = -1;
          q.count--;
          for (int k = 0; k < 4; ++k) if (w.quads[q0].nb[k] == w.group[i]) { w.quads[q0].nb[k] = -1; w.quads[q0].count--; break; }
          break;
        }
      }
    }
    // OpenCV decrements quad_count here AND in the for-increment (a long-standing quirk of calibinit.cpp): every
    // pass unlinks one quad by the hull criterion and silently drops one more from the tail of the group.
    quad_count--;
    w.group[min_idx] = w.group[quad_count];
    w.centers[min_idx] = w.centers[quad_count];
  }
  if (quad_count < 0) quad_count = 0;
  w.n_group = quad_count;
  return quad_count;
}

// ------------------------------------------------------------------------------------------ checkQuadGroup
// out = w.corner_group (indices of corners); returns count (>0 full board, <0 partial)
CB_HD inline int check_quad_group(Work& w, int* corners /*scratch [4*n_group]*/) {
  const int ROW1 = 1000000, ROW2 = 2000000, ROW_ = 3000000;
  const int quad_count = w.n_group;
  int corner_count = 0, result = 0;
  int width = 0, height = 0;
  int hist[5] = {0, 0, 0, 0, 0};
  int* out = w.corner_group;
  int n_out = 0;
  bool ok = true;
  for (int i = 0; i < quad_count && ok; ++i) {
    Quad& q = w.quads[w.group[i]];
    for (int j = 0; j < 4 && ok; ++j) {
      if (q.nb[j] >= 0) {
        const int next_j = (j + 1) & 3;
        const int ai = q.corners[j], bi = q.corners[next_j];
        Corner& a = w.corners[ai];
        Corner& b = w.corners[bi];
        const int row_flag = q.count == 1 ? ROW1 : q.count == 2 ? ROW2 : ROW_;
        if (a.row == 0) { corners[corner_count++] = ai; a.row = row_flag; }
        else if (a.row > row_flag) a.row = row_flag;
        if (q.nb[next_j] >= 0) {
          if (a.count >= 4 || b.count >= 4) { ok = false; break; }
          for (int k = 0; k < 4; ++k) if (a.nb[k] == bi || b.nb[k] == ai) { ok = false; break; }
          if (!ok) break;
          a.nb[a.count++] = bi;
          b.nb[b.count++] = ai;
        }
      }
    }
  }
  if (ok && corner_count != w.pw * w.ph) ok = false;
  if (ok) {
    int first = -1, first2 = -1;
    for (int i = 0; i < corner_count; ++i) {
      const int n = w.corners[corners[i]].count;
      hist[n]++;
      if (first < 0 && n == 2) {
        if (w.corners[corners[i]].row == ROW1) first = corners[i];
        else if (first2 < 0 && w.corners[corners[i]].row == ROW2) first2 = corners[i];
      }
    }
    if (first < 0) first = first2;
    if (first < 0 || hist[0] != 0 || hist[1] != 0 || hist[2] != 4 || hist[3] != (w.pw + w.ph) * 2 - 8) ok = false;
    int cur = first, right = -1, below = -1;
    if (ok) {
      n_out = 0;
      out[n_out++] = cur;
      for (int k = 0; k < 4; ++k) {
        const int c = w.corners[cur].nb[k];
        if (c >= 0) { if (right < 0) right = c; else if (below < 0) below = c; }
      }
      if (right < 0 || (w.corners[right].count != 2 && w.corners[right].count != 3) || below < 0 ||
          (w.corners[below].count != 2 && w.corners[below].count != 3))
        ok = false;
    }
    if (ok) {
      w.corners[cur].row = 0;
      first = below;
      for (;;) {
        w.corners[right].row = 0;
        out[n_out++] = right;
        if (w.corners[right].count == 2) break;
        if (w.corners[right].count != 3 || n_out >= (w.pw > w.ph ? w.pw : w.ph)) { ok = false; break; }
        cur = right;
        for (int k = 0; k < 4; ++k) {
          const int c = w.corners[cur].nb[k];
          if (c >= 0 && w.corners[c].row > 0) {
            int kk = 0;
            for (; kk < 4; ++kk) if (w.corners[c].nb[kk] == below) break;
            if (kk < 4) below = c; else right = c;
          }
        }
      }
    }
    if (ok) {
      width = n_out;
      if (width == w.pw) height = w.ph;
      else if (width == w.ph) height = w.pw;
      else ok = false;
    }
    if (ok) {
      for (int i = 1;; ++i) {
        if (first < 0) break;
        cur = first;
        first = -1;
        int j = 0;
        for (;; ++j) {
          w.corners[cur].row = i;
          out[n_out++] = cur;
          if (w.corners[cur].count == 2 + (i < height - 1) && j > 0) break;
          right = -1;
          for (int k = 0; k < 4; ++k) {
            const int c = w.corners[cur].nb[k];
            if (c >= 0 && w.corners[c].row > i) {
              int kk = 0;
              for (; kk < 4; ++kk) { const int cn = w.corners[c].nb[kk]; if (cn >= 0 && w.corners[cn].row == i - 1) break; }
              if (kk < 4) { right = c; if (j > 0) break; }
              else if (j == 0) first = c;
            }
          }
          if (right < 0) { ok = false; break; }
          cur = right;
          if (n_out >= kMaxQuads * 4 - 1) { ok = false; break; }
        }
        if (!ok) break;
        if (j != width - 1) { ok = false; break; }
      }
    }
    if (ok && n_out != corner_count) ok = false;
    if (ok) {
      if (width != w.pw) {
        const int t = width; width = height; height = t;
        int* tmp = corners;  // reuse scratch
        for (int i = 0; i < n_out; ++i) tmp[i] = out[i];
        for (int i = 0; i < height; ++i) for (int j = 0; j < width; ++j) out[i * width + j] = tmp[j * height + i];
      }
      const Ptf p0 = w.corners[out[0]].pt, p1 = w.corners[out[w.pw - 1]].pt, p2 = w.corners[out[w.pw]].pt;
      if ((p1.x - p0.x) * (p2.y - p1.y) - (p1.y - p0.y) * (p2.x - p1.x) < 0) {
        if (width % 2 == 0) {
          for (int i = 0; i < height; ++i) for (int j = 0; j < width / 2; ++j) { const int t = out[i * width + j]; out[i * width + j] = out[i * width + width - j - 1]; out[i * width + width - j - 1] = t; }
        } else {
          for (int j = 0; j < width; ++j) for (int i = 0; i < height / 2; ++i) { const int t = out[i * width + j]; out[i * width + j] = out[(height - i - 1) * width + j]; out[(height - i - 1) * width + j] = t; }
        }
      }
      result = corner_count;
    }
  }
  if (result <= 0) {
    const int area = w.pw * w.ph;
    corner_count = corner_count < area ? corner_count : area;
    for (int i = 0; i < corner_count; i++) out[i] = corners[i];
    result = -corner_count;
    if (result == -area) result = -result;
  }
  return result;
}

CB_HD inline bool check_board_monotony(const Ptf* c, int pw, int ph) {
  for (int k = 0; k < 2; ++k) {
    const int max_i = (k == 0 ? ph : pw);
    const int max_j = (k == 0 ? pw : ph) - 1;
    for (int i = 0; i < max_i; ++i) {
      const Ptf a = k == 0 ? c[i * pw] : c[i];
      const Ptf b = k == 0 ? c[(i + 1) * pw - 1] : c[(ph - 1) * pw + i];
      const float dx0 = b.x - a.x, dy0 = b.y - a.y;
      if (fabsf(dx0) + fabsf(dy0) < FLT_EPSILON) return false;
      float prevt = 0;
      for (int j = 1; j < max_j; ++j) {
        const Ptf cc = k == 0 ? c[i * pw + j] : c[j * pw + i];
        const float t = ((cc.x - a.x) * dx0 + (cc.y - a.y) * dy0) / (dx0 * dx0 + dy0 * dy0);
        if (t < prevt || t > 1) return false;
        prevt = t;
      }
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------ processQuads
// quads already in w (n_quads, max_quads set).  On success writes pw*ph corners to out_pts and returns true.
// (find_quad_neighbors must have run)
CB_HD inline bool process_groups(Work& w, Ptf* out_pts, int* n_out_pts, int* prev_sqr_size, int* scratch) {
  *n_out_pts = 0;
  if (w.n_quads <= 0) return false;
  w.scan_pos = 0;
  for (int group_idx = 0;; group_idx++) {
    find_connected_quads(w, group_idx);
    if (w.n_group == 0) break;
#ifdef CB_DEBUG
    const int dbg_g = w.n_group;
#endif
    int count = order_found_connected_quads(w);
#ifdef CB_DEBUG
    fprintf(stderr, "  group %d: size %d -> ordered %d", group_idx, dbg_g, count);
#endif
    if (count == 0) {
#ifdef CB_DEBUG
      fprintf(stderr, "\n");
#endif
      continue;
    }
    count = clean_found_connected_quads(w);
#ifdef CB_DEBUG
    fprintf(stderr, " -> cleaned %d", count);
#endif
    count = check_quad_group(w, scratch);
#ifdef CB_DEBUG
    fprintf(stderr, " -> check %d\n", count);
#endif
    int n = count > 0 ? w.pw * w.ph : -count;
    n = n < w.pw * w.ph ? n : w.pw * w.ph;
    float sum_dist = 0;
    int total = 0;
    for (int i = 0; i < n; i++) {
      const Corner& c = w.corners[w.corner_group[i]];
      float sum = 0;
      int ni = 0;
      for (int k = 0; k < 4; ++k) if (c.nb[k] >= 0) { const Ptf o = w.corners[c.nb[k]].pt; sum += sqrtf(normL2Sqrf(o.x - c.pt.x, o.y - c.pt.y)); ni++; }
      sum_dist += sum;
      total += ni;
    }
    *prev_sqr_size = (int)lrint((double)(sum_dist / (float)(total > 1 ? total : 1)));
    if (count > 0 || (-count > *n_out_pts)) {
      for (int i = 0; i < n; ++i) out_pts[i] = w.corners[w.corner_group[i]].pt;
      *n_out_pts = n;
      if (count == w.pw * w.ph && check_board_monotony(out_pts, w.pw, w.ph)) return true;
    }
  }
  return false;
}

CB_HD inline bool process_quads(Work& w, Ptf* out_pts, int* n_out_pts, int* prev_sqr_size, int* scratch) {
  *n_out_pts = 0;
  if (w.n_quads <= 0) return false;
  find_quad_neighbors(w, SerialPar());
  return process_groups(w, out_pts, n_out_pts, prev_sqr_size, scratch);
}

}  // namespace cb
