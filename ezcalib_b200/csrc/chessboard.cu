// chessboard.cu -- batched chessboard detection on the GPU (half (a), chessboard path).
//
// Replaces ChessboardCalibDetector::detectTarget (/root/reference/include/target_detectors/chessboard_detector.hpp:27-50):
//   cv::findChessboardCorners(img, Size(nb_cols-1, nb_rows-1), corners,
//                             ADAPTIVE_THRESH + NORMALIZE_IMAGE + FILTER_QUADS + FAST_CHECK)
//   + cv::cornerSubPix(img, corners, (5,5), (-1,-1), {COUNT+EPS, 100, 0.05}).
// The algorithm lives in OpenCV (un-vendored; "OpenCV 3 or 4", README.md:67); it is restated against the OpenCV 4.13
// behaviour observed in this image (cv2 4.13.0), see cb_logic.h and tests/test_cb_logic.py.
//
// Structure: the image-level passes run as data-parallel kernels over the whole batch -- histogram binarisation,
// 3x3 dilation, white frame, run-based union-find labelling (black 4-connected = the "holes"
// cv::findContours(RETR_CCOMP) reports, white 8-connected = their parents), Suzuki border following
// (CHAIN_APPROX_SIMPLE) per hole, approxPolyDP/quad filtering per hole (one thread for short borders, one warp with the
// contour in shared memory for long ones) -- and the quad-graph logic (neighbour linking on a shared-memory grid,
// grouping, ordering, corner extraction) runs with one block per image on the rules of cb_logic.h.  The attempt loop
// of cv::findChessboardCorners (the histogram-binarised image undilated and then with 1..7 dilations, then 6 adaptive
// thresholds x 8 dilation levels on the equalised image) is driven by the host for all images of the batch in
// lock-step; images that are done drop out.  CALIB_CB_FAST_CHECK (checkChessboardBinary + checkChessboard) is
// evaluated first when the caller asks for it (fast_check_gate below).
#include <algorithm>
#include <chrono>
#include <vector>

#include "cb_logic.h"
#include <cstdlib>

#include "cb_rows.cuh"
#include "images.cuh"
#include "scan_sort.cuh"

namespace ezc {

// ---------------------------------------------------------------------------------------------- histogram binarisation
// 16-pixel chunks, one shared-memory histogram per warp (a chessboard image is bimodal: a single histogram serialises on
// two bins), equal neighbouring pixels counted once
__global__ void cb_hist_kernel(const uint8_t* __restrict__ img, long long pitch, long long istride, bool vec, int W, int H, int* __restrict__ hist) {
  __shared__ int sh[8][256];
  const int im = blockIdx.y;
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  int* mine = sh[threadIdx.x >> 5];
  const uint8_t* I = img + (long long)im * istride;
  const int cpr = (W + 15) >> 4, n_chunks = cpr * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_chunks; idx += gridDim.x * blockDim.x) {
    const int y = idx / cpr, x0 = (idx - y * cpr) << 4;
    const Chunk16 c = ld16(I + (long long)y * pitch, x0, W, vec, 0u);
    const int n = W - x0 < 16 ? W - x0 : 16;
    int prev = -1, cnt = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < n) {
        const int v = (c.w[k >> 2] >> (8 * (k & 3))) & 0xff;
        if (v == prev) ++cnt;
        else { if (cnt) atomicAdd(&mine[prev], cnt); prev = v; cnt = 1; }
      }
    }
    if (cnt) atomicAdd(&mine[prev], cnt);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][i];
    if (t) atomicAdd(&hist[im * 256 + i], t);
  }
}

// icvBinarizationHistogramBased (calibinit.cpp): threshold from the smoothed histogram's maxima; one thread per image
__global__ void cb_select_threshold_kernel(const int* __restrict__ hist_all, int n_img, int W, int H, int* __restrict__ thresh) {
  const int im = blockIdx.x * blockDim.x + threadIdx.x;
  if (im >= n_img) return;
  const int* hist = hist_all + im * 256;
  const int max_pix = W * H, max_pix1 = max_pix / 100;
  int smooth[256], grad[256], maxpos[20];
  for (int i = 0; i < 256; ++i) {
    const int lo = i > 0 ? i - 1 : 0, hi = i < 255 ? i + 1 : 255;
    int s = 0;
    for (int k = lo; k <= hi; ++k) s += hist[k];
    smooth[i] = s / 3;
  }
  grad[0] = 0; grad[255] = 0;
  int prev = 0;
  for (int i = 1; i < 255; ++i) {
    int g = smooth[i - 1] - smooth[i + 1];
    if (abs(g) < 100) g = (prev == 0) ? -100 : prev;
    grad[i] = g;
    prev = g;
  }
  int n = 0;
  for (int i = 254; i > 2 && n < 20; --i) {
    if (grad[i - 1] < 0 && grad[i] > 0) {
      const int s = smooth[i - 1] + smooth[i] + smooth[i + 1];
      if (!(s < max_pix1 && i < 64)) maxpos[n++] = i;
    }
  }
  int t = 0;
  if (n == 0) {
    int acc = 0;
    for (int i = 0; i < 256; ++i) { acc += hist[i]; if (acc > max_pix / 2) { t = i; break; } }
  } else if (n == 1) {
    t = maxpos[0] / 2;
  } else if (n == 2) {
    t = (maxpos[0] + maxpos[1]) / 2;
  } else {
    int idx_acc = 0, acc = 0;
    for (int i = 255; i > 0; --i) { acc += hist[i]; if (acc > max_pix / 18) { idx_acc = i; break; } }
    int idx_bg = 0, bright = maxpos[0];
    for (int k = 0; k < n - 1; ++k) {
      idx_bg = k + 1;
      if (maxpos[k] < idx_acc) break;
      bright = maxpos[k];
    }
    int maxval = hist[maxpos[idx_bg]];
    if (maxpos[idx_bg] >= 250 && idx_bg + 1 < n) { idx_bg++; maxval = hist[maxpos[idx_bg]]; }
    for (int k = idx_bg + 1; k < n; ++k) if (hist[maxpos[k]] >= maxval) { maxval = hist[maxpos[k]]; idx_bg = k; }
    const int dist2 = (bright - maxpos[idx_bg]) / 2;
    t = bright - dist2;
  }
  thresh[im] = t;
}

// ---------------------------------------------------------------------------------------------- labelling
// One union-find over all pixels: black pixels connect 4-way (holes of cv::findContours), white pixels 8-way.
// label = smallest linear index of the component = its raster-first pixel (where the border following starts).
__device__ __forceinline__ int cbl_find(const int* L, int x) {
  int p = L[x];
  while (p != x) { x = p; p = L[x]; }
  return x;
}
__device__ __forceinline__ void cbl_union(int* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = cbl_find(L, a);
    b = cbl_find(L, b);
    if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
    else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
    else done = true;
  }
}
// Labelling in three steps that keep the number of atomic unions at O(#runs) instead of O(#pixels) (the per-pixel
// version spent 93 % of the detector's time contending for the root of the white background):
//   1. cb_label_runs: one warp per image row; every pixel is labelled with the first pixel of its horizontal run
//      (ballot + clz over 32-pixel chunks, run start carried between chunks).
//   2. cb_label_merge: unions between runs of adjacent rows, issued only where a new adjacency starts (the pixel is the
//      first of its run in either row; white diagonals only across a black up-neighbour).
//   3. cb_label_compress + cb_label_flatten: run heads are pointed at their roots first, pixels then need two hops.
__global__ void cb_label_runs_kernel(const uint8_t* __restrict__ bin, int* __restrict__ L, int W, int H, int n_img,
                                     const int* __restrict__ active) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)H * n_img) return;
  const int slot = (int)(gw / H);
  const int y = (int)(gw - (long long)slot * H);
  const int im = active ? active[slot] : slot;
  const long long base = ((long long)im * H + y) * W;
  int carry = 0;
  int prev_last = 3;
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const bool in = x < W;
    const int v = in ? (bin[base + x] != 0 ? 1 : 0) : 2;
    int left = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) left = prev_last;
    const bool start = in && (v != left);
    const unsigned m = __ballot_sync(0xffffffffu, start);
    const unsigned le = m & (0xffffffffu >> (31 - lane));
    const int rs = le ? (x0 + 31 - __clz(le)) : carry;
    if (in) L[base + x] = (int)(base + rs);
    carry = __shfl_sync(0xffffffffu, rs, 31);
    prev_last = __shfl_sync(0xffffffffu, v, 31);
  }
}
// ---- the same steps on 16-pixel row chunks (cb_rows.cuh): run logic on 16-bit masks, labels touched only at run heads.
// After the merge step a pixel is the root of its component iff L[p] == p, a run head points at its root once
// cb_label_heads has run, and every other pixel points at its run head -- so "root of pixel p" is L[L[p]] and no pass
// ever rewrites the 4-byte labels of all pixels.  Holes (black roots) are listed in raster order through per-block
// counts: heads counts them, a scan over the blocks (not the pixels) gives each block its offset, list writes them.
__device__ __forceinline__ void row_masks(const uint8_t* __restrict__ row, int x0, int W, bool vec, uint32_t& v, uint32_t& left, uint32_t& right) {
  v = nz16(ld16(row, x0, W, vec, 0u));
  left = x0 > 0 ? (row[x0 - 1] != 0 ? 1u : 0u) : 0u;
  right = x0 + 16 < W ? (row[x0 + 16] != 0 ? 1u : 0u) : 0u;
}
__global__ void cb_label_merge16_kernel(const uint8_t* __restrict__ bin, int* __restrict__ L, int W, int H, const int* __restrict__ active, bool vec) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok || c.y == 0) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  const long long base = ((long long)im * H + c.y) * W;
  uint32_t v, vl, vr, u, ul, ur;
  row_masks(bin + base, c.x0, W, vec, v, vl, vr);
  row_masks(bin + base - W, c.x0, W, vec, u, ul, ur);
  const uint32_t valid = valid16(c.x0, W);
  const uint32_t first = c.x0 == 0 ? 1u : 0u;
  const uint32_t vprev = ((v << 1) | vl) & 0xffffu, uprev = ((u << 1) | ul) & 0xffffu, unext = (u >> 1) | (ur << 15);
  const uint32_t start_here = (v ^ vprev) | first, start_up = (u ^ uprev) | first;
  uint32_t same = ~(v ^ u) & (start_here | start_up) & valid;   // a new vertical adjacency between runs of one colour starts here
  const uint32_t wb = v & ~u & valid;                            // white under black: the upper diagonals are separate runs
  uint32_t dr = wb & unext, dl = wb & uprev & ~first;
  const int g0 = (int)(base + c.x0);
  while (same) { const int k = __ffs(same) - 1; same &= same - 1; cbl_union(L, g0 + k, g0 + k - W); }
  while (dr) { const int k = __ffs(dr) - 1; dr &= dr - 1; cbl_union(L, g0 + k, g0 + k - W + 1); }
  while (dl) { const int k = __ffs(dl) - 1; dl &= dl - 1; cbl_union(L, g0 + k, g0 + k - W - 1); }
}
// run heads -> roots (pointer jumping; concurrent writes only ever shorten a path to a valid ancestor); counts the roots
// of the wanted colour per block
__global__ void cb_label_heads_kernel(const uint8_t* __restrict__ bin, int* __restrict__ L, int W, int H, const int* __restrict__ active, bool vec,
                                      int white_roots, int* __restrict__ block_count) {
  const ChunkPos c = chunk_pos(W, H);
  int cnt = 0;
  if (c.ok) {
    const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
    const long long base = ((long long)im * H + c.y) * W;
    uint32_t v, vl, vr;
    row_masks(bin + base, c.x0, W, vec, v, vl, vr);
    uint32_t heads = ((v ^ (((v << 1) | vl) & 0xffffu)) | (c.x0 == 0 ? 1u : 0u)) & valid16(c.x0, W);
    const uint32_t want = white_roots ? v : ~v;
    const int g0 = (int)(base + c.x0);
    while (heads) {
      const int k = __ffs(heads) - 1; heads &= heads - 1;
      const int r = cbl_find(L, g0 + k);
      L[g0 + k] = r;
      cnt += (r == g0 + k && ((want >> k) & 1u)) ? 1 : 0;
    }
  }
  __shared__ int wsum[8];
  for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    block_count[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}
// block_off = exclusive scan of block_count; roots of the wanted colour in raster order -> hole_start
__global__ void cb_root_list_kernel(const uint8_t* __restrict__ bin, const int* __restrict__ L, int W, int H, const int* __restrict__ active, bool vec,
                                    int white_roots, const int* __restrict__ block_count, const int* __restrict__ block_off,
                                    int* __restrict__ hole_start) {
  const long long bi = (long long)blockIdx.y * gridDim.x + blockIdx.x;
  if (block_count[bi] == 0) return;
  const ChunkPos c = chunk_pos(W, H);
  uint32_t roots = 0;
  int g0 = 0;
  if (c.ok) {
    const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
    const long long base = ((long long)im * H + c.y) * W;
    uint32_t v, vl, vr;
    row_masks(bin + base, c.x0, W, vec, v, vl, vr);
    uint32_t heads = ((v ^ (((v << 1) | vl) & 0xffffu)) | (c.x0 == 0 ? 1u : 0u)) & valid16(c.x0, W) & (white_roots ? v : ~v);
    g0 = (int)(base + c.x0);
    while (heads) {
      const int k = __ffs(heads) - 1; heads &= heads - 1;
      if (L[g0 + k] == g0 + k) roots |= 1u << k;
    }
  }
  // exclusive scan of the per-thread root counts over the block (threads are in raster order)
  const int mine = __popc(roots);
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  __shared__ int wtot[8];
  if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
  __syncthreads();
  int pos = block_off[bi] + incl - mine;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pos += wtot[w];
  while (roots) { const int k = __ffs(roots) - 1; roots &= roots - 1; hole_start[pos++] = g0 + k; }
}

// ---------------------------------------------------------------------------------------------- border following
// icvFetchContour for a hole border with CHAIN_APPROX_SIMPLE (contours.cpp): starts at the foreground pixel left of the
// hole's raster-first pixel, emits a vertex whenever the chain direction changes.
struct HoleInfo {
  int start;       // global linear index of the hole's first (black) pixel
  int n_pts;       // vertices of the simple-approximated border (-1: too long)
  int minx, miny, maxx, maxy;
  int parent;      // label of the enclosing white component
  int offset;      // into the point pool
  int quad_ok;
  cb::Pt quad[4];
};

// chain-code direction s -> (dx, dy) = ({1,1,0,-1,-1,-1,0,1}, {0,-1,-1,-1,0,1,1,1}) from 2-bit packed tables
__device__ __forceinline__ int dir_dx(int s) { return (int)((0x901Au >> (2 * s)) & 3u) - 1; }
__device__ __forceinline__ int dir_dy(int s) { return (int)((0xA901u >> (2 * s)) & 3u) - 1; }
template <bool STORE>
__global__ void cb_trace_kernel(const uint8_t* __restrict__ bin, const int* __restrict__ L, int W, int H, int n_holes,
                                HoleInfo* __restrict__ holes, cb::Pt* __restrict__ pool, int max_pts) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes) return;
  HoleInfo& hi = holes[h];
  if (STORE && (hi.n_pts <= 0 || hi.offset < 0)) return;
  const long long N = (long long)W * H;
  const int im = (int)(hi.start / N);
  const int p0 = (int)(hi.start - (long long)im * N);
  const uint8_t* B = bin + (long long)im * N;
  const int y0 = p0 / W, x0 = p0 % W - 1;  // origin: the white pixel left of the first hole pixel (frame => x0 >= 2)
  const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1}, DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  auto val = [&](int x, int y) -> bool { return B[(long long)y * W + x] != 0; };
  int n = 0, minx = x0, maxx = x0, miny = y0, maxy = y0;
  cb::Pt* out = STORE ? pool + hi.offset : nullptr;
  int s = 0, s_end = 0;
  do { s = (s - 1) & 7; } while (!val(x0 + dir_dx(s), y0 + dir_dy(s)) && s != s_end);
  if (s == s_end && !val(x0 + dir_dx(s), y0 + dir_dy(s))) {
    if (STORE) out[0] = cb::Pt{x0, y0};
    n = 1;
  } else {
    const int i1x = x0 + dir_dx(s), i1y = y0 + dir_dy(s);
    int i3x = x0, i3y = y0, prev_s = s ^ 4;
    int x = x0, y = y0;
    for (;;) {
      // the 8 neighbours of the current pixel with independent loads (the probe loop of icvFetchContour would issue them
      // one dependent load after the other), then the first foreground neighbour in rotation order from the mask
      unsigned m8 = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) m8 |= (val(i3x + DX[k], i3y + DY[k]) ? 1u : 0u) << k;
      const int s1 = (s + 1) & 7;
      const unsigned rot = ((m8 | (m8 << 8)) >> s1) & 0xffu;
      s = s1 + (__ffs(rot) - 1);   // rot != 0: the pixel we came from is foreground
      s &= 7;
      const int sdx = dir_dx(s), sdy = dir_dy(s);   // arithmetic table: an indexed local array costs a local-memory load per step
      const int i4x = i3x + sdx, i4y = i3y + sdy;
      if (s != prev_s) {
        if (STORE) { if (n < hi.n_pts) out[n] = cb::Pt{x, y}; }
        ++n;
        minx = x < minx ? x : minx; maxx = x > maxx ? x : maxx; miny = y < miny ? y : miny; maxy = y > maxy ? y : maxy;
        prev_s = s;
      }
      x += sdx; y += sdy;
      if (i4x == x0 && i4y == y0 && i3x == i1x && i3y == i1y) break;
      i3x = i4x; i3y = i4y;
      s = (s + 4) & 7;
      if (!STORE && n > max_pts) { n = -1; break; }
    }
  }
  if (!STORE) {
    hi.n_pts = n;
    hi.minx = minx; hi.maxx = maxx; hi.miny = miny; hi.maxy = maxy;
    hi.parent = L[L[(long long)im * N + (long long)y0 * W + x0]];   // pixel -> run head -> root
    hi.quad_ok = 0;
    // only contours that can become a quad get storage: bounding rect area >= 25 and a sane vertex count
    const bool want = n >= 4 && (maxx - minx + 1) * (maxy - miny + 1) >= 25;
    hi.offset = want ? 0 : -1;
  }
}
__global__ void cb_hole_init_kernel(const int* __restrict__ hole_start, int n_holes, HoleInfo* __restrict__ holes, int* __restrict__ need) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes) return;
  holes[h].start = hole_start[h];
  holes[h].n_pts = 0; holes[h].offset = -1; holes[h].quad_ok = 0;
  need[h] = 0;
}
__global__ void cb_hole_need_kernel(const HoleInfo* __restrict__ holes, int n_holes, int* __restrict__ need) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes) return;
  need[h] = (holes[h].offset >= 0 && holes[h].n_pts > 0) ? 3 * holes[h].n_pts + 8 : 0;  // contour + 2 approx buffers
}
constexpr int kWarpContour = 96;    // up to here: one thread per hole (cb_quad_kernel)
constexpr int kWarpContourMid = 512;  // up to here: a warp with 512-vertex buffers, beyond: kMaxContourPts-vertex buffers
// pool offsets; holes with long contours go to the work lists of the warp kernels (counts[0] / counts[1])
__global__ void cb_hole_offset_kernel(HoleInfo* __restrict__ holes, int n_holes, const int* __restrict__ need_scan, int* __restrict__ list_mid,
                                      int* __restrict__ list_big, int* __restrict__ counts) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes) return;
  if (holes[h].offset < 0) return;
  holes[h].offset = need_scan[h];
  const int n = holes[h].n_pts;
  if (n > kWarpContourMid) list_big[atomicAdd(&counts[1], 1)] = h;
  else if (n > kWarpContour) list_mid[atomicAdd(&counts[0], 1)] = h;
}
// approxPolyDP levels 1..7 + convexity + FILTER_QUADS tests: one thread per hole (contours up to kWarpContour vertices)
__global__ void cb_quad_kernel(HoleInfo* __restrict__ holes, int n_holes, cb::Pt* __restrict__ pool, int* __restrict__ stack_pool) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes) return;
  HoleInfo& hi = holes[h];
  if (hi.offset < 0 || hi.n_pts <= 0 || hi.n_pts > kWarpContour) return;
  cb::Pt* c = pool + hi.offset;
  cb::Pt* ta = c + hi.n_pts + 2;
  cb::Pt* tb = ta + hi.n_pts + 2;
  int* stack = stack_pool + (size_t)hi.offset * 2;  // >= 4*n + 16 ints per hole
  cb::Pt q[4];
  if (cb::contour_to_quad(c, hi.n_pts, ta, tb, stack, q)) {
    hi.quad_ok = 1;
    for (int i = 0; i < 4; ++i) hi.quad[i] = q[i];
  }
}

// Long contours (noise blobs of the adaptive-threshold attempts run to thousands of vertices): one WARP per hole.  The
// Douglas-Peucker scans -- "farthest point from the start" and "farthest point from the chord" -- are lane-strided with a
// (distance, first index) arg-max, i.e. exactly the element the sequential scan of cb::approx_poly_dp_closed keeps (it
// uses a strict '>' in contour order); the stack discipline and the clean-up pass are unchanged and run on every lane
// redundantly on lane-uniform data, lane 0 does the writes.
__device__ __forceinline__ void warp_argmax(double& d, int& ord) {
  for (int m = 16; m >= 1; m >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, m);
    const int oo = __shfl_xor_sync(0xffffffffu, ord, m);
    if (od > d || (od == d && oo < ord)) { d = od; ord = oo; }
  }
}
template <class StackT>
__device__ int approx_poly_dp_closed_warp(const cb::Pt* src, int count, cb::Pt* dst, double eps, StackT* stack_buf) {
  using cb::Pt;
  if (count == 0) return 0;
  const int lane = threadIdx.x & 31;
  int top = 0;
  // every lane performs the (identical) stack and output writes itself, so it reads back its own stores: no warp sync
  auto push = [&](int s_, int e_) { stack_buf[2 * top] = (StackT)s_; stack_buf[2 * top + 1] = (StackT)e_; ++top; };
  const int init_iters = 3;
  int slice_s = 0, slice_e = 0, right_s = 0, right_e = 0;
  Pt start_pt = {-1000000, -1000000}, end_pt = {0, 0}, pt = {0, 0};
  int pos = 0, new_count = 0;
  bool le_eps = false;
  eps *= eps;
  right_s = 0;
  for (int i = 0; i < init_iters; i++) {
    pos = (pos + right_s) % count;
    start_pt = src[pos];
    const int base = pos + 1;  // element j (1-based) is src[(base + j - 1) % count]
    double max_dist = 0;
    int best = 0x7fffffff;
    for (int j = 1 + lane; j < count; j += 32) {
      int pp = base + j - 1; if (pp >= count) pp -= count; if (pp >= count) pp -= count;
      pt = src[pp];
      const double dx = pt.x - start_pt.x, dy = pt.y - start_pt.y;
      const double dist = dx * dx + dy * dy;
      if (dist > max_dist) { max_dist = dist; best = j; }
    }
    warp_argmax(max_dist, best);
    if (max_dist > 0) right_s = best;   // sequential scan only assigns on dist > max_dist (initially 0)
    pos = (base + count - 1) % count;   // after the scan the cursor has advanced count times in total
    le_eps = max_dist <= eps;
  }
  if (!le_eps) {
    right_e = slice_s = pos % count;
    slice_e = right_s = (right_s + slice_s) % count;
    push(right_s, right_e);
    push(slice_s, slice_e);
  } else {
    dst[new_count] = start_pt;
    new_count++;
  }
  while (top > 0) {
    --top;
    slice_s = stack_buf[2 * top]; slice_e = stack_buf[2 * top + 1];
    end_pt = src[slice_e];
    pos = slice_s;
    start_pt = src[pos]; if (++pos >= count) pos = 0;
    if (pos != slice_e) {
      const int len = (slice_e - pos + count) % count;  // points pos .. slice_e-1 (cyclic)
      double max_dist = 0;
      int best = 0x7fffffff;
      const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
      const double L2 = dx * dx + dy * dy;
      // short ranges (the bulk on ragged contours) are scanned by every lane redundantly: no shuffles at all
      const bool small = len <= 4;
      for (int o = small ? 0 : lane; o < len; o += small ? 1 : 32) {
        int pp = pos + o; if (pp >= count) pp -= count;
        pt = src[pp];
        const double px = pt.x - start_pt.x, py = pt.y - start_pt.y;
        const double t = px * dx + py * dy;
        double dist;
        if (t < 0) dist = px * px + py * py;
        else if (t > L2) { const double ex = pt.x - end_pt.x, ey = pt.y - end_pt.y; dist = ex * ex + ey * ey; }
        else { const double cr = py * dx - px * dy; dist = cr * cr / L2; }
        if (dist > max_dist) { max_dist = dist; best = o; }
      }
      if (!small) warp_argmax(max_dist, best);
      if (max_dist > 0) { int pp = pos + best; if (pp >= count) pp -= count; right_s = pp; }
      le_eps = max_dist <= eps;
    } else {
      le_eps = true;
      start_pt = src[slice_s];
    }
    if (le_eps) {
      dst[new_count] = start_pt;
      new_count++;
    } else {
      right_e = slice_e;
      slice_e = right_s;
      push(right_s, right_e);
      push(slice_s, slice_e);
    }
  }
  __syncwarp();
  // last stage: remove extra points on [almost] straight lines (sequential; lane 0 writes, every lane tracks the state)
  count = new_count;
  int result = new_count;
  if (lane == 0) {
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
    result = new_count;
  }
  result = __shfl_sync(0xffffffffu, result, 0);
  __syncwarp();
  return result;
}
// One warp (= one block) per long contour; the two approximation buffers and the subdivision stack live in shared memory:
// Douglas-Peucker on a ragged contour is a chain of thousands of tiny dependent steps, each a couple of loads.
// Two capacity classes (shared memory per block decides how many contours an SM works on at once): CAP = 512 holds the
// bulk, CAP = kMaxContourPts the few giants; LO is the exclusive lower bound of the class.
template <int CAP> constexpr size_t quad_warp_smem() { return 2 * sizeof(cb::Pt) * (CAP + 4) + sizeof(unsigned short) * (4 * CAP + 16); }
template <int CAP>
__device__ void quad_warp_one(HoleInfo& hi, const cb::Pt* __restrict__ pool, unsigned char* qw_smem) {
  const int lane = threadIdx.x;
  const int n = hi.n_pts;
  const cb::Pt* c = pool + hi.offset;
  cb::Pt* ta = reinterpret_cast<cb::Pt*>(qw_smem);
  cb::Pt* tb = ta + CAP + 4;
  unsigned short* stack = reinterpret_cast<unsigned short*>(tb + CAP + 4);
  // cb::contour_to_quad with the warp-parallel approximation; the bounding-box test already passed in cb_trace_kernel
  for (int i = lane; i < n; i += 32) ta[i] = c[i];
  __syncwarp();
  // same schedule as cb::contour_to_quad: first call on the previous level's first output, second call on its own output
  cb::Pt* X = ta;   // input of this level's first call
  cb::Pt* Y = tb;
  int nx = n;
  const cb::Pt* fin = nullptr;
  for (int level = 1; level <= cb::kMaxApproxLevel && !fin; level++) {
    const int nf = approx_poly_dp_closed_warp(X, nx, Y, (double)(float)level, stack);
    __syncwarp();
    if (nf == 4) { fin = Y; break; }
    const int ns = approx_poly_dp_closed_warp(Y, nf, X, (double)(float)level, stack);
    __syncwarp();
    if (ns == 4) { fin = X; break; }
    cb::Pt* t = X; X = Y; Y = t;
    nx = nf;
  }
  if (!fin) return;
  if (lane == 0) {
    cb::Pt q[4];
    if (cb::quad_from_4(fin, q)) {
      hi.quad_ok = 1;
      for (int i = 0; i < 4; ++i) hi.quad[i] = q[i];
    }
  }
}
// The holes with more than kWarpContour vertices are few among possibly 10^5 holes of a noisy threshold image: they are
// collected in two work lists (cb_hole_offset_kernel) and a fixed grid of warps walks each list.
template <int CAP>
__global__ void __launch_bounds__(32) cb_quad_warp_kernel(HoleInfo* __restrict__ holes, const int* __restrict__ list, const int* __restrict__ count,
                                                          const cb::Pt* __restrict__ pool) {
  extern __shared__ __align__(16) unsigned char qw_smem[];
  const int n = *count;
  for (int it = blockIdx.x; it < n; it += gridDim.x) {
    quad_warp_one<CAP>(holes[list[it]], pool, qw_smem);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- CALIB_CB_FAST_CHECK
// checkChessboardBinary + checkChessboard (OpenCV calib3d/src/checkchessboard.cpp), the gate the reference switches on
// (chessboard_detector.hpp:33): blobs of the eroded white / dilated black image whose min-area rectangle has a long side
// >= 10 px and aspect within [0.3, 3] are "quad hypotheses"; a board is plausible if some run of similar sizes (ratio
// <= 1.4) holds more than w*h/2 of them with >= 75 % of the expected black and white squares.  Here: 8-connected
// components of the foreground by the run-based labelling above, the run end points of every component that survive an Akl-Toussaint
// filter give its convex hull (no border walk), min-area rectangle over the hull edges; the (size, class) lists go to the host, which runs checkQuads.
struct FcComp {
  int start;        // raster-first pixel of the component (global linear index)
  int n_cand;       // run end points of the component (upper bound of what is stored)
  int fill;         // points stored = candidates strictly outside the extreme-point polygon
  int offset;       // into the point pool, -1 = skipped
  int ex[8], ey[8]; // extreme points in the 8 principal directions, in angular order (left, top-left, top, ...)
  unsigned long long key[8];  // atomicMax keys: (direction value + bias) << 32 | pixel
};
// The convex hull of a component is the hull of the end points of its horizontal runs, so nothing has to be traced:
// three pixel-parallel passes (extreme points by atomicMax, survivor count, survivor scatter) replace a sequential
// border walk whose length for the white background is the whole image frame.
__global__ void cb_fc_init_kernel(int n_comp, FcComp* __restrict__ comps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_comp) return;
  FcComp& ci = comps[c];
  ci.start = 0; ci.n_cand = 0; ci.fill = 0; ci.offset = -1;
  for (int k = 0; k < 8; ++k) ci.key[k] = 0ull;
}
// Run heads of the foreground are pointed at their roots and every root draws a component id from a counter -- the gate
// does not care about the order of the components, so no prefix sum is needed.  cidmap is indexed by the compact pixel
// index (slot * N + pixel) and only written / read at roots.
__global__ void cb_fc_roots16_kernel(const uint8_t* __restrict__ fg, int* __restrict__ L, int W, int H, const int* __restrict__ active, bool vec,
                                     int* __restrict__ cidmap, int* __restrict__ counter) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  const long long N = (long long)W * H, base = (long long)im * N + (long long)c.y * W;
  uint32_t v, vl, vr;
  row_masks(fg + base, c.x0, W, vec, v, vl, vr);
  uint32_t heads = v & ~((v << 1) | vl);   // first pixel of a foreground run (x == 0: vl = 0)
  const int g0 = (int)(base + c.x0);
  while (heads) {
    const int k = __ffs(heads) - 1; heads &= heads - 1;
    const int r = cbl_find(L, g0 + k);
    L[g0 + k] = r;
    if (r == g0 + k) cidmap[(long long)blockIdx.y * N + (r - (long long)im * N)] = atomicAdd(counter, 1);
  }
}
// The run end points of the foreground (first or last pixel of a horizontal run) are the hull candidates.
// MODE 0: extreme points + candidate count; MODE 1: store the candidates strictly outside the polygon of the extreme points
template <int MODE>
__global__ void cb_fc_points16_kernel(const uint8_t* __restrict__ fg, const int* __restrict__ L, const int* __restrict__ cidmap, int W, int H,
                                      const int* __restrict__ active, bool vec, FcComp* __restrict__ comps, cb::Pt* __restrict__ pool) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  const long long N = (long long)W * H, ibase = (long long)im * N, base = ibase + (long long)c.y * W;
  uint32_t v, vl, vr;
  row_masks(fg + base, c.x0, W, vec, v, vl, vr);
  const uint32_t vprev = ((v << 1) | vl) & 0xffffu, vnext = (v >> 1) | (vr << 15);
  uint32_t cand = v & (~vprev | ~vnext) & 0xffffu;   // x == 0 / x == W-1: the missing neighbour reads as background
  const int g0 = (int)(base + c.x0);
  const int y = c.y;
  int last_head = -1, cid = -1, pending = 0;
  while (cand) {
    const int k = __ffs(cand) - 1; cand &= cand - 1;
    const int x = c.x0 + k, g = g0 + k;
    const int head = L[g];                 // pixel -> run head (a run head that is a root points at itself)
    if (head != last_head) {
      const int root = L[head];
      const int ncid = cidmap[(long long)blockIdx.y * N + (root - ibase)];
      if (MODE == 0 && pending && ncid != cid) { atomicAdd(&comps[cid].n_cand, pending); pending = 0; }
      cid = ncid; last_head = head;
      if (MODE == 0 && root == g) comps[cid].start = root;   // the first pixel of a component starts a run
    }
    FcComp& ci = comps[cid];
    if (MODE == 0) {
      ++pending;
      const int p = (int)(g - ibase);
      const int kv[8] = {-x, -x - y, -y, x - y, x, x + y, y, y - x};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned long long key = ((unsigned long long)(unsigned)(kv[q] + (1 << 20)) << 32) | (unsigned)p;
        if (key > ci.key[q]) atomicMax(&ci.key[q], key);
      }
    } else {
      if (ci.offset == -1) continue;
      bool outside = false;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int q2 = (q + 1) & 7;
        const long long ax = ci.ex[q2] - ci.ex[q], ay = ci.ey[q2] - ci.ey[q];
        if (ax == 0 && ay == 0) continue;
        const long long cr = ax * (y - ci.ey[q]) - ay * (x - ci.ex[q]);
        if (cr < 0) { outside = true; break; }
      }
      if (outside) pool[ci.offset + atomicAdd(&ci.fill, 1)] = cb::Pt{x, y};
    }
  }
  if (MODE == 0 && pending) atomicAdd(&comps[cid].n_cand, pending);
}
// decode the extreme points; components whose bounding-box diagonal is below 10 px cannot pass the size test
__global__ void cb_fc_prefilter_kernel(FcComp* __restrict__ comps, int n_comp, int W) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_comp) return;
  FcComp& ci = comps[c];
  for (int k = 0; k < 8; ++k) {
    const int p = (int)(unsigned)(ci.key[k] & 0xffffffffull);
    ci.ey[k] = p / W; ci.ex[k] = p - ci.ey[k] * W;
  }
  const long long bw = ci.ex[4] - ci.ex[0], bh = ci.ey[6] - ci.ey[2];
  ci.offset = (bw * bw + bh * bh >= 100) ? 0 : -1;
}
__global__ void cb_fc_need_kernel(const FcComp* __restrict__ comps, int n, int* __restrict__ need) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  need[c] = comps[c].offset >= 0 ? 2 * (comps[c].n_cand + 8) + 16 : 0;  // stored points (<= candidates) + extreme points, and their hull
}
__global__ void cb_fc_offset_kernel(FcComp* __restrict__ comps, int n, const int* __restrict__ need_scan) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  if (comps[c].offset >= 0) comps[c].offset = need_scan[c];
}
struct FcHyp { float size; int slot_class; };  // slot_class = image * 2 + class
// convex hull (monotone chain on sorted points) + min-area rectangle over the hull edges; emits a hypothesis.
// Serial version on global memory: only for components whose candidate list does not fit the block version below.
__device__ void fc_minrect_serial(const FcComp& ci, cb::Pt* __restrict__ pool, double& bw_out, double& bh_out) {
  cb::Pt* a = pool + ci.offset;
  for (int k = 0; k < 8; ++k) a[ci.fill + k] = cb::Pt{ci.ex[k], ci.ey[k]};
  const int n = ci.fill + 8;
  cb::Pt* h = a + n + 4;
  auto less = [](const cb::Pt& p, const cb::Pt& q) { return p.x < q.x || (p.x == q.x && p.y < q.y); };
  // heap sort by (x, y)
  auto sift = [&](int start, int end) {
    int root = start;
    for (;;) {
      int child = 2 * root + 1;
      if (child > end) break;
      if (child + 1 <= end && less(a[child], a[child + 1])) ++child;
      if (less(a[root], a[child])) { const cb::Pt t = a[root]; a[root] = a[child]; a[child] = t; root = child; }
      else break;
    }
  };
  for (int st = (n - 2) / 2; st >= 0; --st) sift(st, n - 1);
  for (int end = n - 1; end > 0; --end) { const cb::Pt t = a[0]; a[0] = a[end]; a[end] = t; sift(0, end - 1); }
  int k = 0;
  auto cross = [](const cb::Pt& o, const cb::Pt& p, const cb::Pt& q) { return (long long)(p.x - o.x) * (q.y - o.y) - (long long)(p.y - o.y) * (q.x - o.x); };
  for (int i = 0; i < n; ++i) {
    if (i > 0 && a[i].x == a[i - 1].x && a[i].y == a[i - 1].y) continue;
    while (k >= 2 && cross(h[k - 2], h[k - 1], a[i]) <= 0) k--;
    h[k++] = a[i];
  }
  const int lower = k + 1;
  for (int i = n - 2; i >= 0; --i) {
    if (a[i].x == a[i + 1].x && a[i].y == a[i + 1].y) continue;
    while (k >= lower && cross(h[k - 2], h[k - 1], a[i]) <= 0) k--;
    h[k++] = a[i];
  }
  if (k > 1) k--;  // last point == first
  // width / height as cv::minAreaRect (4.13) reports them: "width" is the extent along the box direction whose angle lies
  // in [-90, 0) degrees in image coordinates (x >= 0, y < 0), "height" the other one (measured on cv2)
  double bw = 0, bh = 0;
  if (k == 2) {
    const double dx = h[1].x - h[0].x, dy = h[1].y - h[0].y;
    const double len = sqrt(dx * dx + dy * dy);
    const bool along = (dy < 0 && dx >= 0) || (dy > 0 && dx <= 0);
    bw = along ? len : 0; bh = along ? 0 : len;
  } else if (k >= 3) {
    double best = 1e300;
    for (int e = 0; e < k; ++e) {
      const cb::Pt p = h[e], q = h[(e + 1) % k];
      const double ux = q.x - p.x, uy = q.y - p.y;
      const double len = sqrt(ux * ux + uy * uy);
      const double cx = ux / len, cy = uy / len;
      double mn = 0, mx = 0, mh = 0;
      for (int i = 0; i < k; ++i) {
        const double vx = h[i].x - p.x, vy = h[i].y - p.y;
        const double t = vx * cx + vy * cy, d = fabs(vy * cx - vx * cy);
        mn = t < mn ? t : mn; mx = t > mx ? t : mx; mh = d > mh ? d : mh;
      }
      const double w = mx - mn, area = w * mh;
      if (area < best) {
        best = area;
        const bool along = (uy < 0 && ux >= 0) || (uy > 0 && ux <= 0);
        bw = along ? w : mh; bh = along ? mh : w;
      }
    }
  }
  bw_out = bw; bh_out = bh;
}
// width / height of the min-area rectangle over hull h[0..k) (packed x << 16 | y) for edge e, as the serial version computes them
__device__ __forceinline__ double fc_edge_rect(const uint32_t* __restrict__ h, int k, int e, double& bw, double& bh) {
  const double px = (double)(h[e] >> 16), py = (double)(h[e] & 0xffffu);
  const uint32_t qk = h[e + 1 == k ? 0 : e + 1];
  const double ux = (double)(qk >> 16) - px, uy = (double)(qk & 0xffffu) - py;
  const double len = sqrt(ux * ux + uy * uy);
  const double cx = ux / len, cy = uy / len;
  double mn = 0, mx = 0, mh = 0;
  for (int i = 0; i < k; ++i) {
    const double vx = (double)(h[i] >> 16) - px, vy = (double)(h[i] & 0xffffu) - py;
    const double t = vx * cx + vy * cy, d = fabs(vy * cx - vx * cy);
    mn = t < mn ? t : mn; mx = t > mx ? t : mx; mh = d > mh ? d : mh;
  }
  const double w = mx - mn;
  const bool along = (uy < 0 && ux >= 0) || (uy > 0 && ux <= 0);
  bw = along ? w : mh; bh = along ? mh : w;
  return w * mh;
}
// One block per component: the candidates are sorted (bitonic, keys x << 16 | y) and the hull is built in shared memory,
// the rectangles of the hull edges are evaluated in parallel; the first edge with the smallest area wins, as in the
// sequential scan.  (A single thread on global memory needed up to 1.2 ms for one large blob.)
constexpr int kFcSortCap = 4096;
constexpr int kFcThreads = 64;
__global__ void __launch_bounds__(kFcThreads) cb_fc_minrect_kernel(const FcComp* __restrict__ comps, int n_comp, cb::Pt* __restrict__ pool, long long N,
                                                                   int class_id, int small_coords, FcHyp* __restrict__ hyps, int* __restrict__ n_hyps,
                                                                   int max_hyps) {
  const int c = blockIdx.x;
  const FcComp& ci = comps[c];
  if (ci.offset < 0) return;
  __shared__ uint32_t key[kFcSortCap];
  __shared__ uint32_t hull[kFcSortCap + 1];
  __shared__ int s_k;
  __shared__ double s_area[kFcThreads], s_bw[kFcThreads], s_bh[kFcThreads];
  __shared__ int s_e[kFcThreads];
  const int tid = threadIdx.x;
  const int n = ci.fill + 8;
  double bw = 0, bh = 0;
  if (n > kFcSortCap || !small_coords) {
    if (tid != 0) return;
    fc_minrect_serial(ci, pool, bw, bh);
  } else {
    const cb::Pt* a = pool + ci.offset;
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = tid; i < m; i += kFcThreads) {
      uint32_t kk = 0xffffffffu;
      if (i < ci.fill) kk = ((uint32_t)a[i].x << 16) | (uint32_t)a[i].y;
      else if (i < n) kk = ((uint32_t)ci.ex[i - ci.fill] << 16) | (uint32_t)ci.ey[i - ci.fill];
      key[i] = kk;
    }
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (m >> 1); t += kFcThreads) {
          const int lo = ((t / stride) * stride << 1) + (t % stride), hi = lo + stride;
          const bool up = (lo & size) == 0;
          const uint32_t x = key[lo], y = key[hi];
          if ((x > y) == up) { key[lo] = y; key[hi] = x; }
        }
        __syncthreads();
      }
    if (tid == 0) {
      auto cross = [](uint32_t o, uint32_t p, uint32_t q) {
        return (long long)((int)(p >> 16) - (int)(o >> 16)) * ((int)(q & 0xffffu) - (int)(o & 0xffffu)) -
               (long long)((int)(p & 0xffffu) - (int)(o & 0xffffu)) * ((int)(q >> 16) - (int)(o >> 16));
      };
      int k = 0;
      for (int i = 0; i < n; ++i) {
        if (i > 0 && key[i] == key[i - 1]) continue;
        while (k >= 2 && cross(hull[k - 2], hull[k - 1], key[i]) <= 0) k--;
        hull[k++] = key[i];
      }
      const int lower = k + 1;
      for (int i = n - 2; i >= 0; --i) {
        if (key[i] == key[i + 1]) continue;
        while (k >= lower && cross(hull[k - 2], hull[k - 1], key[i]) <= 0) k--;
        hull[k++] = key[i];
      }
      if (k > 1) k--;  // last point == first
      s_k = k;
    }
    __syncthreads();
    const int k = s_k;
    double best = 1e300, bbw = 0, bbh = 0;
    int best_e = 0x7fffffff;
    if (k >= 3)
      for (int e = tid; e < k; e += kFcThreads) {
        double w_, h_;
        const double area = fc_edge_rect(hull, k, e, w_, h_);
        if (area < best) { best = area; bbw = w_; bbh = h_; best_e = e; }
      }
    s_area[tid] = best; s_bw[tid] = bbw; s_bh[tid] = bbh; s_e[tid] = best_e;
    __syncthreads();
    if (tid != 0) return;
    if (k == 2) {
      const double dx = (double)(hull[1] >> 16) - (double)(hull[0] >> 16), dy = (double)(hull[1] & 0xffffu) - (double)(hull[0] & 0xffffu);
      const double len = sqrt(dx * dx + dy * dy);
      const bool along = (dy < 0 && dx >= 0) || (dy > 0 && dx <= 0);
      bw = along ? len : 0; bh = along ? 0 : len;
    } else if (k >= 3) {
      // the sequential scan keeps the FIRST edge with the strictly smallest area (it starts from best = 1e300)
      double ba = 1e300;
      int be = 0x7fffffff;
      for (int t = 0; t < kFcThreads; ++t)
        if (s_e[t] != 0x7fffffff && (s_area[t] < ba || (s_area[t] == ba && s_e[t] < be))) { ba = s_area[t]; be = s_e[t]; bw = s_bw[t]; bh = s_bh[t]; }
    }
  }
  const float fw = (float)bw, fh = (float)bh;
  const float box_size = fw > fh ? fw : fh;
  if (box_size < 10.0f) return;
  const float aspect = fw / (fh > 1.f ? fh : 1.f);
  if (aspect < 0.3f || aspect > 3.0f) return;
  const int im = (int)(ci.start / N);
  const int idx = atomicAdd(n_hyps, 1);
  if (idx < max_hyps) { hyps[idx].size = box_size; hyps[idx].slot_class = im * 2 + class_id; }
}

// ---------------------------------------------------------------------------------------------- per-image quad logic
struct ImageState {
  int found;            // processQuads succeeded in some attempt
  int prev_sqr_size;
  int n_quads;
  int hole_begin, hole_end;
  int total_quads;               // contours of this attempt that became quads
  unsigned long long best;       // (quads of the winning parent << 32) | (0x7fffffff - index of its last quad's hole)
};

// generateQuads, second half (CALIB_CB_FILTER_QUADS): the parent contour with the most quads wins.  OpenCV decides it in
// one sequential pass -- boardIdx moves to a parent when its running count strictly exceeds the current one's -- which
// ends at the parent that reaches the maximal count FIRST, i.e. among the parents with the maximal final count the one
// whose LAST quad comes earliest.  Exact per-parent counts (indexed by the parent's label = pixel index) and last-hole
// indices by atomics, the winner by one 64-bit atomicMax per image; the touched entries are cleared afterwards.
__global__ void cb_state_reset_kernel(ImageState* __restrict__ st, const int* __restrict__ active, int n_act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_act) return;
  ImageState& S = st[active[i]];
  S.total_quads = 0; S.best = 0ull; S.n_quads = 0;
}
__global__ void cb_parent_count_kernel(const HoleInfo* __restrict__ holes, int n_holes, int* __restrict__ pcnt, int* __restrict__ plast) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes || !holes[h].quad_ok) return;
  const int p = holes[h].parent;
  atomicAdd(&pcnt[p], 1);
  atomicMax(&plast[p], h + 1);
}
__global__ void cb_parent_select_kernel(const HoleInfo* __restrict__ holes, int n_holes, long long N, const int* __restrict__ pcnt,
                                        const int* __restrict__ plast, ImageState* __restrict__ st) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes || !holes[h].quad_ok) return;
  const int p = holes[h].parent;
  ImageState& S = st[(int)(holes[h].start / N)];
  atomicAdd(&S.total_quads, 1);
  if (plast[p] == h + 1)   // one candidate per parent is enough
    atomicMax(&S.best, ((unsigned long long)(unsigned)pcnt[p] << 32) | (unsigned)(0x7fffffff - h));
}
__global__ void cb_parent_clear_kernel(const HoleInfo* __restrict__ holes, int n_holes, int* __restrict__ pcnt, int* __restrict__ plast) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes || !holes[h].quad_ok) return;
  const int p = holes[h].parent;
  pcnt[p] = 0; plast[p] = 0;
}


__global__ void cb_hole_ranges_kernel(const HoleInfo* __restrict__ holes, int n_holes, long long N, int n_img, ImageState* __restrict__ st) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_holes) return;
  const int im = (int)(holes[h].start / N);
  if (h == 0 || (int)(holes[h - 1].start / N) != im) st[im].hole_begin = h;
  if (h == n_holes - 1 || (int)(holes[h + 1].start / N) != im) st[im].hole_end = h + 1;
}

struct WarpPar {
  __device__ int lane() const { return threadIdx.x & 31; }
  __device__ int lanes() const { return 32; }
  __device__ void sync() const { __syncwarp(); }
  __device__ void reduce_min(float& d, int& code) const {
    for (int m = 16; m >= 1; m >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, m);
      const int oc = __shfl_xor_sync(0xffffffffu, code, m);
      if (od < d || (od == d && oc < code)) { d = od; code = oc; }
    }
  }
  __device__ bool any(bool v) const { return __any_sync(0xffffffffu, v) != 0; }
};

// coarse phase timers of cb_process_kernel (ns, summed over images and attempts; printed when EZC_CB_TIMING is set)
__device__ unsigned long long g_cb_ns[6];
__device__ __forceinline__ unsigned long long cb_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// findQuadNeighbors for one image.  Same decisions as cb::find_quad_neighbors (cb_logic.h, the definition the CPU
// harness pins against cv2): the (quad, corner) sweep is sequential because every accepted link changes the state the
// next search sees.  The per-slot state the two scans of a step read (corner position, link, edge length) is staged in
// shared memory -- the global Work struct is two dependent loads away per candidate, which made a noisy attempt (1536
// quads => 6144 steps x 6144 candidates) cost ~400 ms -- and the scans only visit the grid cells their search disc
// touches (find_quad_neighbors_grid below): ~6 ms for such an attempt.
constexpr int kProcThreads = 256;
// 16-bit counters in shared memory (atomicAdd on the containing 32-bit word)
__device__ __forceinline__ void atomicAdd_u16(unsigned short* p) {
  unsigned int* w = reinterpret_cast<unsigned int*>(reinterpret_cast<size_t>(p) & ~size_t(3));
  atomicAdd(w, (reinterpret_cast<size_t>(p) & 2) ? 0x10000u : 1u);
}
__device__ __forceinline__ unsigned short fetch_inc_u16(unsigned short* p) {
  unsigned int* w = reinterpret_cast<unsigned int*>(reinterpret_cast<size_t>(p) & ~size_t(3));
  const bool hi = (reinterpret_cast<size_t>(p) & 2) != 0;
  const unsigned int old = atomicAdd(w, hi ? 0x10000u : 1u);
  return (unsigned short)(hi ? (old >> 16) : (old & 0xffffu));
}
struct NbShared {
  float2* pt;     // [4 * nq]  position of corner slot (quad, j)
  int* nb;        // [4 * nq]  linked quad or -1
  float* edge;    // [nq]
  int* count;     // [nq]
};
// Uniform grid over the corner slots (counting sort by cell): the two scans of a step only have to look at the cells a
// disc of radius sqrt(edge_len) (resp. sqrt(min_dist)) touches, so ONE warp does a step with shuffles only -- no block
// barriers.  Free slots never move (a link averages the two corners it consumes, and consumed slots are skipped), so the
// grid is built once per attempt.  Same decisions as the exhaustive scan: minimum over (distance, slot code).
constexpr int kGridCells = 4608;
struct NbGrid {
  unsigned short* start;   // [kGridCells + 1]
  unsigned short* cursor;  // [kGridCells]
  unsigned short* items;   // [4 * kMaxQuads] slot codes sorted by cell
  int cell, gw, gh;
};
__device__ __forceinline__ int grid_cell_of(const NbGrid& G, float x, float y) {
  int cx = (int)floorf(x / (float)G.cell), cy = (int)floorf(y / (float)G.cell);
  cx = min(max(cx, 0), G.gw - 1); cy = min(max(cy, 0), G.gh - 1);
  return cy * G.gw + cx;
}
template <class F>
__device__ __forceinline__ void grid_query(const NbGrid& G, float x, float y, float r, int lane, F&& f) {
  const float c = (float)G.cell;
  const int cx0 = min(max((int)floorf((x - r) / c), 0), G.gw - 1), cx1 = min(max((int)floorf((x + r) / c), 0), G.gw - 1);
  const int cy0 = min(max((int)floorf((y - r) / c), 0), G.gh - 1), cy1 = min(max((int)floorf((y + r) / c), 0), G.gh - 1);
  for (int cy = cy0; cy <= cy1; ++cy) {
    const int beg = G.start[cy * G.gw + cx0], end = G.start[cy * G.gw + cx1 + 1];
    for (int it = beg + lane; it < end; it += 32) f((int)G.items[it]);
  }
}
// cb::diagonal_neighbour on the staged slot positions
__device__ __forceinline__ bool diagonal_neighbour_s(const NbShared& S, int q, int i, float2 other, float2 other_diag) {
  const float2 a0 = S.pt[4 * q + i], a1 = S.pt[4 * q + ((i + 1) & 3)], a2 = S.pt[4 * q + ((i + 2) & 3)], a3 = S.pt[4 * q + ((i + 3) & 3)];
  const cb::Ptf c0{a0.x, a0.y}, c1{a1.x, a1.y}, c2{a2.x, a2.y}, c3{a3.x, a3.y}, o{other.x, other.y}, od{other_diag.x, other_diag.y};
  const cb::Ptf m1 = cb::mid_pt(c0, c1), m2 = cb::mid_pt(c2, c3), m3 = cb::mid_pt(c1, c2), m4 = cb::mid_pt(c3, c0);
  return cb::same_side_of_line(m1, m2, c0, o) && cb::same_side_of_line(m3, m4, c0, o) && cb::same_side_of_line(m1, m2, c0, od) &&
         cb::same_side_of_line(m3, m4, c0, od);
}
__device__ void find_quad_neighbors_grid(cb::Work& w, const NbShared& S, const NbGrid& G) {
  const int nq = w.n_quads;
  const int tid = threadIdx.x;
  // ---- stage the per-slot state and build the grid (whole block)
  for (int k = tid; k < nq; k += kProcThreads) {
    const cb::Quad& q = w.quads[k];
    S.edge[k] = q.edge_len;
    S.count[k] = q.count;
    for (int j = 0; j < 4; ++j) {
      const cb::Ptf p = w.corners[q.corners[j]].pt;
      S.pt[4 * k + j] = make_float2(p.x, p.y);
      S.nb[4 * k + j] = q.nb[j];
    }
  }
  for (int c = tid; c < kGridCells; c += kProcThreads) G.cursor[c] = 0;
  __syncthreads();
  for (int sl = tid; sl < 4 * nq; sl += kProcThreads) atomicAdd_u16(&G.cursor[grid_cell_of(G, S.pt[sl].x, S.pt[sl].y)]);
  __syncthreads();
  if (tid < 32) {   // exclusive scan of the cell counts by one warp (144 cells per lane)
    constexpr int per = kGridCells / 32;
    int sum = 0;
    for (int c = tid * per; c < (tid + 1) * per; ++c) sum += G.cursor[c];
    int inc = sum;
    for (int m = 1; m < 32; m <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, m); if (tid >= m) inc += o; }
    int run = inc - sum;
    for (int c = tid * per; c < (tid + 1) * per; ++c) { const int n = G.cursor[c]; G.start[c] = (unsigned short)run; G.cursor[c] = (unsigned short)run; run += n; }
    if (tid == 31) G.start[kGridCells] = (unsigned short)run;
  }
  __syncthreads();
  for (int sl = tid; sl < 4 * nq; sl += kProcThreads) {
    const int c = grid_cell_of(G, S.pt[sl].x, S.pt[sl].y);
    G.items[fetch_inc_u16(&G.cursor[c])] = (unsigned short)sl;
  }
  __syncthreads();
  // ---- the sequential sweep, one warp
  if (tid < 32) {
    const int lane = tid;
    for (int idx = 0; idx < nq; idx++) {
      for (int i = 0; i < 4; i++) {
        if (S.nb[4 * idx + i] >= 0) continue;
        float min_dist = FLT_MAX;
        int code = 0x7fffffff;
        const float2 pt = S.pt[4 * idx + i];
        const float cur_edge = S.edge[idx];
        grid_query(G, pt.x, pt.y, sqrtf(cur_edge * cb::kNbSqrScale) + 1.f, lane, [&](int sl) {
          const int k = sl >> 2;
          if (k == idx || S.nb[sl] >= 0) return;
          const float ek = S.edge[k];
          const float2 o = S.pt[sl];
          const float dist = cb::normL2Sqrf(pt.x - o.x, pt.y - o.y);
          if ((dist < min_dist || (dist == min_dist && sl < code)) && dist <= cur_edge * cb::kNbSqrScale && dist <= ek * cb::kNbSqrScale) {
            if (ek > 16 * cur_edge || cur_edge > 16 * ek) return;
            if (!diagonal_neighbour_s(S, idx, i, o, S.pt[4 * k + (((sl & 3) + 2) & 3)])) return;
            if (ek * cb::kNbSqrScale >= cur_edge && !diagonal_neighbour_s(S, k, sl & 3, pt, S.pt[4 * idx + ((i + 2) & 3)])) return;   // see cb_logic.h
            code = sl; min_dist = dist;
          }
        });
        for (int m = 16; m >= 1; m >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, min_dist, m);
          const int oc = __shfl_xor_sync(0xffffffffu, code, m);
          if (od < min_dist || (od == min_dist && oc < code)) { min_dist = od; code = oc; }
        }
        if (code == 0x7fffffff || !(min_dist < FLT_MAX)) continue;
        const int cq = code >> 2, cci = code & 3;
        if (S.count[idx] >= 4 || S.count[cq] >= 4) continue;
        const float2 cc = S.pt[4 * cq + cci];
        int j = 0;
        for (; j < 4; j++) {
          if (S.nb[4 * idx + j] == cq) break;
          const float2 o = S.pt[4 * idx + j];
          if (cb::normL2Sqrf(cc.x - o.x, cc.y - o.y) < min_dist) break;
        }
        if (j < 4) continue;
        for (j = 0; j < 4; j++) if (S.nb[4 * cq + j] == idx) break;
        if (j < 4) continue;
        bool blocked = false;
        grid_query(G, cc.x, cc.y, sqrtf(min_dist) + 1.f, lane, [&](int sl) {
          const int q3 = sl >> 2;
          if (q3 == idx || q3 == cq || S.nb[sl] >= 0) return;
          const float2 o = S.pt[sl];
          if (cb::normL2Sqrf(cc.x - o.x, cc.y - o.y) < min_dist && diagonal_neighbour_s(S, cq, cci, o, S.pt[4 * q3 + (((sl & 3) + 2) & 3)]))
            blocked = true;
        });
        if (__any_sync(0xffffffffu, blocked)) continue;
        if (lane == 0) {
          cb::Quad& cur = w.quads[idx];
          cb::Quad& cqq = w.quads[cq];
          cb::Corner& ccn = w.corners[cqq.corners[cci]];
          const float nx = (pt.x + cc.x) * 0.5f, ny = (pt.y + cc.y) * 0.5f;
          ccn.pt.x = nx; ccn.pt.y = ny;
          cur.count++;
          cur.nb[i] = cq;
          cur.corners[i] = cqq.corners[cci];
          cqq.count++;
          cqq.nb[cci] = idx;
          S.pt[4 * cq + cci] = make_float2(nx, ny);
          S.pt[4 * idx + i] = make_float2(nx, ny);
          S.nb[4 * idx + i] = cq;
          S.nb[4 * cq + cci] = idx;
          S.count[idx]++;
          S.count[cq]++;
        }
        __syncwarp();
      }
    }
  }
  __threadfence_block();
  __syncthreads();
}

// one BLOCK per active image
__global__ void __launch_bounds__(kProcThreads) cb_process_kernel(const HoleInfo* __restrict__ holes, ImageState* __restrict__ st,
                                                                  const int* __restrict__ active, int n_img, int pw, int ph,
                                                                  const int* __restrict__ dil_of, cb::Work* __restrict__ works, int* __restrict__ scratch_all,
                                                                  cb::Ptf* __restrict__ corners_out, int W, int H) {
  extern __shared__ __align__(16) unsigned char cb_smem[];
  const int slot = blockIdx.x;
  const int tid = threadIdx.x;
  if (slot >= n_img) return;
  const int im = active[slot];
  const int dilations = dil_of[im];   // generateQuads' edge_len compensation is per attempt
  cb::Work& w = works[im];
  ImageState& S = st[im];
  int* scratch = scratch_all + (size_t)im * cb::kMaxQuads * 4;
  const unsigned long long t0 = cb_now();
  // the winning parent was selected by cb_parent_select_kernel; its quads are appended in hole order (the order
  // generateQuads visits the contours) by the whole block: flags, block-wide exclusive scan, independent quad set-up
  __shared__ int scan_w[kProcThreads / 32];
  __shared__ int s_base;
  const int total = S.total_quads;
  int max_quads = total * 3 > 2 ? total * 3 : 2;
  if (max_quads > cb::kMaxQuads) max_quads = cb::kMaxQuads;
  const int boardIdx = S.best ? holes[0x7fffffff - (int)(unsigned)(S.best & 0xffffffffull)].parent : -1;
  if (tid == 0) { s_base = 0; w.max_quads = max_quads; w.pw = pw; w.ph = ph; }
  __syncthreads();
  for (int h0 = S.hole_begin; h0 < S.hole_end; h0 += kProcThreads) {
    const int h = h0 + tid;
    const bool take = h < S.hole_end && holes[h].quad_ok && holes[h].parent == boardIdx;
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) scan_w[warp] = __popc(bal);
    __syncthreads();
    int before = s_base, all = 0;
    for (int wv = 0; wv < kProcThreads / 32; ++wv) { if (wv < warp) before += scan_w[wv]; all += scan_w[wv]; }
    const int idx = before + __popc(bal & ((1u << lane) - 1u));
    if (take && idx < max_quads) cb::add_quad_at(w, idx, holes[h].quad, dilations);
    __syncthreads();
    if (tid == 0) s_base += all;
    __syncthreads();
    if (s_base >= max_quads) break;
  }
  if (tid == 0) {
    w.n_quads = s_base < max_quads ? s_base : max_quads;
    S.n_quads = w.n_quads;
    atomicAdd(&g_cb_ns[0], cb_now() - t0);
  }
  __threadfence_block();
  __syncthreads();
  const unsigned long long t1 = cb_now();
  if (tid == 0) atomicAdd(&g_cb_ns[1], t1 - t0);
  // a complete board needs all its inner squares: fewer quads than that cannot succeed (and leave no trace in
  // prev_sqr_size either), so the quad-graph stage is skipped
  const int nq = w.n_quads;
  if (nq < ((pw - 1) * (ph - 1)) / 2) return;
  {
    NbShared sh;
    sh.pt = reinterpret_cast<float2*>(cb_smem);
    sh.nb = reinterpret_cast<int*>(sh.pt + 4 * cb::kMaxQuads);
    sh.edge = reinterpret_cast<float*>(sh.nb + 4 * cb::kMaxQuads);
    sh.count = reinterpret_cast<int*>(sh.edge + cb::kMaxQuads);
    NbGrid G;
    G.start = reinterpret_cast<unsigned short*>(sh.count + cb::kMaxQuads);
    G.cursor = G.start + kGridCells + 2;
    G.items = G.cursor + kGridCells;
    G.cell = max(16, (int)ceilf(sqrtf((float)W * (float)H / 4096.f)));
    G.gw = (W + G.cell - 1) / G.cell; G.gh = (H + G.cell - 1) / G.cell;
    while (G.gw * G.gh > kGridCells) { ++G.cell; G.gw = (W + G.cell - 1) / G.cell; G.gh = (H + G.cell - 1) / G.cell; }
    find_quad_neighbors_grid(w, sh, G);
  }
  const unsigned long long t2 = cb_now();
  if (tid == 0) atomicAdd(&g_cb_ns[2], t2 - t1);
  if (tid == 0) {
    int n_out = 0, prev = S.prev_sqr_size;
    cb::Ptf* out = corners_out + (size_t)im * pw * ph;
    cb::Ptf* tmp = reinterpret_cast<cb::Ptf*>(w.hull);  // partial corner sets are written too; copy on success only
    const bool ok = cb::process_groups(w, tmp, &n_out, &prev, scratch);
    S.prev_sqr_size = prev;
    if (ok) {
      for (int i = 0; i < pw * ph; ++i) out[i] = tmp[i];
      S.found = 1;
    }
    atomicAdd(&g_cb_ns[3], cb_now() - t2);
  }
}
constexpr size_t kProcSmem = (size_t)cb::kMaxQuads * (4 * sizeof(float2) + 4 * sizeof(int) + sizeof(float) + sizeof(int)) +
                             sizeof(unsigned short) * (2 * kGridCells + 4 + 4 * cb::kMaxQuads);

// final checks of cv::findChessboardCorners after a successful attempt: board monotony, 8-px border, even/even flip
__global__ void cb_finalize_kernel(ImageState* __restrict__ st, int n_img, int pw, int ph, int W, int H, cb::Ptf* __restrict__ corners,
                                   uint8_t* __restrict__ found_out) {
  const int im = blockIdx.x * blockDim.x + threadIdx.x;
  if (im >= n_img) return;
  cb::Ptf* c = corners + (size_t)im * pw * ph;
  bool found = st[im].found != 0;
  if (found) found = cb::check_board_monotony(c, pw, ph);
  if (found) {
    const int BORDER = 8;
    for (int k = 0; k < pw * ph; ++k)
      if (c[k].x <= BORDER || c[k].x > W - BORDER || c[k].y <= BORDER || c[k].y > H - BORDER) { found = false; break; }
  }
  if (found && (ph & 1) == 0 && (pw & 1) == 0) {
    const int last_row = (ph - 1) * pw;
    const double dy0 = c[last_row].y - c[0].y;
    if (dy0 < 0) {
      const int n = pw * ph;
      for (int i = 0; i < n / 2; i++) { const cb::Ptf t = c[i]; c[i] = c[n - i - 1]; c[n - i - 1] = t; }
    }
  }
  found_out[im] = found ? 1 : 0;
  if (!found) for (int k = 0; k < pw * ph; ++k) c[k] = cb::Ptf{0.f, 0.f};
}

// ---------------------------------------------------------------------------------------------- fallback image ops
// cv::equalizeHist LUT from the histogram (one thread per image)
__global__ void cb_equalize_lut_kernel(const int* __restrict__ hist_all, int n_img, int total, uint8_t* __restrict__ lut_all) {
  const int im = blockIdx.x * blockDim.x + threadIdx.x;
  if (im >= n_img) return;
  const int* hist = hist_all + im * 256;
  uint8_t* lut = lut_all + im * 256;
  int i = 0;
  while (!hist[i]) ++i;
  if (hist[i] == total) { for (int k = 0; k < 256; ++k) lut[k] = (uint8_t)i; return; }
  const float scale = (256 - 1.f) / (total - hist[i]);
  int sum = 0;
  for (int k = 0; k <= i; ++k) lut[k] = 0;
  for (++i; i < 256; ++i) {
    sum += hist[i];
    const int v = __float2int_rn(sum * scale);
    lut[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
  }
}
__global__ void cb_apply_lut_kernel(const uint8_t* __restrict__ img, long long pitch, long long istride, bool vec_src, int W, int H,
                                    const int* __restrict__ active, const uint8_t* __restrict__ lut, uint8_t* __restrict__ out, bool vec_dst) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  Chunk16 v = ld16(img + (long long)im * istride + (long long)c.y * pitch, c.x0, W, vec_src, 0u);
  const uint8_t* T = lut + im * 256;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t w = v.w[k];
    v.w[k] = (uint32_t)T[w & 0xff] | ((uint32_t)T[(w >> 8) & 0xff] << 8) | ((uint32_t)T[(w >> 16) & 0xff] << 16) | ((uint32_t)T[w >> 24] << 24);
  }
  st16(out + ((long long)im * H + c.y) * W, c.x0, W, vec_dst, v);
}
// cv::adaptiveThreshold(MEAN_C, BINARY): separable box sums with replicated border, per-image block size (up to 0.2 x
// the short image side: ~200 taps).  Horizontal pass: one block per row, inclusive prefix sums of the row in shared
// memory, each output the difference of two of them plus the replicated end pixels.  Vertical pass: one thread per
// column and 32-row segment, a running sum that adds the row entering the window and drops the one leaving it.
// Integer sums, so the result does not depend on how they are formed.
// `src_of` (optional) names the image each listed slot reads: a speculative attempt works in a slot of its own on the
// equalised copy of the view it belongs to.
constexpr int kBoxRowThreads = 256;
__global__ void __launch_bounds__(kBoxRowThreads) cb_box_h_kernel(const uint8_t* __restrict__ src, int W, int H, const int* __restrict__ list,
                                                                  const int* __restrict__ block, int* __restrict__ out,
                                                                  const int* __restrict__ src_of) {
  extern __shared__ int box_pre[];   // W + 1 prefix sums, then one partial per warp
  const int y = blockIdx.x, im = list[blockIdx.y];
  const long long N = (long long)W * H;
  const uint8_t* S = src + (long long)(src_of ? src_of[im] : im) * N + (long long)y * W;
  const int r = block[im] / 2;
  int* warp_tot = box_pre + W + 1;
  __shared__ int carry;
  if (threadIdx.x == 0) { carry = 0; box_pre[0] = 0; }
  __syncthreads();
  for (int x0 = 0; x0 < W; x0 += kBoxRowThreads) {
    const int x = x0 + threadIdx.x;
    int v = x < W ? (int)S[x] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    int before = carry;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (x < W) box_pre[x + 1] = before + v;
    __syncthreads();
    if (threadIdx.x == kBoxRowThreads - 1) carry = before + v;
    __syncthreads();
  }
  const int first = (int)S[0], last = (int)S[W - 1];
  int* O = out + (long long)im * N + (long long)y * W;
  for (int x = threadIdx.x; x < W; x += kBoxRowThreads) {
    const int lo = max(0, x - r), hi = min(W - 1, x + r);
    O[x] = box_pre[hi + 1] - box_pre[lo] + max(0, r - x) * first + max(0, x + r - (W - 1)) * last;
  }
}
constexpr int kBoxSegRows = 32;
__global__ void cb_box_v_thresh_kernel(const uint8_t* __restrict__ src, const int* __restrict__ hs, int W, int H,
                                       const int* __restrict__ list, const int* __restrict__ block, const int* __restrict__ delta,
                                       uint8_t* __restrict__ out, const int* __restrict__ src_of) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int im = list[blockIdx.z];
  const int y0 = blockIdx.y * kBoxSegRows, y1 = min(H, y0 + kBoxSegRows);
  const long long N = (long long)W * H;
  const int b = block[im], r = b / 2, dlt = delta[im];
  const int* S = hs + (long long)im * N + x;
  const uint8_t* I = src + (long long)(src_of ? src_of[im] : im) * N + x;
  uint8_t* O = out + (long long)im * N + x;
  const float scale = (float)(1.0 / ((double)b * b));
  int s = 0;
  for (int d = -r; d <= r; ++d) s += S[(long long)min(H - 1, max(0, y0 + d)) * W];
  for (int y = y0; y < y1; ++y) {
    int mean = __float2int_rn((float)s * scale);
    mean = mean > 255 ? 255 : mean;
    O[(long long)y * W] = ((int)I[(long long)y * W] - mean > -dlt) ? 255 : 0;
    s += S[(long long)min(H - 1, y + 1 + r) * W] - S[(long long)max(0, y - r) * W];
  }
}
// (src and dst may be the same buffer when `src_of` moves whole images between slots: no slot is then both read and written)
// 16 bytes per thread when the image size allows (every image then starts 16-byte aligned in the packed buffers).
__global__ void cb_copy_active_kernel(const uint8_t* src, uint8_t* dst, long long N, int n_img, const int* __restrict__ list,
                                      const int* __restrict__ src_of) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int im = list[blockIdx.y];
  const long long so = (long long)(src_of ? src_of[im] : im) * N, dof = (long long)im * N;
  if ((N & 15) == 0) {
    if (i * 16 < N) reinterpret_cast<uint4*>(dst + dof)[i] = reinterpret_cast<const uint4*>(src + so)[i];
  } else {
    for (long long k = i * 16; k < min(N, i * 16 + 16); ++k) dst[dof + k] = src[so + k];
  }
}
// Results of the views that finished, kept apart from the per-slot state while speculative attempts reuse the slots.
__global__ void cb_save_results_kernel(const ImageState* __restrict__ st, const cb::Ptf* __restrict__ corners, int n_img, int npts,
                                       int* __restrict__ res_found, cb::Ptf* __restrict__ res_corners) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * npts) return;
  const int im = i / npts;
  if (i == im * npts) res_found[im] = st[im].found;
  res_corners[i] = corners[i];
}
// pairs[2 j] = slot whose attempt is the one the sequential search would have stopped at, pairs[2 j + 1] = its view
__global__ void cb_accept_kernel(const int* __restrict__ pairs, int n_pairs, int npts, const cb::Ptf* __restrict__ corners,
                                 int* __restrict__ res_found, cb::Ptf* __restrict__ res_corners) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs * npts) return;
  const int j = i / npts, k = i - j * npts;
  const int slot = pairs[2 * j], view = pairs[2 * j + 1];
  if (k == 0) res_found[view] = 1;
  res_corners[(size_t)view * npts + k] = corners[(size_t)slot * npts + k];
}
__global__ void cb_restore_results_kernel(ImageState* __restrict__ st, cb::Ptf* __restrict__ corners, int n_img, int npts,
                                          const int* __restrict__ res_found, const cb::Ptf* __restrict__ res_corners) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * npts) return;
  const int im = i / npts;
  if (i == im * npts) st[im].found = res_found[im];
  corners[i] = res_corners[i];
}
__global__ void cb_gather_corners_kernel(const cb::Ptf* __restrict__ corners, const uint8_t* __restrict__ found, int n_img, int npts,
                                         float2* __restrict__ out, int* __restrict__ img_index) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * npts) return;
  const int im = i / npts;
  // images without a board keep a harmless in-image position so that the refinement kernel can run over the batch
  out[i] = found[im] ? make_float2(corners[i].x, corners[i].y) : make_float2(16.f, 16.f);
  img_index[i] = im;
}
__global__ void cb_scatter_corners_kernel(const float2* __restrict__ refined, const uint8_t* __restrict__ found, int n_img, int npts,
                                          float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * npts) return;
  out[i] = found[i / npts] ? refined[i] : make_float2(0.f, 0.f);
}

}  // namespace ezc

using namespace ezc;

namespace {

struct CbPools {
  DevBuf hist, thresh, bin0, bin1, prevbin, eq, lut, labels, flags, scan_tmp, total, hole_start, holes, need, pool, stack_pool;
  DevBuf works, scratch, state, active, corners, found, block, delta, hsum, sub_corners, sub_idx, refresh;
  DevBuf pcnt, plast;   // per-pixel-label quad counters (kept zero between attempts)
  DevBuf blk_cnt, blk_off;   // roots per block of the chunk grid and their exclusive scan
  DevBuf dil, srcmap, res_found, res_corners, pairs;   // per-slot dilation counts / source views, results of finished views
  DevBuf qlist, qcount;      // work lists of the holes with long contours
  DevBuf fc_w0, fc_w1, fc_w2, fc_k0, fc_k1, fc_k2, fc_fg, fc_comps, fc_hyps, fc_count, fc_list;
  HostPinned fc_h;
  HostPinned h;
};

CbPools& cb_pools(ezc_context* ctx) {
  if (!ctx->cb_pools) {
    ctx->cb_pools = new CbPools();
    ctx->cb_pools_free = [](void* p) { delete static_cast<CbPools*>(p); };
  }
  return *static_cast<CbPools*>(ctx->cb_pools);
}

#define CB_LAUNCH(kernel, n, ...)                                                     \
  do {                                                                                \
    if ((n) > 0) {                                                                    \
      const long long _n = (n);                                                       \
      kernel<<<(unsigned)((_n + 255) / 256), 256, 0, ctx->stream>>>(__VA_ARGS__);     \
      EZC_CUDA(cudaGetLastError());                                                   \
      ctx->kernel_launches++;                                                         \
    }                                                                                 \
  } while (0)

// per-pixel kernels: grid.x covers one image, grid.y the images (no 64-bit division to split a flat index)
#define CB_LAUNCH_PIX(kernel, n_pix, n_images, ...)                                                         \
  do {                                                                                                      \
    if ((n_pix) > 0 && (n_images) > 0) {                                                                    \
      kernel<<<dim3((unsigned)(((n_pix) + 255) / 256), (unsigned)(n_images)), 256, 0, ctx->stream>>>(__VA_ARGS__); \
      EZC_CUDA(cudaGetLastError());                                                                         \
      ctx->kernel_launches++;                                                                               \
    }                                                                                                       \
  } while (0)

// 16-pixel row chunks (cb_rows.cuh): grid.x covers the chunks of one image, grid.y the images
#define CB_LAUNCH_CH(kernel, W_, H_, n_images, ...)                                                         \
  do {                                                                                                      \
    if ((n_images) > 0) {                                                                                   \
      kernel<<<dim3(chunk_blocks((W_), (H_)), (unsigned)(n_images)), 256, 0, ctx->stream>>>(__VA_ARGS__);    \
      EZC_CUDA(cudaGetLastError());                                                                         \
      ctx->kernel_launches++;                                                                               \
    }                                                                                                       \
  } while (0)

#define CB_LAUNCH_FRAME(img_, n_images, active_)                                                                                       \
  do {                                                                                                                                  \
    if ((n_images) > 0) {                                                                                                               \
      cb_frame_border_kernel<<<dim3((unsigned)((6 * (W + H) + 255) / 256), (unsigned)(n_images)), 256, 0, ctx->stream>>>((img_), W, H, (active_)); \
      EZC_CUDA(cudaGetLastError());                                                                                                     \
      ctx->kernel_launches++;                                                                                                           \
    }                                                                                                                                   \
  } while (0)

// Upload a compact list of image indices; returns its device pointer.
ezc_status upload_list(ezc_context* ctx, DevBuf& buf, const std::vector<int>& list) {
  EZC_CUDA(buf.reserve(sizeof(int) * std::max<size_t>(list.size(), 1)));
  if (!list.empty()) EZC_CUDA(cudaMemcpyAsync(buf.p, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
  return EZC_OK;
}

// cv::adaptiveThreshold of the equalised images named by `list` (n entries) into `dst`, block sizes / deltas per slot
ezc_status adaptive_threshold(ezc_context* ctx, CbPools& P, int W, int H, int n, const int* list, uint8_t* dst, const int* src_of) {
  if (n <= 0) return EZC_OK;
  const size_t smem = sizeof(int) * ((size_t)W + 1 + kBoxRowThreads / 32);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    EZC_CUDA(cudaFuncSetAttribute(cb_box_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cb_box_h_kernel<<<dim3((unsigned)H, (unsigned)n), kBoxRowThreads, smem, ctx->stream>>>(P.eq.as<uint8_t>(), W, H, list, P.block.as<int>(), P.hsum.as<int>(), src_of);
  EZC_CUDA(cudaGetLastError());
  cb_box_v_thresh_kernel<<<dim3((unsigned)((W + 127) / 128), (unsigned)((H + kBoxSegRows - 1) / kBoxSegRows), (unsigned)n), 128, 0, ctx->stream>>>(
      P.eq.as<uint8_t>(), P.hsum.as<int>(), W, H, list, P.block.as<int>(), P.delta.as<int>(), dst, src_of);
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches += 2;
  return EZC_OK;
}

// One attempt for the images in `list` (device copy in P.active, n_act entries): quads from `bin` (already dilated +
// framed) -> processQuads.  Per-pixel buffers are indexed by the REAL image index; only the hole flags are compact.
ezc_status run_attempt(ezc_context* ctx, CbPools& P, const uint8_t* bin, int W, int H, int n_act, int pw, int ph, const int* d_dil) {
  if (n_act <= 0) return EZC_OK;
  const bool timing = getenv("EZC_CB_TIMING") != nullptr && getenv("EZC_CB_TIMING")[0] == '2';
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(ctx->stream);
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[cb]   %-12s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
    t_prev = t;
  };
  const long long N = (long long)W * H;
  const int* list = P.active.as<int>();
  int* d_tot = P.total.as<int>();
  const bool vec = (W % 16) == 0;
  const long long n_blk = (long long)chunk_blocks(W, H) * n_act;
  EZC_CUDA(P.blk_cnt.reserve(sizeof(int) * (size_t)n_blk)); EZC_CUDA(P.blk_off.reserve(sizeof(int) * (size_t)n_blk));
  CB_LAUNCH(cb_label_runs_kernel, (long long)H * n_act * 32, bin, P.labels.as<int>(), W, H, n_act, list);
  CB_LAUNCH_CH(cb_label_merge16_kernel, W, H, n_act, bin, P.labels.as<int>(), W, H, list, vec);
  CB_LAUNCH_CH(cb_label_heads_kernel, W, H, n_act, bin, P.labels.as<int>(), W, H, list, vec, 0, P.blk_cnt.as<int>());
  if (ezc_status st = exclusive_scan_i32(ctx, P.blk_cnt.as<int>(), P.blk_off.as<int>(), n_blk, P.scan_tmp, d_tot)) return st;
  EZC_CUDA(cudaMemcpyAsync(P.h.p, d_tot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int nH = *P.h.as<int>();
  lap("labels");
  ImageState* st = P.state.as<ImageState>();
  if (nH > 0) {
    EZC_CUDA(P.hole_start.reserve(sizeof(int) * (size_t)nH));
    EZC_CUDA(P.holes.reserve(sizeof(HoleInfo) * (size_t)nH));
    EZC_CUDA(P.need.reserve(sizeof(int) * ((size_t)nH + 1)));
    CB_LAUNCH_CH(cb_root_list_kernel, W, H, n_act, bin, P.labels.as<int>(), W, H, list, vec, 0, P.blk_cnt.as<int>(), P.blk_off.as<int>(), P.hole_start.as<int>());
    CB_LAUNCH(cb_hole_init_kernel, nH, P.hole_start.as<int>(), nH, P.holes.as<HoleInfo>(), P.need.as<int>());
    CB_LAUNCH(cb_trace_kernel<false>, nH, bin, P.labels.as<int>(), W, H, nH, P.holes.as<HoleInfo>(), nullptr, cb::kMaxContourPts);
    CB_LAUNCH(cb_hole_need_kernel, nH, P.holes.as<HoleInfo>(), nH, P.need.as<int>());
    if (ezc_status s2 = exclusive_scan_i32(ctx, P.need.as<int>(), P.need.as<int>(), nH, P.scan_tmp, d_tot)) return s2;
    EZC_CUDA(cudaMemcpyAsync(P.h.p, d_tot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nPts = *P.h.as<int>();
    lap("trace count");
    if (timing) fprintf(stderr, "[cb]   %d contours, %d points\n", nH, nPts);
    EZC_CUDA(P.pool.reserve(sizeof(cb::Pt) * ((size_t)nPts + 16)));
    EZC_CUDA(P.stack_pool.reserve(sizeof(int) * 2 * ((size_t)nPts + 16)));
    EZC_CUDA(P.qlist.reserve(sizeof(int) * 2 * (size_t)nH)); EZC_CUDA(P.qcount.reserve(64));
    EZC_CUDA(cudaMemsetAsync(P.qcount.p, 0, 2 * sizeof(int), ctx->stream));
    int* list_mid = P.qlist.as<int>();
    int* list_big = list_mid + nH;
    CB_LAUNCH(cb_hole_offset_kernel, nH, P.holes.as<HoleInfo>(), nH, P.need.as<int>(), list_mid, list_big, P.qcount.as<int>());
    CB_LAUNCH(cb_trace_kernel<true>, nH, bin, P.labels.as<int>(), W, H, nH, P.holes.as<HoleInfo>(), P.pool.as<cb::Pt>(), cb::kMaxContourPts);
    lap("trace store");
    CB_LAUNCH(cb_quad_kernel, nH, P.holes.as<HoleInfo>(), nH, P.pool.as<cb::Pt>(), P.stack_pool.as<int>());
    lap("quads small");
    {
      static bool qw_attr = false;
      if (!qw_attr) {
        EZC_CUDA(cudaFuncSetAttribute(cb_quad_warp_kernel<cb::kMaxContourPts>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)quad_warp_smem<cb::kMaxContourPts>()));
        qw_attr = true;
      }
      const int g_mid = std::min(nH, ctx->sm_count * 16), g_big = std::min(nH, ctx->sm_count * 3);
      cb_quad_warp_kernel<kWarpContourMid><<<g_mid, 32, quad_warp_smem<kWarpContourMid>(), ctx->stream>>>(P.holes.as<HoleInfo>(), list_mid,
                                                                                                          P.qcount.as<int>(), P.pool.as<cb::Pt>());
      EZC_CUDA(cudaGetLastError());
      cb_quad_warp_kernel<cb::kMaxContourPts><<<g_big, 32, quad_warp_smem<cb::kMaxContourPts>(), ctx->stream>>>(
          P.holes.as<HoleInfo>(), list_big, P.qcount.as<int>() + 1, P.pool.as<cb::Pt>());
      EZC_CUDA(cudaGetLastError());
      ctx->kernel_launches += 2;
    }
    lap("quads warp");
    CB_LAUNCH(cb_hole_ranges_kernel, nH, P.holes.as<HoleInfo>(), nH, N, 0, st);
  }
  CB_LAUNCH(cb_state_reset_kernel, n_act, st, list, n_act);
  if (nH > 0) {
    CB_LAUNCH(cb_parent_count_kernel, nH, P.holes.as<HoleInfo>(), nH, P.pcnt.as<int>(), P.plast.as<int>());
    CB_LAUNCH(cb_parent_select_kernel, nH, P.holes.as<HoleInfo>(), nH, N, P.pcnt.as<int>(), P.plast.as<int>(), st);
    CB_LAUNCH(cb_parent_clear_kernel, nH, P.holes.as<HoleInfo>(), nH, P.pcnt.as<int>(), P.plast.as<int>());
  }
  {
    static bool attr_set = false;
    if (!attr_set) {
      EZC_CUDA(cudaFuncSetAttribute(cb_process_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProcSmem));
      attr_set = true;
    }
    cb_process_kernel<<<n_act, kProcThreads, kProcSmem, ctx->stream>>>(P.holes.as<HoleInfo>(), st, list, n_act, pw, ph, d_dil,
                                                                       P.works.as<cb::Work>(), P.scratch.as<int>(), P.corners.as<cb::Ptf>(), W, H);
    EZC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  lap("process");
  return EZC_OK;
}

// refresh the host copy of the per-image state and rebuild the compact list of images still searching
ezc_status sync_state(ezc_context* ctx, CbPools& P, int nb, std::vector<ImageState>& h_state, std::vector<int>& list) {
  EZC_CUDA(cudaMemcpyAsync(h_state.data(), P.state.p, sizeof(ImageState) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int> next;
  for (int im : list) if (!h_state[im].found) next.push_back(im);
  list.swap(next);
  for (int i = 0; i < nb; ++i) { h_state[i].hole_begin = 0; h_state[i].hole_end = 0; }
  EZC_CUDA(cudaMemcpyAsync(P.state.p, h_state.data(), sizeof(ImageState) * nb, cudaMemcpyHostToDevice, ctx->stream));
  return upload_list(ctx, P.active, list);
}

// checkQuads (checkchessboard.cpp) on the host: a few hundred (size, class) pairs per image.
static bool fc_check_quads(std::vector<std::pair<float, int>>& quads, int pw, int ph) {
  const size_t min_quads_count = (size_t)(pw * ph / 2);
  std::sort(quads.begin(), quads.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
  const float size_rel_dev = 0.4f;
  const int black_count = (int)lrint(std::ceil(pw / 2.0) * std::ceil(ph / 2.0));
  const int white_count = (int)lrint(std::floor(pw / 2.0) * std::floor(ph / 2.0));
  for (size_t i = 0; i < quads.size(); i++) {
    size_t j = i + 1;
    for (; j < quads.size(); j++)
      if (quads[j].first / quads[i].first > 1.0f + size_rel_dev) break;
    if (j + 1 > min_quads_count + i) {
      int counts[2] = {0, 0};
      for (size_t k = i; k != j; k++) counts[quads[k].second]++;
      if (counts[0] < black_count * 0.75 || counts[1] < white_count * 0.75) continue;
      return true;
    }
  }
  return false;
}

// fillQuads + checkQuads for the images in `alist` (real indices); images that pass are flagged and leave the list.
static ezc_status fc_round(ezc_context* ctx, CbPools& P, const uint8_t* white_src, const uint8_t* black_src, int white_thresh, int black_thresh,
                           int W, int H, int pw, int ph, std::vector<int>& alist, std::vector<char>& pass) {
  const int n_act = (int)alist.size();
  if (n_act == 0) return EZC_OK;
  const long long N = (long long)W * H, total = N * n_act;
  if (ezc_status st = upload_list(ctx, P.fc_list, alist)) return st;
  const int* list = P.fc_list.as<int>();
  const bool vec = (W % 16) == 0;
  const int max_hyps = (int)std::min<long long>(4 << 20, total / 64 + 4096);
  EZC_CUDA(P.fc_hyps.reserve(sizeof(FcHyp) * (size_t)max_hyps));
  EZC_CUDA(P.fc_count.reserve(64));
  EZC_CUDA(cudaMemsetAsync(P.fc_count.p, 0, sizeof(int), ctx->stream));
  int* d_tot = P.total.as<int>();
  uint8_t* fg = P.fc_fg.as<uint8_t>();
  for (int cls = 1; cls >= 0; --cls) {
    CB_LAUNCH_CH(cb_fg16_kernel, W, H, n_act, cls ? white_src : black_src, fg, W, H, list, cls ? white_thresh : black_thresh, cls ? 0 : 1, vec);
    CB_LAUNCH(cb_label_runs_kernel, (long long)H * n_act * 32, fg, P.labels.as<int>(), W, H, n_act, list);
    CB_LAUNCH_CH(cb_label_merge16_kernel, W, H, n_act, fg, P.labels.as<int>(), W, H, list, vec);
    EZC_CUDA(cudaMemsetAsync(d_tot, 0, sizeof(int), ctx->stream));
    CB_LAUNCH_CH(cb_fc_roots16_kernel, W, H, n_act, fg, P.labels.as<int>(), W, H, list, vec, P.flags.as<int>(), d_tot);
    EZC_CUDA(cudaMemcpyAsync(P.h.p, d_tot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nC = *P.h.as<int>();
    if (nC <= 0) continue;
    EZC_CUDA(P.fc_comps.reserve(sizeof(FcComp) * (size_t)nC));
    EZC_CUDA(P.need.reserve(sizeof(int) * ((size_t)nC + 1)));
    CB_LAUNCH(cb_fc_init_kernel, nC, nC, P.fc_comps.as<FcComp>());
    CB_LAUNCH_CH(cb_fc_points16_kernel<0>, W, H, n_act, fg, P.labels.as<int>(), P.flags.as<int>(), W, H, list, vec, P.fc_comps.as<FcComp>(), nullptr);
    CB_LAUNCH(cb_fc_prefilter_kernel, nC, P.fc_comps.as<FcComp>(), nC, W);
    CB_LAUNCH(cb_fc_need_kernel, nC, P.fc_comps.as<FcComp>(), nC, P.need.as<int>());
    if (ezc_status st = exclusive_scan_i32(ctx, P.need.as<int>(), P.need.as<int>(), nC, P.scan_tmp, d_tot)) return st;
    EZC_CUDA(cudaMemcpyAsync(P.h.p, d_tot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    const int nPts = *P.h.as<int>();
    if (nPts <= 0) continue;
    EZC_CUDA(P.pool.reserve(sizeof(cb::Pt) * ((size_t)nPts + 16)));
    CB_LAUNCH(cb_fc_offset_kernel, nC, P.fc_comps.as<FcComp>(), nC, P.need.as<int>());
    CB_LAUNCH_CH(cb_fc_points16_kernel<1>, W, H, n_act, fg, P.labels.as<int>(), P.flags.as<int>(), W, H, list, vec, P.fc_comps.as<FcComp>(), P.pool.as<cb::Pt>());
    cb_fc_minrect_kernel<<<nC, kFcThreads, 0, ctx->stream>>>(P.fc_comps.as<FcComp>(), nC, P.pool.as<cb::Pt>(), N, cls, (W < 65536 && H < 65536) ? 1 : 0,
                                                             P.fc_hyps.as<FcHyp>(), P.fc_count.as<int>(), max_hyps);
    EZC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  EZC_CUDA(cudaMemcpyAsync(P.h.p, P.fc_count.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int nHyp = *P.h.as<int>();
  std::vector<int> next;
  if (nHyp > max_hyps) {
    // hypothesis buffer overflow (pathological noise): let everything through to the full search
    for (int im : alist) pass[im] = 1;
    alist.clear();
    return EZC_OK;
  }
  EZC_CUDA(P.fc_h.reserve(sizeof(FcHyp) * (size_t)std::max(nHyp, 1)));
  if (nHyp > 0) {
    EZC_CUDA(cudaMemcpyAsync(P.fc_h.p, P.fc_hyps.p, sizeof(FcHyp) * (size_t)nHyp, cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  const FcHyp* hh = P.fc_h.as<FcHyp>();
  std::vector<std::vector<std::pair<float, int>>> per(pass.size());
  for (int i = 0; i < nHyp; ++i) per[hh[i].slot_class >> 1].emplace_back(hh[i].size, hh[i].slot_class & 1);
  const bool dbg = getenv("EZC_FC_DEBUG") != nullptr;
  for (int im : alist) {
    int n0 = 0, n1 = 0;
    if (dbg) for (auto& q : per[im]) (q.second ? n1 : n0)++;
    const bool ok = fc_check_quads(per[im], pw, ph);
    if (dbg) {
      fprintf(stderr, "[fc] image %d thresholds (%d,%d): %d black + %d white hypotheses -> %s |", im, white_thresh, black_thresh, n0, n1, ok ? "pass" : "fail");
      for (size_t i = 0; i < per[im].size() && i < 60; ++i) fprintf(stderr, " %.1f%c", per[im][i].first, per[im][i].second ? 'w' : 'b');
      fprintf(stderr, "\n");
    }
    if (ok) pass[im] = 1; else next.push_back(im);
  }
  alist.swap(next);
  return EZC_OK;
}

// The CALIB_CB_FAST_CHECK gate of cv::findChessboardCorners: images for which neither checkChessboardBinary (on the
// histogram-binarised image B) nor checkChessboard (on the gray image) sees enough similar blobs leave `list`.
static ezc_status fast_check_gate(ezc_context* ctx, CbPools& P, const uint8_t* d_img, long long pitch, long long istride, const uint8_t* B,
                                  int W, int H, int nb, int pw, int ph, std::vector<int>& list) {
  const long long N = (long long)W * H, total = N * nb;
  const bool vec = (W % 16) == 0;
  const bool vec_src = vec && pitch % 16 == 0 && istride % 16 == 0 && ((uintptr_t)d_img & 15) == 0;
  EZC_CUDA(P.fc_w0.reserve(total)); EZC_CUDA(P.fc_w1.reserve(total)); EZC_CUDA(P.fc_w2.reserve(total));
  EZC_CUDA(P.fc_k0.reserve(total)); EZC_CUDA(P.fc_k1.reserve(total)); EZC_CUDA(P.fc_k2.reserve(total));
  EZC_CUDA(P.fc_fg.reserve(total));
  std::vector<char> pass(nb, 0);
  std::vector<int> alist = list;
  const uint8_t* white = B;
  const uint8_t* black = B;
  uint8_t* wbuf[2] = {P.fc_w0.as<uint8_t>(), P.fc_w1.as<uint8_t>()};   // reused by the gray-level check afterwards
  uint8_t* kbuf[2] = {P.fc_k0.as<uint8_t>(), P.fc_k1.as<uint8_t>()};
  // checkChessboardBinary passes if ANY of its four erosion levels passes, so the order of evaluation is free: level 1
  // (touching squares just separated) is tried first, it is where almost every real board passes; level 0 (squares still
  // joined at their corners: a handful of huge blobs) is the most expensive and the least likely.  Erosions are cumulative,
  // so the images of levels 1..3 are built in order and kept; level 0 is the binarised image itself.
  uint8_t* wl[4] = {nullptr, P.fc_w0.as<uint8_t>(), P.fc_w1.as<uint8_t>(), P.fc_w2.as<uint8_t>()};
  uint8_t* kl[4] = {nullptr, P.fc_k0.as<uint8_t>(), P.fc_k1.as<uint8_t>(), P.fc_k2.as<uint8_t>()};
  const int order[4] = {1, 2, 0, 3};
  int built = 0;
  for (int oi = 0; oi < 4 && !alist.empty(); ++oi) {
    const int e = order[oi];
    for (; built < e; ++built) {   // build level built+1 from level built for the images still undecided
      const int n_act = (int)alist.size();
      if (ezc_status st = upload_list(ctx, P.fc_list, alist)) return st;
      const uint8_t* ws = built == 0 ? B : wl[built];
      const uint8_t* ks = built == 0 ? B : kl[built];
      CB_LAUNCH_CH(cb_morph16_kernel<false>, W, H, n_act, ws, (long long)W, N, vec, wl[built + 1], W, H, P.fc_list.as<int>(), vec);
      CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, n_act, ks, (long long)W, N, vec, kl[built + 1], W, H, P.fc_list.as<int>(), vec);
    }
    white = e == 0 ? B : wl[e];
    black = e == 0 ? B : kl[e];
    if (ezc_status st = fc_round(ctx, P, white, black, 128, 128, W, H, pw, ph, alist, pass)) return st;
  }
  if (!alist.empty()) {
    const int n_act = (int)alist.size();
    if (ezc_status st = upload_list(ctx, P.fc_list, alist)) return st;
    CB_LAUNCH_CH(cb_morph16_kernel<false>, W, H, n_act, d_img, pitch, istride, vec_src, wbuf[0], W, H, P.fc_list.as<int>(), vec);
    CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, n_act, d_img, pitch, istride, vec_src, kbuf[0], W, H, P.fc_list.as<int>(), vec);
    for (int t = 20; t < 130 && !alist.empty(); t += 20)
      if (ezc_status st = fc_round(ctx, P, wbuf[0], kbuf[0], t + 70, t, W, H, pw, ph, alist, pass)) return st;
  }
  std::vector<int> keep;
  for (int im : list) if (pass[im]) keep.push_back(im);
  list.swap(keep);
  return EZC_OK;
}

ezc_status detect_sub_batch(ezc_context* ctx, const ezc_images* imgs, int img0, int nb, int pw, int ph, bool final_subpix, bool fast_check,
                            float2* d_corners_out, uint8_t* d_found_out) {
  CbPools& P = cb_pools(ctx);
  const int W = imgs->width, H = imgs->height;
  const long long N = (long long)W * H, total = N * nb;
  const uint8_t* d_img = imgs->d + (long long)img0 * imgs->image_stride;
  EZC_CUDA(images_wait(ctx, imgs, img0 + nb - 1));
  const int npts = pw * ph;
  EZC_CUDA(P.hist.reserve(sizeof(int) * 256 * nb)); EZC_CUDA(P.thresh.reserve(sizeof(int) * nb));
  EZC_CUDA(P.bin0.reserve(total)); EZC_CUDA(P.bin1.reserve(total));
  EZC_CUDA(P.labels.reserve(sizeof(int) * total)); EZC_CUDA(P.flags.reserve(sizeof(int) * (total + 1)));
  EZC_CUDA(P.total.reserve(64)); EZC_CUDA(P.h.reserve(4096));
  EZC_CUDA(P.works.reserve(sizeof(cb::Work) * (size_t)nb)); EZC_CUDA(P.scratch.reserve(sizeof(int) * cb::kMaxQuads * 4 * (size_t)nb));
  EZC_CUDA(P.state.reserve(sizeof(ImageState) * nb));
  EZC_CUDA(P.pcnt.reserve(sizeof(int) * total)); EZC_CUDA(P.plast.reserve(sizeof(int) * total));
  EZC_CUDA(cudaMemsetAsync(P.pcnt.p, 0, sizeof(int) * total, ctx->stream)); EZC_CUDA(cudaMemsetAsync(P.plast.p, 0, sizeof(int) * total, ctx->stream));
  EZC_CUDA(P.corners.reserve(sizeof(cb::Ptf) * (size_t)npts * nb)); EZC_CUDA(P.found.reserve(nb));
  EZC_CUDA(P.holes.reserve(sizeof(HoleInfo) * 16));
  std::vector<ImageState> h_state(nb, ImageState{0, 0, 0, 0, 0});
  std::vector<int> list(nb);
  for (int i = 0; i < nb; ++i) list[i] = i;
  EZC_CUDA(cudaMemcpyAsync(P.state.p, h_state.data(), sizeof(ImageState) * nb, cudaMemcpyHostToDevice, ctx->stream));
  if (ezc_status st = upload_list(ctx, P.active, list)) return st;
  const bool vec = (W % 16) == 0;   // packed working images: rows and images start 16-byte aligned
  const bool vec_src = imgs->pitch % 16 == 0 && imgs->image_stride % 16 == 0 && ((uintptr_t)d_img & 15) == 0 && vec;
  EZC_CUDA(cudaMemsetAsync(P.hist.p, 0, sizeof(int) * 256 * nb, ctx->stream));
  {
    dim3 grid(std::max(1, std::min(64, (int)((N + 255) / 256))), nb);
    cb_hist_kernel<<<grid, 256, 0, ctx->stream>>>(d_img, imgs->pitch, imgs->image_stride, vec_src, W, H, P.hist.as<int>());
    EZC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  CB_LAUNCH(cb_select_threshold_kernel, nb, P.hist.as<int>(), nb, W, H, P.thresh.as<int>());
  uint8_t* b0 = P.bin0.as<uint8_t>();
  uint8_t* b1 = P.bin1.as<uint8_t>();
  CB_LAUNCH_CH(cb_binarize16_kernel, W, H, nb, d_img, (long long)imgs->pitch, (long long)imgs->image_stride, vec_src, W, H, P.thresh.as<int>(),
               (const int*)nullptr, b0, vec);
  if (fast_check) {
    if (ezc_status st = fast_check_gate(ctx, P, d_img, imgs->pitch, imgs->image_stride, b0, W, H, nb, pw, ph, list)) return st;
    if (ezc_status st = upload_list(ctx, P.active, list)) return st;
  }
  std::vector<int> h_dil(nb, 0);
  EZC_CUDA(P.dil.reserve(sizeof(int) * nb));
  auto upload_dil = [&](int d) -> ezc_status {
    std::fill(h_dil.begin(), h_dil.end(), d);
    EZC_CUDA(cudaMemcpyAsync(P.dil.p, h_dil.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
    return EZC_OK;
  };
  bool phase2_ran = false;
  std::vector<int> h_phase2(nb, 0);
  // The search is a sequence of 8 + 48 attempts per view (threshold image x dilation count) that stops at the first
  // success.  Most views stop at attempt 0; the few that go on used to walk the rest one attempt at a time -- a serial
  // tail of up to 55 x ~2.7 ms with one view on the GPU.  The only state an attempt inherits is prev_sqr_size, and only
  // through the adaptive-threshold block size, so the remaining attempts of the unfinished views run CONCURRENTLY in
  // the slots the finished views left free, each on the block size its predecessors are expected to leave behind; the
  // host then walks the attempts in order and keeps exactly what the one-at-a-time search would have kept: an attempt
  // whose assumed block size turns out wrong ends the walk and is rerun in the next round.  EZC_CB_SEQUENTIAL=1 keeps
  // the one-at-a-time search (A/B parity check in tests/test_chessboard_gpu.py).
  if (getenv("EZC_CB_SEQUENTIAL")) {
    // ---- phase 1: histogram-binarised image; attempt 0 runs on it undilated, attempts 1..7 add one dilation each (the
    // frame of the previous pass is dilated too). cv2 4.13's first failing level on pre-dilated ideal boards pins this.
    for (int dil = 0; dil <= 7 && !list.empty(); ++dil) {
      const int n_act = (int)list.size();
      if (getenv("EZC_CB_TIMING")) fprintf(stderr, "[cb] phase 1 attempt %d: %d of %d views\n", dil, n_act, nb);
      if (dil > 0) {
        CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, n_act, b0, (long long)W, N, vec, b1, W, H, P.active.as<int>(), vec);
        std::swap(b0, b1);
      }
      CB_LAUNCH_FRAME(b0, n_act, P.active.as<int>());
      if (ezc_status st = upload_dil(dil)) return st;
      if (ezc_status st = run_attempt(ctx, P, b0, W, H, n_act, pw, ph, P.dil.as<int>())) return st;
      if (ezc_status st = sync_state(ctx, P, nb, h_state, list)) return st;
    }
    // ---- phase 2: equalised image, 6 adaptive thresholds x 8 dilations, for the images still without a board
    phase2_ran = !list.empty();
    for (int im : list) h_phase2[im] = 1;
    if (phase2_ran) {
      EZC_CUDA(P.eq.reserve(total)); EZC_CUDA(P.lut.reserve(256 * nb)); EZC_CUDA(P.prevbin.reserve(total));
      EZC_CUDA(P.block.reserve(sizeof(int) * nb)); EZC_CUDA(P.delta.reserve(sizeof(int) * nb)); EZC_CUDA(P.hsum.reserve(sizeof(int) * total));
      CB_LAUNCH(cb_equalize_lut_kernel, nb, P.hist.as<int>(), nb, (int)N, P.lut.as<uint8_t>());
      CB_LAUNCH_CH(cb_apply_lut_kernel, W, H, (int)list.size(), d_img, (long long)imgs->pitch, (long long)imgs->image_stride, vec_src, W, H, P.active.as<int>(),
                   P.lut.as<uint8_t>(), P.eq.as<uint8_t>(), vec);   // only the images that enter the fallback
      std::vector<int> h_block(nb, 0), h_delta(nb, 0), prev_block(nb, -1);
      for (int i = 0; i < nb; ++i) h_state[i].prev_sqr_size = 0;  // prev_sqr_size restarts at 0 with the fallback
      EZC_CUDA(cudaMemcpyAsync(P.state.p, h_state.data(), sizeof(ImageState) * nb, cudaMemcpyHostToDevice, ctx->stream));
      uint8_t* th = P.bin0.as<uint8_t>();       // thresh_img (gets the frame)
      uint8_t* tmp = P.bin1.as<uint8_t>();
      uint8_t* prevb = P.prevbin.as<uint8_t>(); // prev_thresh_img (no frame)
      for (int k = 0; k < 6 && !list.empty(); ++k) {
        std::fill(prev_block.begin(), prev_block.end(), -1);
        for (int dil = 0; dil <= 7 && !list.empty(); ++dil) {
          std::vector<int> fresh, old;
          for (int im : list) {
            const int ps = h_state[im].prev_sqr_size;
            int bs = (int)lrint(ps == 0 ? std::min(W, H) * (k % 2 == 0 ? 0.2 : 0.1) : ps * 2.0);
            bs |= 1;
            if (bs != prev_block[im]) fresh.push_back(im); else old.push_back(im);
            h_block[im] = bs;
            h_delta[im] = (k / 2) * 5;
            prev_block[im] = bs;
          }
          EZC_CUDA(cudaMemcpyAsync(P.block.p, h_block.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
          EZC_CUDA(cudaMemcpyAsync(P.delta.p, h_delta.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
          if (!fresh.empty()) {
            // new block size: adaptive threshold of the equalised image, then `dil` dilations -> prev_thresh_img
            if (ezc_status st = upload_list(ctx, P.refresh, fresh)) return st;
            const int nf = (int)fresh.size();
            if (ezc_status st = adaptive_threshold(ctx, P, W, H, nf, P.refresh.as<int>(), tmp, nullptr)) return st;
            for (int d = 0; d < dil; ++d) {
              CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, nf, tmp, (long long)W, N, vec, th, W, H, P.refresh.as<int>(), vec);
              CB_LAUNCH_PIX(cb_copy_active_kernel, (N + 15) / 16, nf, th, tmp, N, nf, P.refresh.as<int>(), (const int*)nullptr);
            }
            CB_LAUNCH_PIX(cb_copy_active_kernel, (N + 15) / 16, nf, tmp, prevb, N, nf, P.refresh.as<int>(), (const int*)nullptr);
          }
          if (!old.empty() && dil > 0) {
            // same block size as before: one more dilation of prev_thresh_img
            if (ezc_status st = upload_list(ctx, P.refresh, old)) return st;
            const int no = (int)old.size();
            CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, no, prevb, (long long)W, N, vec, tmp, W, H, P.refresh.as<int>(), vec);
            CB_LAUNCH_PIX(cb_copy_active_kernel, (N + 15) / 16, no, tmp, prevb, N, no, P.refresh.as<int>(), (const int*)nullptr);
          }
          const int n_act = (int)list.size();
          if (getenv("EZC_CB_TIMING")) fprintf(stderr, "[cb] phase 2 k %d dil %d: %d of %d views\n", k, dil, n_act, nb);
          CB_LAUNCH_PIX(cb_copy_active_kernel, (N + 15) / 16, n_act, prevb, th, N, n_act, P.active.as<int>(), (const int*)nullptr);
          CB_LAUNCH_FRAME(th, n_act, P.active.as<int>());
          // generateQuads gets the dilation count in the fallback too (edge_len compensation (sqrt(e) + 2 dil)^2), as OpenCV's
          // source reads.  Round 1 ran the fallback without it -- indistinguishable on every fixture it had; the randomised
          // degradations of tests/cb_fuzz.py separate the two: cv2 finds strongly blurred boards at 6-7 dilations only with it.
          if (ezc_status st = upload_dil(dil)) return st;
          if (ezc_status st = run_attempt(ctx, P, th, W, H, n_act, pw, ph, P.dil.as<int>())) return st;
          if (ezc_status st = sync_state(ctx, P, nb, h_state, list)) return st;
        }
      }
    }
  } else {
    // ---- attempt 0 for every view that passed the gate
    if (!list.empty()) {
      const int n_act = (int)list.size();
      if (getenv("EZC_CB_TIMING")) fprintf(stderr, "[cb] attempt 0: %d of %d views\n", n_act, nb);
      CB_LAUNCH_FRAME(b0, n_act, P.active.as<int>());
      if (ezc_status st = upload_dil(0)) return st;
      if (ezc_status st = run_attempt(ctx, P, b0, W, H, n_act, pw, ph, P.dil.as<int>())) return st;
      if (ezc_status st = sync_state(ctx, P, nb, h_state, list)) return st;
    }
    if (!list.empty()) {
      // ---- the other 7 + 48 attempts, concurrently
      struct Att { int phase, k, dil; };
      std::vector<Att> att;
      for (int d = 1; d <= 7; ++d) att.push_back({1, 0, d});
      const int kP2 = (int)att.size();   // index of the first attempt of the fallback
      for (int k = 0; k < 6; ++k) for (int d = 0; d <= 7; ++d) att.push_back({2, k, d});
      const int kAtt = (int)att.size();
      auto block_size = [&](int k, int ps) { return (int)lrint(ps == 0 ? std::min(W, H) * (k % 2 == 0 ? 0.2 : 0.1) : ps * 2.0) | 1; };
      struct Slot { int view, j, block, head; };
      std::vector<Slot> slots;
      std::vector<int> pos(nb, 0), ps(nb, 0), h_src(nb, 0), h_block(nb, 0), h_delta(nb, 0), pairs;
      std::vector<int> l_p1, l_head, l_follow, l_all, l_step;
      EZC_CUDA(P.eq.reserve(total)); EZC_CUDA(P.lut.reserve(256 * nb)); EZC_CUDA(P.prevbin.reserve(total));
      EZC_CUDA(P.block.reserve(sizeof(int) * nb)); EZC_CUDA(P.delta.reserve(sizeof(int) * nb)); EZC_CUDA(P.hsum.reserve(sizeof(int) * total));
      EZC_CUDA(P.srcmap.reserve(sizeof(int) * nb)); EZC_CUDA(P.refresh.reserve(sizeof(int) * nb)); EZC_CUDA(P.pairs.reserve(sizeof(int) * 2 * nb));
      EZC_CUDA(P.res_found.reserve(sizeof(int) * nb)); EZC_CUDA(P.res_corners.reserve(sizeof(cb::Ptf) * (size_t)npts * nb));
      CB_LAUNCH(cb_save_results_kernel, (long long)nb * npts, P.state.as<ImageState>(), P.corners.as<cb::Ptf>(), nb, npts, P.res_found.as<int>(),
                P.res_corners.as<cb::Ptf>());
      // equalised copies (CALIB_CB_NORMALIZE_IMAGE) are made when a view's first fallback attempt is scheduled
      std::vector<char> has_eq(nb, 0);
      std::vector<int> l_eq;
      bool lut_ready = false;
      const uint8_t* base1 = b0;                   // attempt 0's image with its frame: phase 1 dilates it d times
      uint8_t* tmp = b1;
      uint8_t* th = P.prevbin.as<uint8_t>();
      while (!list.empty()) {
        const int R = (int)list.size();
        // Attempts per view and round.  A slot is a whole image of per-pixel and per-contour work, so what is run ahead
        // should be likely to be needed: views that fail attempts 0 and 1 mostly succeed a dilation or two later (cheap,
        // clean images: the rest of phase 1 goes in one round, never across the phase boundary), and a view that reaches
        // the fallback mostly fails all 48 of its attempts (noise: its 6 undilated thresholds cost ~15 ms each, the rest
        // ~1.5 ms -- side by side they cost one of the former).  EZC_CB_SLOTS caps the slots of a round.
        const int slot_cap = std::min(nb, getenv("EZC_CB_SLOTS") ? std::max(1, atoi(getenv("EZC_CB_SLOTS"))) : 64);
        int c = std::max(1, slot_cap / R);
        slots.clear(); l_p1.clear(); l_head.clear(); l_follow.clear(); l_all.clear();
        for (int v : list) {
          int assumed = ps[v], head = -1, head_k = -1;
          // phase 1: dilations 2-3 first (where most late successes are), then the rest up to the phase boundary
          const int j_end = pos[v] < kP2 ? std::min(pos[v] + std::min(c, pos[v] < 3 ? 2 : kP2), kP2) : std::min(pos[v] + c, kAtt);
          for (int j = pos[v]; j < j_end; ++j) {
            const Att& a = att[j];
            const int sl = (int)slots.size();
            Slot S{v, j, 0, -1};
            h_dil[sl] = a.dil;
            if (a.phase == 1) {
              h_src[sl] = v; h_block[sl] = 0; h_delta[sl] = 0;
              l_p1.push_back(sl);
            } else {
              if (j == kP2) assumed = 0;   // prev_sqr_size restarts at 0 with the fallback
              S.block = block_size(a.k, assumed);
              h_block[sl] = S.block; h_delta[sl] = (a.k / 2) * 5;
              if (head >= 0 && head_k == a.k) { S.head = head; h_src[sl] = head; l_follow.push_back(sl); }
              else { head = sl; head_k = a.k; h_src[sl] = v; l_head.push_back(sl); }
            }
            l_all.push_back(sl);
            slots.push_back(S);
          }
        }
        const int n_slots = (int)slots.size();
        l_eq.clear();
        for (int sl : l_head) if (!has_eq[slots[sl].view]) { has_eq[slots[sl].view] = 1; l_eq.push_back(slots[sl].view); }
        if (!l_eq.empty()) {
          if (!lut_ready) { CB_LAUNCH(cb_equalize_lut_kernel, nb, P.hist.as<int>(), nb, (int)N, P.lut.as<uint8_t>()); lut_ready = true; }
          if (ezc_status st = upload_list(ctx, P.refresh, l_eq)) return st;
          CB_LAUNCH_CH(cb_apply_lut_kernel, W, H, (int)l_eq.size(), d_img, (long long)imgs->pitch, (long long)imgs->image_stride, vec_src, W, H,
                       P.refresh.as<int>(), P.lut.as<uint8_t>(), P.eq.as<uint8_t>(), vec);
        }
        const bool timing = getenv("EZC_CB_TIMING") != nullptr;
        const auto t_round = std::chrono::steady_clock::now();
        auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
          cudaStreamSynchronize(ctx->stream);
          return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        };
        EZC_CUDA(cudaMemcpyAsync(P.srcmap.p, h_src.data(), sizeof(int) * n_slots, cudaMemcpyHostToDevice, ctx->stream));
        EZC_CUDA(cudaMemcpyAsync(P.block.p, h_block.data(), sizeof(int) * n_slots, cudaMemcpyHostToDevice, ctx->stream));
        EZC_CUDA(cudaMemcpyAsync(P.delta.p, h_delta.data(), sizeof(int) * n_slots, cudaMemcpyHostToDevice, ctx->stream));
        EZC_CUDA(cudaMemcpyAsync(P.dil.p, h_dil.data(), sizeof(int) * n_slots, cudaMemcpyHostToDevice, ctx->stream));
        const int* d_src = P.srcmap.as<int>();
        // Slot images.  Fallback slots start from the adaptive threshold of their view (one per (view, k): the other
        // dilation levels of that k copy it) in buffer A; phase-1 slots start from attempt 0's framed image.  Dilation
        // step t takes every slot with >= t dilations from one buffer to the other (A -> B for odd t), so a slot ends in
        // B when its dilation count is odd; the others are copied over once.  B then gets the frames.
        uint8_t* A = tmp;
        uint8_t* B = th;
        if (!l_head.empty()) {
          if (ezc_status st = upload_list(ctx, P.refresh, l_head)) return st;
          if (ezc_status st = adaptive_threshold(ctx, P, W, H, (int)l_head.size(), P.refresh.as<int>(), A, d_src)) return st;
          // a head that is not at dilation 0 (the walk resumed in the middle of a k) is dilated below like the others
        }
        if (!l_follow.empty()) {
          if (ezc_status st = upload_list(ctx, P.refresh, l_follow)) return st;
          CB_LAUNCH_PIX(cb_copy_active_kernel, (N + 15) / 16, (int)l_follow.size(), A, A, N, (int)l_follow.size(), P.refresh.as<int>(), d_src);
        }
        if (!l_p1.empty()) {   // every phase-1 slot has >= 1 dilation: its first step reads the view's image directly
          if (ezc_status st = upload_list(ctx, P.refresh, l_p1)) return st;
          CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, (int)l_p1.size(), base1, (long long)W, N, vec, B, W, H, P.refresh.as<int>(), vec, d_src);
        }
        for (int t = 1; t <= 7; ++t) {
          l_step.clear();
          for (int sl = 0; sl < n_slots; ++sl) if (h_dil[sl] >= t && !(t == 1 && att[slots[sl].j].phase == 1)) l_step.push_back(sl);
          if (l_step.empty()) continue;
          if (ezc_status st = upload_list(ctx, P.refresh, l_step)) return st;
          const int ns = (int)l_step.size();
          const uint8_t* from = (t & 1) ? A : B;
          uint8_t* to = (t & 1) ? B : A;
          CB_LAUNCH_CH(cb_morph16_kernel<true>, W, H, ns, from, (long long)W, N, vec, to, W, H, P.refresh.as<int>(), vec);
        }
        l_step.clear();
        for (int sl = 0; sl < n_slots; ++sl) if ((h_dil[sl] & 1) == 0) l_step.push_back(sl);
        if (!l_step.empty()) {
          if (ezc_status st = upload_list(ctx, P.refresh, l_step)) return st;
          CB_LAUNCH_PIX(cb_copy_active_kernel, (N + 15) / 16, (int)l_step.size(), A, B, N, (int)l_step.size(), P.refresh.as<int>(), (const int*)nullptr);
        }
        if (ezc_status st = upload_list(ctx, P.active, l_all)) return st;
        CB_LAUNCH_FRAME(B, n_slots, P.active.as<int>());
        for (int sl = 0; sl < n_slots; ++sl) { h_state[sl] = ImageState{0, 0, 0, 0, 0}; h_state[sl].prev_sqr_size = -1; }   // -1: not written by this attempt
        EZC_CUDA(cudaMemcpyAsync(P.state.p, h_state.data(), sizeof(ImageState) * n_slots, cudaMemcpyHostToDevice, ctx->stream));
        const double ms_prep = timing ? ms_since(t_round) : 0.0;
        if (ezc_status st = run_attempt(ctx, P, B, W, H, n_slots, pw, ph, P.dil.as<int>())) return st;
        EZC_CUDA(cudaMemcpyAsync(h_state.data(), P.state.p, sizeof(ImageState) * n_slots, cudaMemcpyDeviceToHost, ctx->stream));
        EZC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (timing)
          fprintf(stderr, "[cb] concurrent round: %d views x up to %d attempts = %d slots of %d: images %.2f ms, attempts %.2f ms\n", R, c, n_slots, nb,
                  ms_prep, ms_since(t_round) - ms_prep);
        // walk every view's attempts in order
        std::vector<int> next;
        pairs.clear();
        for (int sl = 0; sl < n_slots;) {
          const int v = slots[sl].view;
          int cur = ps[v], j_next = pos[v];
          bool done = false;
          int e = sl;
          for (; e < n_slots && slots[e].view == v; ++e) {
            if (done) continue;
            const Att& a = att[slots[e].j];
            if (a.phase == 2) {
              h_phase2[v] = 1;
              if (slots[e].j == kP2) cur = 0;
              if (block_size(a.k, cur) != slots[e].block) { done = true; continue; }   // ran on the wrong threshold image: rerun from here
            }
            if (h_state[e].found) { pairs.push_back(e); pairs.push_back(v); j_next = kAtt; done = true; continue; }
            if (h_state[e].prev_sqr_size >= 0) cur = h_state[e].prev_sqr_size;
            j_next = slots[e].j + 1;
          }
          // a walk that stops exactly at the phase boundary re-enters with the fallback's reset applied by the next round
          pos[v] = j_next; ps[v] = cur;
          if (j_next < kAtt) next.push_back(v);
          sl = e;
        }
        if (!pairs.empty()) {
          EZC_CUDA(cudaMemcpyAsync(P.pairs.p, pairs.data(), sizeof(int) * pairs.size(), cudaMemcpyHostToDevice, ctx->stream));
          CB_LAUNCH(cb_accept_kernel, (long long)(pairs.size() / 2) * npts, P.pairs.as<int>(), (int)(pairs.size() / 2), npts, P.corners.as<cb::Ptf>(),
                    P.res_found.as<int>(), P.res_corners.as<cb::Ptf>());
        }
        list.swap(next);
      }
      for (int i = 0; i < nb; ++i) phase2_ran = phase2_ran || h_phase2[i] != 0;
      CB_LAUNCH(cb_restore_results_kernel, (long long)nb * npts, P.state.as<ImageState>(), P.corners.as<cb::Ptf>(), nb, npts, P.res_found.as<int>(),
                P.res_corners.as<cb::Ptf>());
    }
  }
  // ---- final checks, internal (2,2) refinement on the image the search ended on, then the reference's (5,5) refinement
  CB_LAUNCH(cb_finalize_kernel, nb, P.state.as<ImageState>(), nb, pw, ph, W, H, P.corners.as<cb::Ptf>(), P.found.as<uint8_t>());
  EZC_CUDA(P.sub_corners.reserve(sizeof(float2) * (size_t)npts * nb)); EZC_CUDA(P.sub_idx.reserve(sizeof(int) * (size_t)npts * nb));
  CB_LAUNCH(cb_gather_corners_kernel, (long long)nb * npts, P.corners.as<cb::Ptf>(), P.found.as<uint8_t>(), nb, npts, P.sub_corners.as<float2>(),
            P.sub_idx.as<int>());
  // cv::cornerSubPix(img, corners, (2,2), (-1,-1), {EPS+MAX_ITER, 15, 0.1}): `img` is the original image when the board was
  // found on the histogram-binarised image, the equalised copy when the fallback had to run (NORMALIZE_IMAGE).
  ezc_images view;
  view.ctx = ctx; view.n = nb; view.width = W; view.height = H;
  if (!phase2_ran) {
    view.d = d_img; view.pitch = imgs->pitch; view.image_stride = imgs->image_stride;
  } else {
    // dense "used" image: equalised for the images that entered the fallback, original otherwise (LUT = identity)
    std::vector<uint8_t> ident(256 * (size_t)nb);
    for (int i = 0; i < nb; ++i) for (int k = 0; k < 256; ++k) ident[(size_t)i * 256 + k] = (uint8_t)k;
    std::vector<uint8_t> h_lut(256 * (size_t)nb);
    EZC_CUDA(cudaMemcpyAsync(h_lut.data(), P.lut.p, 256 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nb; ++i) if (!h_phase2[i]) std::copy(&ident[(size_t)i * 256], &ident[(size_t)i * 256] + 256, &h_lut[(size_t)i * 256]);
    EZC_CUDA(cudaMemcpyAsync(P.lut.p, h_lut.data(), 256 * (size_t)nb, cudaMemcpyHostToDevice, ctx->stream));
    CB_LAUNCH_CH(cb_apply_lut_kernel, W, H, nb, d_img, (long long)imgs->pitch, (long long)imgs->image_stride, vec_src, W, H, (const int*)nullptr,
                 P.lut.as<uint8_t>(), P.eq.as<uint8_t>(), vec);
    view.d = P.eq.as<uint8_t>(); view.pitch = W; view.image_stride = N;
  }
  if (ezc_status st = subpix_device(ctx, &view, P.sub_corners.as<float2>(), P.sub_idx.as<int>(), nb * npts, 2, 2, 15, 0.1)) return st;
  if (final_subpix) {
    // chessboard_detector.hpp:39-40 -- on the ORIGINAL image
    view.d = d_img; view.pitch = imgs->pitch; view.image_stride = imgs->image_stride;
    if (ezc_status st = subpix_device(ctx, &view, P.sub_corners.as<float2>(), P.sub_idx.as<int>(), nb * npts, 5, 5, 100, 0.05)) return st;
  }
  CB_LAUNCH(cb_scatter_corners_kernel, (long long)nb * npts, P.sub_corners.as<float2>(), P.found.as<uint8_t>(), nb, npts,
            d_corners_out + (size_t)img0 * npts);
  EZC_CUDA(cudaMemcpyAsync(d_found_out + img0, P.found.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
  return EZC_OK;
}

}  // namespace

extern "C" {

ezc_status ezc_find_chessboard_corners_batch(ezc_context* ctx, const ezc_images* imgs, int32_t pattern_w, int32_t pattern_h,
                                             int32_t final_subpix, int32_t fast_check, float* corners_out, uint8_t* found_out,
                                             int32_t max_images_in_flight) {
  EZC_CHECK(ctx && imgs && corners_out && found_out, EZC_ERR_INVALID, "ezc_find_chessboard_corners_batch: null argument");
  EZC_CHECK(pattern_w > 2 && pattern_h > 2, EZC_ERR_INVALID, "ezc_find_chessboard_corners_batch: both pattern dimensions must be > 2 (OpenCV rule)");
  EZC_CHECK(pattern_w * pattern_h <= cb::kMaxQuads, EZC_ERR_INVALID, "ezc_find_chessboard_corners_batch: pattern too large");
  EZC_CHECK(imgs->width >= 16 && imgs->height >= 16, EZC_ERR_INVALID, "ezc_find_chessboard_corners_batch: image too small");
  EZC_CUDA(cudaSetDevice(ctx->device));
  const int n = imgs->n;
  if (n == 0) return EZC_OK;
  const long long N = (long long)imgs->width * imgs->height;
  // images in flight: ~32 B/px of per-pixel scratch (binary images, labels, component ids, parent counters, gate images),
  // bounded to 48 GiB and to half of what the device has free (180 GB HBM).  Large sub-batches matter: a view that passes
  // the gate and fails the search runs its 56 attempts in ~0.16 s while the rest of its sub-batch is long done, so the
  // fewer sub-batches the fewer serial tails.  (2^30 pixels per sub-batch keeps linear pixel indices in 32 bits.)
  size_t free_b = 0, total_b = 0;
  EZC_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const long long budget = std::min<long long>(48LL << 30, (long long)((free_b + ezc::dev_cache_bytes(ctx->device)) / 2));
  int sub = (int)std::max<long long>(1, std::min<long long>((1LL << 30) / N, budget / (N * 32)));
  sub = std::min(sub, 512);
  if (max_images_in_flight > 0) sub = std::min(sub, max_images_in_flight);
  sub = std::min(sub, n);
  const int npts = pattern_w * pattern_h;
  ezc::DevBuf d_corners, d_found;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  EZC_CUDA(d_corners.reserve(sizeof(float2) * (size_t)n * npts));
  EZC_CUDA(d_found.reserve((size_t)n));
  // No sub-batch ramp under ezc_images_upload_async here (unlike the AprilGrid path): a chessboard sub-batch costs ~10x its
  // upload time and every extra sub-batch risks another serial tail, so waiting for the upload of the first sub-batch is cheaper.
  for (int i0 = 0; i0 < n; i0 += sub) {
    const int nb = std::min(sub, n - i0);
    if (ezc_status st = detect_sub_batch(ctx, imgs, i0, nb, pattern_w, pattern_h, final_subpix != 0, fast_check != 0, d_corners.as<float2>(), d_found.as<uint8_t>()))
      return st;
  }
  EZC_CUDA(cudaMemcpyAsync(corners_out, d_corners.p, sizeof(float2) * (size_t)n * npts, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(found_out, d_found.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (getenv("EZC_CB_TIMING")) {
    unsigned long long t[6] = {0};
    cudaMemcpyFromSymbol(t, g_cb_ns, sizeof(t));
    fprintf(stderr, "[cb timing, cumulative ms] parent-count %.1f | +add_quad %.1f | neighbours %.1f | groups %.1f\n", t[0] * 1e-6, t[1] * 1e-6,
            t[2] * 1e-6, t[3] * 1e-6);
  }
  return EZC_OK;
}

// ChessboardCalibDetector::detectTarget for every image: nb_rows x nb_cols SQUARES (the reference's constructor
// arguments), i.e. pattern (nb_cols-1) x (nb_rows-1) inner corners, findChessboardCorners + cornerSubPix (5,5).
ezc_status ezc_detect_chessboard_batch(ezc_context* ctx, const ezc_images* imgs, int32_t nb_rows, int32_t nb_cols, float* corners_out,
                                       uint8_t* found_out, int32_t max_images_in_flight) {
  return ezc_find_chessboard_corners_batch(ctx, imgs, nb_cols - 1, nb_rows - 1, 1, 1, corners_out, found_out, max_images_in_flight);
}

}  // extern "C"
