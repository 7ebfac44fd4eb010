// apriltag.cu -- batched AprilGrid detection on the GPU (half (a), AprilGrid path).
//
// Replaces AprilTagCalibDetector::detectTarget (/root/reference/include/target_detectors/aprilgrid_detector.hpp:30-71)
// = AprilTags::TagDetector::extractTags (/root/reference/thirdparty/ethz_apriltag2/src/TagDetector.cc:40-586)
// + per-tag cv::cornerSubPix + scatter into 4*nb_tags corner slots.
//
// Not a port: the reference walks one image at a time through std::vector / std::map / pointer lists on one
// core.  Here a whole batch of images resident in HBM goes through data-parallel stages, and every stage is
// arranged so that its result is IDENTICAL to the reference's sequential semantics (float32 operation order
// included; this translation unit is compiled with -fmad=false and uses bit-exact libm restatements):
//   1  ONE tiled kernel: u8 -> float, 3-tap Gaussian blur (with the reference's border quirk) in shared memory, gradient
//      magnitude, activity bit mask (1 bit / pixel) and -- only for the ~5 % active pixels -- the bit-exact atan2
//   2  compaction of the active pixels from the scanned per-word popcounts (raster order == compact order)
//   3  edge emission in generation order (count, scan, scatter)
//   4  connected components of the edge graph (lock-free union-find); edges of different components never
//      interact in Edge::mergeEdges, so the reference's single sequential sweep over the cost-sorted edge
//      list factorises into one independent sequential sweep PER COMPONENT
//   5  stable radix sort of the edges by (component, cost) -- generation order inside equal keys
//   6  one WARP per component: the exact union-find / span-merge sweep of Edge.cc:69-117 (roots of the next 32 edges
//      resolved in parallel, only candidate merges replayed in order)
//   7  stable radix sort of the clustered pixels by representative (== std::map order, raster inside)
//   8  one thread per cluster: weighted line fit, segment orientation vote (float32 sums in raster order)
//   9  10-px spatial hash, child lists, depth-4 quad search, tag decoding, de-duplication, cornerSubPix
// Stages 4-6 are what removes the reference's serial bottleneck (~1e6 dependent union-find steps per image).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "images.cuh"
#include "libm_exact.cuh"
#include "scan_sort.cuh"
#include "tag36h11.cuh"

namespace ezc {

// ---- constants of the reference ------------------------------------------------------------
__device__ __constant__ float c_filt[3];           // Gaussian::makeGaussianFilter(0.8, 3), computed on the host
constexpr float kMinMag = 0.004f;                  // Edge::minMag          (Edge.cc:8)
constexpr float kThetaThresh = 100.f;              // Edge::thetaThresh     (Edge.cc:11)
constexpr float kMagThresh = 1200.f;               // Edge::magThresh       (Edge.cc:12)
constexpr int kMinSegmentSize = 4;                 // Segment::minimumSegmentSize (Segment.h:14)
constexpr float kMinLineLength = 4.f;              // Segment::minimumLineLength  (Segment.cc:6)
constexpr int kMinEdgeLength = 6;                  // Quad::minimumEdgeLength (Quad.h:22)
constexpr float kMaxQuadAspect = 32.f;             // Quad::maxQuadAspectRatio (Quad.cc:11)
constexpr float kPiF = 3.14159274101257324f;       // (float)M_PI

__device__ __forceinline__ float max_edge_cost() { return 30.f * kPiF / 180.f; }  // Edge.cc:9

// MathUtil::mod2pi (MathUtil.h:30-38)
__device__ __forceinline__ float mod2pi(float vin) {
  const float twopi = 2 * kPiF;
  const float twopi_inv = 1.f / (2.f * kPiF);
  const float absv = fabsf(vin);
  const float q = absv * twopi_inv + 0.5f;
  const int qi = (int)q;
  const float r = absv - qi * twopi;
  return (vin < 0) ? -r : r;
}
__device__ __forceinline__ float mod2pi_ref(float ref, float v) { return ref + mod2pi(v - ref); }
__device__ __forceinline__ float dist2d(float ax, float ay, float bx, float by) {
  const float dx = ax - bx, dy = ay - by;
  return sqrtf(dx * dx + dy * dy);
}
__device__ __forceinline__ float fminr(float a, float b) { return (b < a) ? b : a; }  // std::min
__device__ __forceinline__ float fmaxr(float a, float b) { return (a < b) ? b : a; }  // std::max

// ---- stage 1: preprocessing -----------------------------------------------------------------
// Horizontal pass of FloatImage::filterFactoredCentered (FloatImage.cc:46-52 -> Gaussian.cc:33-69).
// NOTE the reference compares ABSOLUTE offsets (aoff + i) against (alen + j): rows y >= 1 therefore get
// unfiltered border columns, and row 1 reads the last pixel of row 0.  Reproduced as is.
__global__ void at_blur_h_kernel(const uint8_t* __restrict__ img, long long pitch, long long istride, int W, int H,
                                 int n_img, float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long N = (long long)W * H;
  if (gid >= N * n_img) return;
  const int im = (int)(gid / N);
  const int p = (int)(gid - (long long)im * N);
  const int y = p / W, x = p - y * W;
  const uint8_t* I = img + (long long)im * istride;
  auto px = [&](int xx, int yy) -> float { return (float)((double)I[(long long)yy * pitch + xx] / 255.); };
  const float f0 = c_filt[0], f1 = c_filt[1], f2 = c_filt[2];
  double acc = 0;
  if (x >= 2 && x <= W - 2) {
    acc += (double)(px(x + 1, y) * f0); acc += (double)(px(x, y) * f1); acc += (double)(px(x - 1, y) * f2);
  } else if (y == 0) {
    if (x == 0) { acc += (double)(px(1, 0) * f0); acc += (double)(px(0, 0) * f1); acc += (double)(px(0, 0) * f2); }
    else if (x == 1) { acc += (double)(px(2, 0) * f0); acc += (double)(px(1, 0) * f1); acc += (double)(px(0, 0) * f2); }
    else { acc += (double)(px(W - 1, 0) * f0); acc += (double)(px(W - 1, 0) * f1); acc += (double)(px(W - 2, 0) * f2); }
  } else if (y == 1 && x == 0) {
    acc += (double)(px(0, 1) * f0); acc += (double)(px(0, 1) * f1); acc += (double)(px(W - 1, 0) * f2);
  } else {
    const float a = (x == W - 1) ? px(W - 1, y) : px(0, y);  // columns 0,1 replicate in(0,y); column W-1 in(W-1,y)
    acc += (double)(a * f0); acc += (double)(a * f1); acc += (double)(a * f2);
  }
  out[gid] = (float)acc;
}

// Vertical pass (true replicated border: the column is copied out with aoff = 0).
__global__ void at_blur_v_kernel(const float* __restrict__ r, int W, int H, int n_img, float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long N = (long long)W * H;
  if (gid >= N * n_img) return;
  const int im = (int)(gid / N);
  const int p = (int)(gid - (long long)im * N);
  const int y = p / W, x = p - y * W;
  const float* R = r + (long long)im * N;
  const float f0 = c_filt[0], f1 = c_filt[1], f2 = c_filt[2];
  const int yp = min(y + 1, H - 1), ym = max(y - 1, 0);
  double acc = 0;
  acc += (double)(R[(long long)yp * W + x] * f0);
  acc += (double)(R[(long long)y * W + x] * f1);
  acc += (double)(R[(long long)ym * W + x] * f2);
  out[gid] = (float)acc;
}

// Gradient (TagDetector.cc:155-169): interior pixels only; border stays 0.  Also emits the activity flag
// of stage 2: x < W-1, y < H-1, mag >= minMag (TagDetector.cc:214-219).
__global__ void at_gradient_kernel(const float* __restrict__ seg, int W, int H, int n_img, float* __restrict__ theta,
                                   float* __restrict__ mag, int* __restrict__ flag) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long N = (long long)W * H;
  if (gid >= N * n_img) return;
  const int im = (int)(gid / N);
  const int p = (int)(gid - (long long)im * N);
  const int y = p / W, x = p - y * W;
  float th = 0.f, mg = 0.f;
  if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
    const float* S = seg + (long long)im * N;
    const float Ix = S[p + 1] - S[p - 1];
    const float Iy = S[p + W] - S[p - W];
    mg = Ix * Ix + Iy * Iy;
    th = fd_atan2f(Iy, Ix);
  }
  theta[gid] = th;
  mag[gid] = mg;
  flag[gid] = (x + 1 < W && y + 1 < H && mg >= kMinMag) ? 1 : 0;
}

// test hook: the bit-exact libm restatements (libm_exact.cuh) on arbitrary arguments (tests/test_libm_exact.py)
__global__ void at_libm_kernel(int which, const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = which == 0 ? fd_atan2f(a[i], b[i]) : which == 1 ? gl_sinf(a[i]) : gl_cosf(a[i]);
}

// ---- stage 1, fused (production path) ------------------------------------------------------------
// One tile-based kernel does u8 -> float, both blur passes (incl. the border quirk), the gradient, the activity test and
// -- only for active pixels (~5 %) -- the bit-exact atan2f.  Same expressions, same evaluation order as the three
// kernels above (which remain as the debug/export path and as the cross-check of this one), but 1 B/px read and
// ~0.5 B/px written instead of ~25 B/px moved: the blurred images live in shared memory only.
//   out: amask[(im*H + y)*WPR + x/32]  bit x%32 = pixel active (mag >= minMag; implies 1 <= x <= W-2, 1 <= y <= H-2)
//        acount[...] = popc(amask)   (scanned -> compact ids in raster order)
//        theta/mag images, written ONLY at active pixels.
constexpr int kTX = 64, kTY = 32;               // output tile
constexpr int kUW = kTX + 4, kUH = kTY + 4;     // u8 tile incl. halo 2
constexpr int kHW = kTX + 2, kHH = kTY + 4;     // blur_h tile: cols x-1..x+TX, rows y-2..y+TY+1
constexpr int kSW = kTX + 2, kSH = kTY + 2;     // blurred tile: halo 1
__device__ __constant__ float c_u8f[256];       // (float)((double)v / 255.)   (FloatImage from cv::Mat, TagDetector.cc:131-139)

// Stage D of at_pre (shared by both tile flavours): gradient, activity mask, lazy theta.  Warp w owns rows 4w..4w+3, two
// 32-pixel words per row.
template <bool INTERIOR>
__device__ __forceinline__ void at_pre_gradient(const float (*sb)[kSW], int tid, int im, int x0, int y0, int W, int H, int wpr,
                                                uint32_t* __restrict__ amask, int* __restrict__ acount, float* __restrict__ theta,
                                                float* __restrict__ mag) {
  const int warp = tid >> 5, lane = tid & 31;
  const long long N = (long long)W * H;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const int r = warp * 4 + (k >> 1), cw = (k & 1) * 32;
    const int y = y0 + r, xw = x0 + cw;
    if (!INTERIOR && (y >= H || xw >= W)) continue;  // warp-uniform
    const int x = xw + lane, c = cw + lane;
    float mg = 0.f, Ix = 0.f, Iy = 0.f;
    if (INTERIOR || (x >= 1 && x < W - 1 && y >= 1 && y < H - 1)) {
      Ix = sb[r + 1][c + 2] - sb[r + 1][c];
      Iy = sb[r + 2][c + 1] - sb[r][c + 1];
      mg = Ix * Ix + Iy * Iy;
    }
    const bool act = mg >= kMinMag;
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (lane == 0) {
      const long long wi = ((long long)im * H + y) * wpr + (xw >> 5);
      amask[wi] = m;
      acount[wi] = __popc(m);
    }
    if (act) {
      // the gradient itself; theta = atan2f(Iy, Ix) and mag are taken per ACTIVE pixel in at_theta_kernel: evaluating the
      // ~250-instruction bit-exact atan2f here costs the whole warp whenever one lane of the word is active (~5 % of the
      // pixels, ~40 % of the words)
      const long long g = (long long)im * N + (long long)y * W + x;
      theta[g] = Ix;
      mag[g] = Iy;
    }
  }
}

__global__ void __launch_bounds__(256) at_pre_kernel(const uint8_t* __restrict__ img, long long pitch, long long istride, int W, int H,
                                                      int tiles_x, int tiles_y, int wpr, uint32_t* __restrict__ amask,
                                                      int* __restrict__ acount, float* __restrict__ theta, float* __restrict__ mag) {
  // The blur accumulates three float products in a double and rounds once to float.  Round 1's ncu showed this kernel issue
  // bound at 252 thread-instructions per pixel: border tests, clamps and 2-D index arithmetic around ~40 instructions of
  // real work.  Tiles that do not touch the image border (92 % at 2448x2048) now run a branch-free path: aligned 32-bit
  // pixel loads, flat constant-trip loops, and the horizontal pass -- whose inputs are BYTES -- takes its addends from
  // tables of the already widened products (double)(lut[v] * f_k) (no f32 -> f64 conversion).  The addends span < 2^11, so
  // the sum of three is exact in a double and the rounded float is the same for any summation order.
  __shared__ double lutd[3][256];
  __shared__ __align__(16) uint32_t uw_store[kUH * 19];   // u8 tile as words
  __shared__ float hb_store[kHH * (kHW + 1)];
  __shared__ float sb[kSH][kSW];
  uint32_t (*uw)[18] = reinterpret_cast<uint32_t (*)[18]>(uw_store);        // border tiles: byte b of row r = pixel x0 - 2 + b
  uint32_t (*uwi)[19] = reinterpret_cast<uint32_t (*)[19]>(uw_store);       // interior tiles: byte b of row r = pixel x0 - 4 + b
  float (*hb)[kHW] = reinterpret_cast<float (*)[kHW]>(hb_store);
  float (*hbi)[kHW + 1] = reinterpret_cast<float (*)[kHW + 1]>(hb_store);   // odd row stride for the row-per-lane pass B
  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int im = t / tiles_y;
  const int x0 = tx * kTX, y0 = ty * kTY;
  const uint8_t* I = img + (long long)im * istride;
  const float f0 = c_filt[0], f1 = c_filt[1], f2 = c_filt[2];
  {
    const float v = c_u8f[tid];
    lutd[0][tid] = (double)(v * f0);
    lutd[1][tid] = (double)(v * f1);
    lutd[2][tid] = (double)(v * f2);
  }
  const bool interior = x0 >= 4 && x0 + kTX + 4 <= W - 2 && y0 >= 2 && y0 + kTY + 2 <= H - 1 &&
                        ((reinterpret_cast<uintptr_t>(I) | (uintptr_t)pitch) & 3) == 0;
  if (interior) {
    // A. pixels: 36 rows x 18 aligned words (columns x0-4 .. x0+67), row stride 19 words (odd: lanes that differ in the row
    //    hit different banks in pass B)
    for (int i = tid; i < kUH * 18; i += 256) {
      const int r = i / 18, w = i - r * 18;
      uwi[r][w] = *reinterpret_cast<const uint32_t*>(I + (long long)(y0 - 2 + r) * pitch + (x0 - 4) + 4 * w);
    }
    __syncthreads();
    // B. horizontal pass, sliding window: task = (row r, segment of 11 outputs); the widened products of every input pixel are
    //    looked up once and reused by the outputs they feed.  h(r, c) at x = x0 - 1 + c  <-  bytes c+4 (x+1), c+3 (x), c+2 (x-1).
    if (tid < kHH * 6) {
      const int seg = tid / kHH, r = tid - seg * kHH;      // consecutive lanes = consecutive rows
      const int c0 = seg * 11;
      const uint8_t* urow = reinterpret_cast<const uint8_t*>(uwi[r]);
      double dm = lutd[2][urow[c0 + 2]];                    // f2 * p(x-1)
      int vc = urow[c0 + 3];
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const int vp = urow[c0 + k + 4];
        double acc = lutd[0][vp];                           // f0 * p(x+1)
        acc += lutd[1][vc];                                 // f1 * p(x)
        acc += dm;
        hbi[r][c0 + k] = (float)acc;
        dm = lutd[2][vc];
        vc = vp;
      }
    }
    __syncthreads();
    // C. vertical pass, sliding window: task = (column c, segment of rows); each h value is multiplied and widened once per
    //    tap instead of once per use.  s(r, c) at y = y0 - 1 + r  <-  h rows r+2 (y+1, f0), r+1 (y, f1), r (y-1, f2).
    for (int task = tid; task < kSW * 4; task += 256) {
      const int seg = task / kSW, c = task - seg * kSW;    // consecutive lanes = consecutive columns
      const int r0 = seg * 9, r1 = min(r0 + 9, kSH);
      float hm = hbi[r0][c], hc = hbi[r0 + 1][c];
      double q2 = (double)(hm * f2);                        // f2 * h(y-1)
      double q1 = (double)(hc * f1);                        // f1 * h(y)
      for (int r = r0; r < r1; ++r) {
        const float hp = hbi[r + 2][c];
        double acc = 0;
        acc += (double)(hp * f0);
        acc += q1;
        acc += q2;
        sb[r][c] = (float)acc;
        q2 = (double)(hc * f2);
        q1 = (double)(hp * f1);
        hc = hp;
      }
    }
    __syncthreads();
    at_pre_gradient<true>(sb, tid, im, x0, y0, W, H, wpr, amask, acount, theta, mag);
    return;
  }
  // ---- border tiles: clamped coordinates and the reference's border rules
  uint8_t (*ub)[72] = reinterpret_cast<uint8_t (*)[72]>(uw);   // byte b of row r = pixel x0 - 2 + b here
  // thread (lx, ly) = (tid & 63, tid >> 6) walks rows ly, ly+4, ... and columns tx (+64 for the halo columns)
  const int lx = tid & 63, ly = tid >> 6;
  {
    const int xa = min(max(x0 - 2 + lx, 0), W - 1), xb = min(max(x0 - 2 + 64 + lx, 0), W - 1);
    for (int r = ly; r < kUH; r += 4) {
      const uint8_t* row = I + (long long)min(max(y0 - 2 + r, 0), H - 1) * pitch;
      ub[r][lx] = row[xa];
      if (lx < kUW - 64) ub[r][64 + lx] = row[xb];
    }
  }
  __syncthreads();
  // B. horizontal pass (Gaussian.cc:40-68 with its absolute-offset comparison, see at_blur_h_kernel)
  {
    auto hval = [&](int r, int c) -> float {
      const int y = y0 - 2 + r, x = x0 - 1 + c;
      if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
      double acc = 0;
      if (x >= 2 && x <= W - 2) {
        acc += lutd[0][ub[r][c + 2]]; acc += lutd[1][ub[r][c + 1]]; acc += lutd[2][ub[r][c]];
      } else {
        auto px = [&](int xx, int yy) -> int { return I[(long long)yy * pitch + xx]; };
        if (y == 0) {
          if (x == 0) { acc += lutd[0][px(1, 0)]; acc += lutd[1][px(0, 0)]; acc += lutd[2][px(0, 0)]; }
          else if (x == 1) { acc += lutd[0][px(2, 0)]; acc += lutd[1][px(1, 0)]; acc += lutd[2][px(0, 0)]; }
          else { acc += lutd[0][px(W - 1, 0)]; acc += lutd[1][px(W - 1, 0)]; acc += lutd[2][px(W - 2, 0)]; }
        } else if (y == 1 && x == 0) {
          acc += lutd[0][px(0, 1)]; acc += lutd[1][px(0, 1)]; acc += lutd[2][px(W - 1, 0)];
        } else {
          const int a = (x == W - 1) ? px(W - 1, y) : px(0, y);
          acc += lutd[0][a]; acc += lutd[1][a]; acc += lutd[2][a];
        }
      }
      return (float)acc;
    };
    for (int r = ly; r < kHH; r += 4) {
      hb[r][lx] = hval(r, lx);
      if (lx < kHW - 64) hb[r][64 + lx] = hval(r, 64 + lx);
    }
  }
  __syncthreads();
  // C. vertical pass (replicated border)
  {
    auto sval = [&](int r, int c) -> float {
      const int y = y0 - 1 + r, x = x0 - 1 + c;
      if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
      const int rp = min(y + 1, H - 1) - (y0 - 2), rm = max(y - 1, 0) - (y0 - 2), rc = y - (y0 - 2);
      double acc = 0;
      acc += (double)(hb[rp][c] * f0);
      acc += (double)(hb[rc][c] * f1);
      acc += (double)(hb[rm][c] * f2);
      return (float)acc;
    };
    for (int r = ly; r < kSH; r += 4) {
      sb[r][lx] = sval(r, lx);
      if (lx < kSW - 64) sb[r][64 + lx] = sval(r, 64 + lx);
    }
  }
  __syncthreads();
  at_pre_gradient<false>(sb, tid, im, x0, y0, W, H, wpr, amask, acount, theta, mag);
}

// compact id of an ACTIVE pixel from the scanned word counts
__device__ __forceinline__ int compact_id(const uint32_t* __restrict__ amask, const int* __restrict__ aoff, long long wi, int bit) {
  return aoff[wi] + __popc(amask[wi] & ((1u << bit) - 1u));
}

// Union-find forest (UnionFindSimple's parent/size + Edge.cc's theta/mag spans).  The PARENT links live in an int array of
// their own: the root look-ups of the merge sweep -- several hops per edge end, only the link of each hop is needed -- then
// touch 4 bytes per node instead of a 32-byte record, and the links of a whole 214-view sub-batch (~120 MB) stay largely
// L2-resident where the records (~950 MB) streamed from HBM (round 2: 29 GB of DRAM reads for 7.5 M edges, 390 B per edge).
// One 32-byte record per node keeps the statistics: the sweep reads all five values of both ROOTS of a candidate edge, two
// sectors instead of ten scattered ones.
struct __align__(32) UfNode {
  int unused, size;
  float tmin, tmax, mmin, mmax;
  int pad0, pad1;
};

// stage 2 for the fused path: one thread per mask word, emits its active pixels in raster order
__global__ void at_compact_words_kernel(const uint32_t* __restrict__ amask, const int* __restrict__ aoff, long long n_words, int wpr,
                                        int W, int H, const float* __restrict__ theta, const float* __restrict__ mag,
                                        int* __restrict__ apix, float* __restrict__ a_theta, float* __restrict__ a_mag,
                                        UfNode* __restrict__ nodes) {
  const long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= n_words) return;
  uint32_t m = amask[wi];
  if (!m) return;
  const long long row = wi / wpr;                 // im * H + y
  const int xw = (int)(wi - row * wpr) << 5;
  const long long g0 = row * W + xw;              // == im * N + y * W + xw
  int c = aoff[wi];
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    const long long g = g0 + b;
    apix[c] = (int)g;
    a_theta[c] = theta[g];   // Ix  (see at_pre_gradient)
    a_mag[c] = mag[g];       // Iy
    ++c;
  }
}

// theta / mag of the active pixels (TagDetector.cc:159-166) and the union-find nodes, one thread per active pixel: every lane
// of a warp runs the bit-exact atan2f.  In: a_theta = Ix, a_mag = Iy; out: a_theta = atan2f(Iy, Ix), a_mag = Ix^2 + Iy^2.
__global__ void at_theta_kernel(int n, float* __restrict__ a_theta, float* __restrict__ a_mag, UfNode* __restrict__ nodes,
                                int* __restrict__ parent) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const float Ix = a_theta[c], Iy = a_mag[c];
  const float mg = Ix * Ix + Iy * Iy;
  const float th = fd_atan2f(Iy, Ix);
  a_theta[c] = th; a_mag[c] = mg;
  nodes[c] = UfNode{0, 1, th, th, mg, mg, 0, 0};
  parent[c] = c;
}

// ---- stage 3: edges (Edge.cc:14-67) ------------------------------------------------------------------
__device__ __forceinline__ int edge_cost(float theta0, float theta1, float mag1) {
  if (mag1 < kMinMag) return -1;
  const float thetaErr = fabsf(mod2pi(theta1 - theta0));
  if (thetaErr > max_edge_cost()) return -1;
  const float normErr = thetaErr / max_edge_cost();
  return (int)(normErr * 100);
}

// The neighbour's activity bit replaces the mag test (active <=> mag >= minMag), its compact id comes
// from the scanned word counts and its theta from the compact array -- no full-size theta / mag / index images are read.
template <bool EMIT>
__global__ void at_edges_words_kernel(const int* __restrict__ apix, const float* __restrict__ a_theta, const uint32_t* __restrict__ amask,
                                      const int* __restrict__ aoff, int wpr, int n_active, int W, int H, int* __restrict__ ecount,
                                      const int* __restrict__ eoff, int* __restrict__ eA, int* __restrict__ eB,
                                      uint8_t* __restrict__ ecost) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_active) return;
  const long long g = apix[c];
  const long long row = g / W;                    // im * H + y
  const int x = (int)(g - row * W);
  const float th0 = a_theta[c];
  int n = 0;
  int o = EMIT ? eoff[c] : 0;
  auto consider = [&](long long rown, int xn) {
    const long long wi = rown * wpr + (xn >> 5);
    const int bit = xn & 31;
    const uint32_t m = amask[wi];
    if (!((m >> bit) & 1u)) return;               // mag1 < minMag
    const int cn = aoff[wi] + __popc(m & ((1u << bit) - 1u));
    const int cost = edge_cost(th0, a_theta[cn], kMinMag);
    if (cost >= 0) {
      if (EMIT) { eA[o] = c; eB[o] = cn; ecost[o] = (uint8_t)cost; ++o; }
      ++n;
    }
  };
  consider(row, x + 1);          // horizontal
  consider(row + 1, x);          // vertical
  consider(row + 1, x + 1);      // downward diagonal
  if (x != 0) consider(row + 1, x - 1);  // upward diagonal
  if (!EMIT) ecount[c] = n;
}

// ---- stage 4: connected components of the edge graph ------------------------------------------------
__device__ __forceinline__ int ccl_find(const int* L, int x) {
  int p = L[x];
  while (p != x) { x = p; p = L[x]; }
  return x;
}
__global__ void at_ccl_init_kernel(int* L, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) L[i] = i;
}
// Initial forest = the horizontal runs: a pixel whose left neighbour has an edge to it (the 'right' edge is the first one a
// pixel emits, and its target is then the next compact id) starts under the head of its run -- no atomics, and the union
// pass that follows finds most end points one hop from a root.
__global__ void at_ccl_init_runs_kernel(int* __restrict__ L, const int* __restrict__ eoff, const int* __restrict__ eB, int n, int n_edges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int h = i;
  // at most 16 steps back (a thread per pixel walking a 300-pixel run to its head would be quadratic work): a pixel deep
  // inside a long run hangs under a pixel 16 places earlier, which hangs under ... the head
  for (int step = 0; step < 16 && h > 0; ++step) {
    const int e0 = eoff[h - 1], e1 = eoff[h];
    if (e1 > e0 && eB[e0] == h) --h; else break;   // an edge (h-1 -> h): same run (any first edge to the next id links them)
  }
  L[i] = h;
}
// (pointer jumping inside the find, as in ECL-CC, was measured: 5.13 -> 5.83 ms for 214 views -- the extra stores cost more than
// the shorter chains save on these thin components.)
__global__ void at_ccl_union_kernel(int* L, const int* __restrict__ eA, const int* __restrict__ eB, int n_edges) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int a = eA[e], b = eB[e];
  bool done = false;
  while (!done) {
    a = ccl_find(L, a);
    b = ccl_find(L, b);
    if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
    else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
    else done = true;
  }
}
__global__ void at_ccl_flatten_kernel(int* L, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) L[i] = ccl_find(L, i);
}

// ---- stage 5: sort functors --------------------------------------------------------------------------
struct DigitCost {
  const uint8_t* cost;
  __device__ int operator()(uint32_t e) const { return cost[e]; }
};
struct DigitLabelOfEdge {
  const int* label; const int* eA; int shift;
  __device__ int operator()(uint32_t e) const { return (label[eA[e]] >> shift) & 255; }
};
struct DigitArray {
  const int* key; int shift;
  __device__ int operator()(uint32_t v) const { return (key[v] >> shift) & 255; }
};
// key of an edge = label of its component (root compact id); materialised once for the key-value sort
__global__ void at_edge_keys_kernel(const int* __restrict__ label, const int* __restrict__ eA, int n, uint32_t* __restrict__ keys) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) keys[e] = (uint32_t)label[eA[e]];
}
// Dense component numbers.  An edge always runs from a pixel to one LATER in raster order (right, down, both down
// diagonals), and compact ids are raster ordered, so the root of a component (its smallest id) is the first end point of at
// least one edge: the components that own edges are exactly the roots with a non-empty edge range.  Numbering those (scan)
// gives sort keys of ~15 bits instead of the 25 bits of a compact id -- one 9-bit pass less -- and a key that IS the
// component index, so the start of every component in the sorted sequence is written directly (no scan over the edges).
__global__ void at_root_flag_kernel(const int* __restrict__ label, const int* __restrict__ eoff, int n_active, int n_edges, int* __restrict__ flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_active) return;
  const int cnt = (c + 1 < n_active ? eoff[c + 1] : n_edges) - eoff[c];
  flag[c] = (label[c] == c && cnt > 0) ? 1 : 0;
}
__global__ void at_dense_label_kernel(const int* __restrict__ label, const int* __restrict__ dense, int n_active, int* __restrict__ dlabel) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_active) dlabel[c] = dense[label[c]];
}
__global__ void at_starts_dense_kernel(const uint32_t* __restrict__ keys, int n, int n_comp, int* __restrict__ starts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = keys[i];
  if (i == 0 || keys[i - 1] != k) starts[k] = i;
  if (i == n - 1) starts[n_comp] = n;
}
__global__ void at_iota_kernel(uint32_t* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
// flag[i] = 1 where key(order[i]) differs from key(order[i-1])
__global__ void at_boundary_kernel(const uint32_t* __restrict__ order, int n, const int* __restrict__ key,
                                   const int* __restrict__ via, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = via ? key[via[order[i]]] : key[order[i]];
  int f = 1;
  if (i > 0) { const int kp = via ? key[via[order[i - 1]]] : key[order[i - 1]]; f = (k != kp); }
  flag[i] = f;
}
__global__ void at_starts_kernel(const int* __restrict__ flag_scan, const int* __restrict__ total, int n, int* __restrict__ starts) {
  // flag_scan = exclusive scan of boundary flags; element i is a segment start iff scan[i+1] != scan[i]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = flag_scan[i];
  const int nx = (i + 1 < n) ? flag_scan[i + 1] : *total;
  if (nx != s) starts[s] = i;
  if (i == n - 1) starts[*total] = n;
}

// ---- stage 6: the sequential merge sweep, one thread per component (Edge.cc:69-117) -----------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int root = x;
  while (parent[root] != root) root = parent[root];
  while (parent[x] != root) { const int nx = parent[x]; parent[x] = root; x = nx; }  // path compression (UnionFindSimple.cc:6-18)
  return root;
}

// One WARP per component.  The sweep order is strictly sequential in the reference; what a warp adds is
// latency hiding that cannot change the result: (A) the 32 lanes look up the current roots of the next 32
// edges in parallel -- an edge whose end points already share a root can never merge later (sets only grow),
// so it is dropped -- and load the statistics of both roots; (B) the remaining candidates are replayed one by one, in
// order, by the whole warp from REGISTERS (uniform control flow, lane 0 writes the forest): each lane tracks what the
// earlier candidates of the batch did to its roots.  Before, every replayed candidate re-walked its roots and re-read both
// records from memory -- ~1-2 us of dependent loads each, the whole kernel latency-bound on them.
__device__ __forceinline__ int uf_find_ro(const int* parent, int x) {
  int p = parent[x];
  while (p != x) { x = p; p = parent[x]; }
  return x;
}

// Persistent warps pull components from a global counter: most of the ~2000 components of an image hold a handful of
// edges, and with one warp-sized block per component the kernel spent its time launching blocks (round 2: 51 % of the warp
// slots occupied, 1.3 ms for ONE image against 7.4 ms for 214).
__global__ void __launch_bounds__(128) at_merge_kernel(const uint32_t* __restrict__ order, const int* __restrict__ starts, int n_comp,
                                                         const int* __restrict__ eA, const int* __restrict__ eB, UfNode* nodes,
                                                         int* parent, int* __restrict__ work_counter) {
  const int lane = threadIdx.x & 31;
  for (;;) {
  int c = 0;
  if (lane == 0) c = atomicAdd(work_counter, 1);
  c = __shfl_sync(0xffffffffu, c, 0);
  if (c >= n_comp) return;
  const int beg = starts[c], end = starts[c + 1];
  // the end points of the NEXT 32 edges are fetched while this batch is processed: they do not depend on the merges
  int a_nx = 0, b_nx = 0;
  if (beg + lane < end) { const uint32_t e = order[beg + lane]; a_nx = eA[e]; b_nx = eB[e]; }
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    int ra = 0, rb = 0;
    const int a = a_nx, b = b_nx;
    if (i + 32 < end) { const uint32_t e = order[i + 32]; a_nx = eA[e]; b_nx = eB[e]; }
    if (i < end) {
      ra = uf_find_ro(parent, a);
      rb = uf_find_ro(parent, b);
      // path shortening (UnionFindSimple compresses too; it never changes a representative)
      if (parent[a] != ra) parent[a] = ra;
      if (parent[b] != rb) parent[b] = rb;
    }
    // statistics of both roots, loaded by every candidate lane at once; from here on the batch needs no global loads:
    // every merge of this component happens in this warp, so each lane keeps its roots and their statistics current from the
    // broadcast winner / loser / merged record of the merges replayed before its turn.
    int sza = 0, szb = 0;
    float tmina = 0.f, tmaxa = 0.f, mmina = 0.f, mmaxa = 0.f, tminb = 0.f, tmaxb = 0.f, mminb = 0.f, mmaxb = 0.f;
    const bool is_cand = i < end && ra != rb;
    if (is_cand) {
      const int4 a0 = *reinterpret_cast<const int4*>(&nodes[ra]);
      const float2 a1 = *reinterpret_cast<const float2*>(&nodes[ra].mmin);
      const int4 b0 = *reinterpret_cast<const int4*>(&nodes[rb]);
      const float2 b1 = *reinterpret_cast<const float2*>(&nodes[rb].mmin);
      sza = a0.y; tmina = __int_as_float(a0.z); tmaxa = __int_as_float(a0.w); mmina = a1.x; mmaxa = a1.y;
      szb = b0.y; tminb = __int_as_float(b0.z); tmaxb = __int_as_float(b0.w); mminb = b1.x; mmaxb = b1.y;
    }
    unsigned cand = __ballot_sync(0xffffffffu, is_cand);
    while (cand) {
      const int l = __ffs(cand) - 1;
      cand &= cand - 1;
      const int ida = __shfl_sync(0xffffffffu, ra, l), idb = __shfl_sync(0xffffffffu, rb, l);
      if (ida == idb) continue;   // joined by an earlier candidate of this batch
      const int c_sza = __shfl_sync(0xffffffffu, sza, l), c_szb = __shfl_sync(0xffffffffu, szb, l);
      const float c_tmina = __shfl_sync(0xffffffffu, tmina, l), c_tmaxa = __shfl_sync(0xffffffffu, tmaxa, l);
      const float c_tminb = __shfl_sync(0xffffffffu, tminb, l), c_tmaxb = __shfl_sync(0xffffffffu, tmaxb, l);
      const float c_mmina = __shfl_sync(0xffffffffu, mmina, l), c_mmaxa = __shfl_sync(0xffffffffu, mmaxa, l);
      const float c_mminb = __shfl_sync(0xffffffffu, mminb, l), c_mmaxb = __shfl_sync(0xffffffffu, mmaxb, l);
      const float costa = (c_tmaxa - c_tmina);
      const float costb = (c_tmaxb - c_tminb);
      const float bshift = mod2pi_ref((c_tmina + c_tmaxa) / 2, (c_tminb + c_tmaxb) / 2) - (c_tminb + c_tmaxb) / 2;
      const float tminab = fminr(c_tmina, c_tminb + bshift);
      float tmaxab = fmaxr(c_tmaxa, c_tmaxb + bshift);
      if (tmaxab - tminab > 2 * kPiF) tmaxab = tminab + 2 * kPiF;
      const float mminab = fminr(c_mmina, c_mminb);
      const float mmaxab = fmaxr(c_mmaxa, c_mmaxb);
      const float costab = (tmaxab - tminab);
      if (costab <= (fminr(costa, costb) + kThetaThresh / (c_sza + c_szb)) &&
          (mmaxab - mminab) <= fminr(c_mmaxa - c_mmina, c_mmaxb - c_mminb) + kMagThresh / (c_sza + c_szb)) {
        // UnionFindSimple::connectNodes (UnionFindSimple.cc:25-44): the smaller set hangs under the larger one
        const int win = c_sza > c_szb ? ida : idb, lose = c_sza > c_szb ? idb : ida;
        const int szab = c_sza + c_szb;
        if (lane == 0) { parent[lose] = win; nodes[win] = UfNode{0, szab, tminab, tmaxab, mminab, mmaxab, 0, 0}; }
        if (ra == lose) ra = win;
        if (rb == lose) rb = win;
        if (ra == win) { sza = szab; tmina = tminab; tmaxa = tmaxab; mmina = mminab; mmaxa = mmaxab; }
        if (rb == win) { szb = szab; tminb = tminab; tmaxb = tmaxab; mminb = mminab; mmaxb = mmaxab; }
      }
    }
    __syncwarp();
  }
  }   // next component
}

// size array of the forest (debug export)
__global__ void at_split_nodes_kernel(const UfNode* __restrict__ nodes, int n, int* __restrict__ setsize) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  setsize[i] = nodes[i].size;
}

__global__ void at_rep_kernel(const UfNode* __restrict__ nodes, const int* __restrict__ parent, int n_active, int* __restrict__ rep,
                              int* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_active) return;
  const int r = uf_find_ro(parent, i);
  rep[i] = r;
  keep[i] = nodes[r].size >= kMinSegmentSize ? 1 : 0;  // TagDetector.cc:248
}
__global__ void at_keep_compact_kernel(const int* __restrict__ keep_scan, const int* __restrict__ rep, const UfNode* __restrict__ nodes,
                                       int n_active, uint32_t* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_active) return;
  if (nodes[rep[i]].size >= kMinSegmentSize) list[keep_scan[i]] = (uint32_t)i;
}

// ---- stage 8: line fit per cluster (GLine2D.cc:54-87, GLineSegment2D.cc:9-24, TagDetector.cc:265-325) ---------
struct Segment {
  float x0, y0, x1, y1, theta, length;
  int image;
  int pad;
};

__global__ void at_fit_kernel(const uint32_t* __restrict__ pts, const int* __restrict__ cstart, int n_clusters,
                              const int* __restrict__ apix, const float* __restrict__ a_theta, const float* __restrict__ a_mag,
                              int W, int H, Segment* __restrict__ segs, int* __restrict__ valid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clusters) return;
  const int beg = cstart[c], end = cstart[c + 1];
  const long long N = (long long)W * H;
  float mXX = 0, mYY = 0, mXY = 0, mX = 0, mY = 0, n = 0;
  for (int i = beg; i < end; ++i) {
    const int a = pts[i];
    const int p = (int)(apix[a] % N);
    const float y = (float)(p / W), x = (float)(p % W);
    const float alpha = a_mag[a];
    mY += y * alpha;
    mX += x * alpha;
    mYY += y * y * alpha;
    mXX += x * x * alpha;
    mXY += x * y * alpha;
    n += alpha;
  }
  const float Ex = mX / n, Ey = mY / n;
  const float Cxx = mXX / n - (mX / n) * (mX / n);
  const float Cyy = mYY / n - (mY / n) * (mY / n);
  const float Cxy = mXY / n - (mX / n) * (mY / n);
  const float phi = 0.5f * fd_atan2f(-2 * Cxy, (Cyy - Cxx));
  float dx = -gl_sinf(phi), dy = gl_cosf(phi);
  float px = Ex, py = Ey;
  // GLine2D::normalizeSlope
  {
    const float mag = sqrtf(dx * dx + dy * dy);
    dx /= mag; dy /= mag;
  }
  float maxc = -INFINITY, minc = INFINITY;
  for (int i = beg; i < end; ++i) {
    const int a = pts[i];
    const int p = (int)(apix[a] % N);
    const float y = (float)(p / W), x = (float)(p % W);
    const float coord = x * dx + y * dy;
    maxc = fmaxr(maxc, coord);
    minc = fminr(minc, coord);
  }
  // GLine2D::normalizeP
  {
    const float dotprod = -dy * px + dx * py;
    px = -dy * dotprod; py = dx * dotprod;
  }
  const float p0x = px + minc * dx, p0y = py + minc * dy;
  const float p1x = px + maxc * dx, p1y = py + maxc * dy;
  const float length = dist2d(p0x, p0y, p1x, p1y);
  Segment s;
  s.image = (int)(apix[pts[beg]] / N);
  s.pad = 0;
  if (length < kMinLineLength) { valid[c] = 0; s.x0 = s.y0 = s.x1 = s.y1 = s.theta = s.length = 0; segs[c] = s; return; }
  const float sdy = p1y - p0y, sdx = p1x - p0x;
  float theta = fd_atan2f(sdy, sdx);
  float flip = 0, noflip = 0;
  for (int i = beg; i < end; ++i) {
    const int a = pts[i];
    const float err = mod2pi(a_theta[a] - theta);
    if (err < 0) noflip += a_mag[a]; else flip += a_mag[a];
  }
  if (flip > noflip) theta = theta + kPiF;
  const float dot = sdx * gl_cosf(theta) + sdy * gl_sinf(theta);
  if (dot > 0) { s.x0 = p1x; s.y0 = p1y; s.x1 = p0x; s.y1 = p0y; }
  else { s.x0 = p0x; s.y0 = p0y; s.x1 = p1x; s.y1 = p1y; }
  s.theta = theta; s.length = length;
  segs[c] = s;
  valid[c] = 1;
}

__global__ void at_seg_compact_kernel(const Segment* __restrict__ in, const int* __restrict__ valid, const int* __restrict__ vscan,
                                      int n, Segment* __restrict__ out, int* __restrict__ img_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !valid[i]) return;
  out[vscan[i]] = in[i];
  atomicAdd(&img_count[in[i].image], 1);
}

// ---- stage 9a: spatial hash (Gridder.h) + children (TagDetector.cc:344-388) ---------------------------------
struct GridDims { int gw, gh; };
__host__ __device__ inline GridDims grid_dims(int W, int H) {
  GridDims g;
  g.gw = (int)(((float)W - 0.f) / 10.f + 1);
  g.gh = (int)(((float)H - 0.f) / 10.f + 1);
  return g;
}
__global__ void at_cell_count_kernel(const Segment* __restrict__ segs, int n, int W, int H, int* __restrict__ cell_of,
                                     int* __restrict__ cell_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GridDims g = grid_dims(W, H);
  const int ix = (int)((segs[i].x0 - 0.f) / 10.f);
  const int iy = (int)((segs[i].y0 - 0.f) / 10.f);
  int cell = -1;
  if (ix >= 0 && iy >= 0 && ix < g.gw && iy < g.gh) {
    cell = (segs[i].image * g.gh + iy) * g.gw + ix;
    atomicAdd(&cell_count[cell], 1);
  }
  cell_of[i] = cell;
}
__global__ void at_cell_fill_kernel(const int* __restrict__ cell_of, int n, const int* __restrict__ cell_start, int* __restrict__ cell_fill,
                                    int* __restrict__ cell_items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = cell_of[i];
  if (cell < 0) return;
  cell_items[cell_start[cell] + atomicAdd(&cell_fill[cell], 1)] = i;
}
// LIFO order of Gridder::add: later (higher index) segments come first in a cell's list.
__global__ void at_cell_sort_kernel(const int* __restrict__ cell_start, const int* __restrict__ cell_count, int n_cells,
                                    int* __restrict__ cell_items) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int b = cell_start[c], m = cell_count[c];
  for (int i = 1; i < m; ++i) {
    const int v = cell_items[b + i];
    int j = i - 1;
    while (j >= 0 && cell_items[b + j] < v) { cell_items[b + j + 1] = cell_items[b + j]; --j; }
    cell_items[b + j + 1] = v;
  }
}

// GLine2D::intersectionWith for lines through (ax0,ay0)-(ax1,ay1) and (bx0,by0)-(bx1,by1); returns false if parallel
__device__ __forceinline__ bool line_intersect(float ax0, float ay0, float ax1, float ay1, float bx0, float by0, float bx1,
                                               float by1, float& ox, float& oy) {
  const float dx = ax1 - ax0, dy = ay1 - ay0;
  const float ldx = bx1 - bx0, ldy = by1 - by0;
  const float m00 = dx, m01 = -ldx, m10 = dy, m11 = -ldy;
  const float det = m00 * m11 - m01 * m10;
  if (fabs((double)det) < 1e-10) { ox = -1; oy = 0; return false; }
  const float i00 = m11 / det;
  const float i01 = -m01 / det;
  const float b00 = bx0 - ax0;
  const float b10 = by0 - ay0;
  const float x00 = i00 * b00 + i01 * b10;
  ox = dx * x00 + ax0;
  oy = dy * x00 + ay0;
  return true;
}

template <bool EMIT>
__global__ void at_children_kernel(const Segment* __restrict__ segs, int n, int W, int H, const int* __restrict__ cell_start,
                                   const int* __restrict__ cell_count, const int* __restrict__ cell_items,
                                   int* __restrict__ ccount, const int* __restrict__ coff, int* __restrict__ children) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Segment P = segs[i];
  const GridDims g = grid_dims(W, H);
  const float range = 0.5f * P.length;
  int ix0 = (int)((P.x1 - range - 0.f) / 10.f), iy0 = (int)((P.y1 - range - 0.f) / 10.f);
  int ix1 = (int)((P.x1 + range - 0.f) / 10.f), iy1 = (int)((P.y1 + range - 0.f) / 10.f);
  ix0 = min(g.gw - 1, max(0, ix0)); ix1 = min(g.gw - 1, max(0, ix1));
  iy0 = min(g.gh - 1, max(0, iy0)); iy1 = min(g.gh - 1, max(0, iy1));
  int cnt = 0;
  int o = EMIT ? coff[i] : 0;
  for (int iy = iy0; iy <= iy1; ++iy)
    for (int ix = ix0; ix <= ix1; ++ix) {
      const int cell = (P.image * g.gh + iy) * g.gw + ix;
      const int b = cell_start[cell], m = cell_count[cell];
      for (int q = 0; q < m; ++q) {
        const int ci = cell_items[b + q];
        const Segment C = segs[ci];
        if (mod2pi(C.theta - P.theta) > 0) continue;
        float px, py;
        line_intersect(P.x0, P.y0, P.x1, P.y1, C.x0, C.y0, C.x1, C.y1, px, py);
        if (px == -1) continue;
        const float parentDist = dist2d(px, py, P.x1, P.y1);
        const float childDist = dist2d(px, py, C.x0, C.y0);
        if (fmaxr(parentDist, childDist) > P.length) continue;
        if (EMIT) children[o++] = ci;
        ++cnt;
      }
    }
  if (!EMIT) ccount[i] = cnt;
}

// ---- stage 9b: quad search (Quad.cc:52-157) --------------------------------------------------------------
struct QuadRec {
  float p[4][2];
  float perimeter;
  int image;
};

__device__ bool make_quad(const Segment* __restrict__ segs, const int path[5], QuadRec& q) {
  float p[4][2];
  float perim = 0;
  bool bad = false;
  for (int i = 0; i < 4; ++i) {
    const Segment A = segs[path[i]], B = segs[path[i + 1]];
    line_intersect(A.x0, A.y0, A.x1, A.y1, B.x0, B.y0, B.x1, B.y1, p[i][0], p[i][1]);
    perim += A.length;
    if (p[i][0] == -1) bad = true;
  }
  if (!bad) {
    const float t0 = fd_atan2f(p[1][1] - p[0][1], p[1][0] - p[0][0]);
    const float t1 = fd_atan2f(p[2][1] - p[1][1], p[2][0] - p[1][0]);
    const float t2 = fd_atan2f(p[3][1] - p[2][1], p[3][0] - p[2][0]);
    const float t3 = fd_atan2f(p[0][1] - p[3][1], p[0][0] - p[3][0]);
    const float ttheta = mod2pi(t1 - t0) + mod2pi(t2 - t1) + mod2pi(t3 - t2) + mod2pi(t0 - t3);
    if (ttheta < -7 || ttheta > -5) bad = true;
  }
  if (!bad) {
    const float d0 = dist2d(p[0][0], p[0][1], p[1][0], p[1][1]);
    const float d1 = dist2d(p[1][0], p[1][1], p[2][0], p[2][1]);
    const float d2 = dist2d(p[2][0], p[2][1], p[3][0], p[3][1]);
    const float d3 = dist2d(p[3][0], p[3][1], p[0][0], p[0][1]);
    const float d4 = dist2d(p[0][0], p[0][1], p[2][0], p[2][1]);
    const float d5 = dist2d(p[1][0], p[1][1], p[3][0], p[3][1]);
    if (d0 < kMinEdgeLength || d1 < kMinEdgeLength || d2 < kMinEdgeLength || d3 < kMinEdgeLength || d4 < kMinEdgeLength ||
        d5 < kMinEdgeLength)
      bad = true;
    const float dmax = fmaxr(fmaxr(d0, d1), fmaxr(d2, d3));
    const float dmin = fminr(fminr(d0, d1), fminr(d2, d3));
    if (dmax > dmin * kMaxQuadAspect) bad = true;
  }
  if (bad) return false;
  for (int i = 0; i < 4; ++i) { q.p[i][0] = p[i][0]; q.p[i][1] = p[i][1]; }
  q.perimeter = perim;
  q.image = segs[path[0]].image;
  return true;
}

template <bool EMIT>
__global__ void at_quads_kernel(const Segment* __restrict__ segs, int n, const int* __restrict__ coff, const int* __restrict__ children,
                                int* __restrict__ qcount, const int* __restrict__ qoff, QuadRec* __restrict__ quads) {
  const int s0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (s0 >= n) return;
  int path[5];
  path[0] = s0;
  const float th0 = segs[s0].theta;
  int cnt = 0;
  int o = EMIT ? qoff[s0] : 0;
  // depth-first over children lists, depth 4 (iterative form of Quad::search)
  for (int a = coff[s0]; a < coff[s0 + 1]; ++a) {
    const int c1 = children[a];
    if (segs[c1].theta > th0) continue;
    path[1] = c1;
    for (int b = coff[c1]; b < coff[c1 + 1]; ++b) {
      const int c2 = children[b];
      if (segs[c2].theta > th0) continue;
      path[2] = c2;
      for (int c = coff[c2]; c < coff[c2 + 1]; ++c) {
        const int c3 = children[c];
        if (segs[c3].theta > th0) continue;
        path[3] = c3;
        for (int d = coff[c3]; d < coff[c3 + 1]; ++d) {
          const int c4 = children[d];
          if (segs[c4].theta > th0) continue;
          path[4] = c4;
          if (c4 != s0) continue;
          QuadRec q;
          if (make_quad(segs, path, q)) {
            if (EMIT) quads[o++] = q;
            ++cnt;
          }
        }
      }
    }
  }
  if (!EMIT) qcount[s0] = cnt;
}

// ---- stage 9c: decode (TagDetector.cc:432-532, GrayModel.cc, TagFamily.cc:45-112) -------------------------------
struct Detection {
  float p[4][2];
  float cxy[2];
  float perimeter;
  int id, hamming, rotation, good, image;
  unsigned long long obs_code;
};

struct GrayModelD {
  double A[4][4];
  double b[4];
  double v[4];
  int nobs;
  __device__ void init() {
    for (int i = 0; i < 4; ++i) { b[i] = 0; v[i] = 0; for (int j = 0; j < 4; ++j) A[i][j] = 0; }
    nobs = 0;
  }
  __device__ void add(float x, float y, float gray) {
    const float xy = x * y;
    A[0][0] += x * x; A[0][1] += x * y; A[0][2] += x * xy; A[0][3] += x;
    A[1][1] += y * y; A[1][2] += y * xy; A[1][3] += y;
    A[2][2] += xy * xy; A[2][3] += xy;
    A[3][3] += 1;
    b[0] += x * gray; b[1] += y * gray; b[2] += xy * gray; b[3] += gray;
    nobs++;
  }
  // cofactor inverse as Eigen's compute_inverse_size4 (scalar path); invertible iff |det| > 1e-12
  __device__ void compute() {
    if (nobs >= 6) {
      for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) A[j][i] = A[i][j];
      double res[4][4];
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          const int i1 = (i + 1) & 3, i2 = (i + 2) & 3, i3 = (i + 3) & 3;
          const int j1 = (j + 1) & 3, j2 = (j + 2) & 3, j3 = (j + 3) & 3;
          const double h0 = A[i1][j1] * (A[i2][j2] * A[i3][j3] - A[i2][j3] * A[i3][j2]);
          const double h1 = A[i2][j1] * (A[i3][j2] * A[i1][j3] - A[i3][j3] * A[i1][j2]);
          const double h2 = A[i3][j1] * (A[i1][j2] * A[i2][j3] - A[i1][j3] * A[i2][j2]);
          const double cof = h0 + h1 + h2;
          res[j][i] = ((i + j) & 1) ? -cof : cof;
        }
      double det = 0;
      for (int i = 0; i < 4; ++i) det += A[i][0] * res[0][i];
      if (fabs(det) > 1e-12) {
        for (int i = 0; i < 4; ++i) {
          double s = (res[i][0] / det) * b[0];
          for (int k = 1; k < 4; ++k) s += (res[i][k] / det) * b[k];
          v[i] = s;
        }
        return;
      }
    }
    v[0] = v[1] = v[2] = 0;
    v[3] = b[3] / nobs;
  }
  __device__ float interp(float x, float y) const { return (float)(v[0] * x + v[1] * y + v[2] * x * y + v[3]); }
};

// Quad::interpolate with INTERPOLATE defined (Quad.cc:37-50)
__device__ __forceinline__ void quad_interp01(const QuadRec& q, float x01, float y01, float& ox, float& oy) {
  const float x = 2 * x01 - 1, y = 2 * y01 - 1;
  const float kx = (float)((double)x + 1.);  // (x+1.) is a double expression, narrowed to the vector's Scalar
  const float p01x = q.p[1][0] - q.p[0][0], p01y = q.p[1][1] - q.p[0][1];
  const float p32x = q.p[2][0] - q.p[3][0], p32y = q.p[2][1] - q.p[3][1];
  const float r1x = q.p[0][0] + (p01x * kx) / 2.f, r1y = q.p[0][1] + (p01y * kx) / 2.f;
  const float r2x = q.p[3][0] + (p32x * kx) / 2.f, r2y = q.p[3][1] + (p32y * kx) / 2.f;
  const float ky = y + 1;
  ox = r1x + ((r2x - r1x) * ky) / 2.f;
  oy = r1y + ((r2y - r1y) * ky) / 2.f;
}

__device__ __forceinline__ unsigned long long rotate90(unsigned long long w, int d) {
  unsigned long long wr = 0;
  for (int r = d - 1; r >= 0; r--)
    for (int c = 0; c < d; c++) {
      const int b = r + d * c;
      wr = wr << 1;
      if ((w & (1ULL << b)) != 0) wr |= 1;
    }
  return wr;
}

// exact homography through (-1,-1),(1,-1),(1,1),(-1,1) -> quad points minus the optical centre; only used to
// find which quad corner is nearest to H*(-1,-1) after the decoded rotation (TagDetector.cc:496-524)
__device__ void homography4(const float dst[4][2], double H[3][3]) {
  const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
  double A[8][9];
  for (int i = 0; i < 4; ++i) {
    const double x = sx[i], y = sy[i], u = dst[i][0], v = dst[i][1];
    const double r0[9] = {x, y, 1, 0, 0, 0, -u * x, -u * y, u};
    const double r1[9] = {0, 0, 0, x, y, 1, -v * x, -v * y, v};
    for (int k = 0; k < 9; ++k) { A[2 * i][k] = r0[k]; A[2 * i + 1][k] = r1[k]; }
  }
  for (int c = 0; c < 8; ++c) {
    int p = c;
    for (int r = c + 1; r < 8; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (p != c) for (int k = 0; k < 9; ++k) { const double t = A[p][k]; A[p][k] = A[c][k]; A[c][k] = t; }
    const double piv = A[c][c];
    for (int r = 0; r < 8; ++r) {
      if (r == c || piv == 0.0) continue;
      const double f = A[r][c] / piv;
      for (int k = c; k < 9; ++k) A[r][k] -= f * A[c][k];
    }
  }
  double h[9];
  for (int i = 0; i < 8; ++i) h[i] = A[i][i] != 0.0 ? A[i][8] / A[i][i] : 0.0;
  h[8] = 1.0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H[i][j] = h[i * 3 + j];
}

__global__ void at_decode_kernel(const QuadRec* __restrict__ quads, int n, const uint8_t* __restrict__ imgs, long long pitch,
                                 long long istride, int W, int H, int img0, Detection* __restrict__ dets) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= n) return;
  const QuadRec q = quads[qi];
  const uint8_t* I = imgs + (long long)(img0 + q.image) * istride;
  auto fim = [&](int x, int y) -> float { return (float)((double)I[(long long)y * pitch + x] / 255.); };
  const int dd = 2 * 2 + 6;  // 2*blackBorder + dimension
  GrayModelD black, white;
  black.init(); white.init();
  for (int iy = -1; iy <= dd; iy++) {
    const float y = (iy + 0.5f) / dd;
    for (int ix = -1; ix <= dd; ix++) {
      const float x = (ix + 0.5f) / dd;
      float px, py;
      quad_interp01(q, x, y, px, py);
      const int irx = (int)((double)px + 0.5);
      const int iry = (int)((double)py + 0.5);
      if (irx < 0 || irx >= W || iry < 0 || iry >= H) continue;
      const float v = fim(irx, iry);
      if (iy == -1 || iy == dd || ix == -1 || ix == dd) white.add(x, y, v);
      else if (iy == 0 || iy == (dd - 1) || ix == 0 || ix == (dd - 1)) black.add(x, y, v);
    }
  }
  black.compute(); white.compute();
  bool bad = false;
  unsigned long long tagCode = 0;
  for (int iy = 5; iy >= 0; iy--) {
    const float y = (2 + iy + 0.5f) / dd;
    for (int ix = 0; ix < 6; ix++) {
      const float x = (2 + ix + 0.5f) / dd;
      float px, py;
      quad_interp01(q, x, y, px, py);
      const int irx = (int)((double)px + 0.5);
      const int iry = (int)((double)py + 0.5);
      if (irx < 0 || irx >= W || iry < 0 || iry >= H) { bad = true; continue; }
      const float threshold = (black.interp(x, y) + white.interp(x, y)) * 0.5f;
      const float v = fim(irx, iry);
      tagCode = tagCode << 1;
      if (v > threshold) tagCode |= 1;
    }
  }
  Detection d;
  d.good = 0; d.id = -1; d.hamming = 0; d.rotation = 0; d.image = q.image; d.obs_code = tagCode;
  d.perimeter = q.perimeter; d.cxy[0] = d.cxy[1] = 0;
  for (int i = 0; i < 4; ++i) { d.p[i][0] = q.p[i][0]; d.p[i][1] = q.p[i][1]; }
  if (!bad) {
    // TagFamily::decode: min Hamming over 587 codes x 4 rotations, first minimum in (id, rot) order wins
    unsigned long long rc[4];
    rc[0] = tagCode; rc[1] = rotate90(rc[0], 6); rc[2] = rotate90(rc[1], 6); rc[3] = rotate90(rc[2], 6);
    int bestId = -1, bestH = 0x7fffffff, bestRot = 0;
    for (int id = 0; id < kNumTag36h11; ++id) {
      const unsigned long long code = c_tag36h11[id];
      for (int rot = 0; rot < 4; ++rot) {
        const int hd = __popcll(rc[rot] ^ code);
        if (hd < bestH) { bestH = hd; bestRot = rot; bestId = id; }
      }
    }
    d.id = bestId; d.hamming = bestH; d.rotation = bestRot;
    d.good = (bestH <= 1) ? 1 : 0;  // errorRecoveryBits = 1
    // homography (optical centre = (W/2, H/2) as ints) rotated by the decoded rotation
    const float cx = (float)(W / 2), cy = (float)(H / 2);
    float dst[4][2];
    for (int i = 0; i < 4; ++i) { dst[i][0] = q.p[i][0] - cx; dst[i][1] = q.p[i][1] - cy; }
    double Hm[3][3];
    homography4(dst, Hm);
    const float c = gl_cosf(bestRot * kPiF / 2);
    const float s = gl_sinf(bestRot * kPiF / 2);
    double R[3][3] = {{c, -s, 0}, {s, c, 0}, {0, 0, 1}};
    double T[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { double a = Hm[i][0] * R[0][j]; a += Hm[i][1] * R[1][j]; a += Hm[i][2] * R[2][j]; T[i][j] = a; }
    // TagDetection::interpolate(-1,-1)
    float blx = 0, bly = 0;
    {
      const float x = -1, y = -1;
      const float z = (float)(T[2][0] * x + T[2][1] * y + T[2][2]);
      if (z != 0) {
        blx = (float)((T[0][0] * x + T[0][1] * y + T[0][2]) / z + cx);
        bly = (float)((T[1][0] * x + T[1][1] * y + T[1][2]) / z + cy);
      }
    }
    int bestR = -1;
    float bestDist = FLT_MAX;
    for (int i = 0; i < 4; ++i) {
      const float dist = dist2d(blx, bly, q.p[i][0], q.p[i][1]);
      if (dist < bestDist) { bestDist = dist; bestR = i; }
    }
    for (int i = 0; i < 4; ++i) { d.p[i][0] = q.p[(i + bestR) % 4][0]; d.p[i][1] = q.p[(i + bestR) % 4][1]; }
    if (d.good) quad_interp01(q, 0.5f, 0.5f, d.cxy[0], d.cxy[1]);
  }
  dets[qi] = d;
}

// ---- stage 9d: de-duplication per image (TagDetector.cc:547-580) and output assembly -------------------------
__device__ bool overlaps_too_much(const Detection& a, const Detection& b) {
  const float radius = (dist2d(a.p[0][0], a.p[0][1], a.p[1][0], a.p[1][1]) + dist2d(a.p[1][0], a.p[1][1], a.p[2][0], a.p[2][1]) +
                        dist2d(a.p[2][0], a.p[2][1], a.p[3][0], a.p[3][1]) + dist2d(a.p[3][0], a.p[3][1], a.p[0][0], a.p[0][1]) +
                        dist2d(b.p[0][0], b.p[0][1], b.p[1][0], b.p[1][1]) + dist2d(b.p[1][0], b.p[1][1], b.p[2][0], b.p[2][1]) +
                        dist2d(b.p[2][0], b.p[2][1], b.p[3][0], b.p[3][1]) + dist2d(b.p[3][0], b.p[3][1], b.p[0][0], b.p[0][1])) /
                       16.0f;
  const float dist = dist2d(a.cxy[0], a.cxy[1], b.cxy[0], b.cxy[1]);
  return dist < radius;
}

// one thread per image; quads of image i are dets[qstart[i] .. qstart[i+1]); the surviving detections are
// written back, in the reference's order, to the front of the same range; ngood[i] = their number.
__global__ void at_dedup_kernel(Detection* __restrict__ dets, const int* __restrict__ qstart, int n_img, int* __restrict__ ngood) {
  const int im = blockIdx.x * blockDim.x + threadIdx.x;
  if (im >= n_img) return;
  const int beg = qstart[im], end = qstart[im + 1];
  int ng = 0;
  for (int i = beg; i < end; ++i) {
    const Detection t = dets[i];
    if (!t.good) continue;
    bool newFeature = true;
    for (int o = 0; o < ng; ++o) {
      Detection& other = dets[beg + o];
      if (t.id != other.id || !overlaps_too_much(t, other)) continue;
      newFeature = false;
      if (t.hamming > other.hamming) continue;
      if (t.hamming < other.hamming || t.perimeter > other.perimeter) other = t;
    }
    if (newFeature) { dets[beg + ng] = t; ++ng; }
  }
  ngood[im] = ng;
}

// gather the corners to refine: detections with id < nb_tags whose 4 corners all lie inside the image
// (cv::cornerSubPix asserts that; the reference would throw otherwise)
__global__ void at_gather_corners_kernel(const Detection* __restrict__ dets, const int* __restrict__ qstart, const int* __restrict__ ngood,
                                         int n_img, int nb_tags, int W, int H, int img0, float2* __restrict__ corners,
                                         int* __restrict__ cimg, int* __restrict__ refine_flag) {
  const int im = blockIdx.x, t = threadIdx.x;
  if (im >= n_img) return;
  const int beg = qstart[im];
  for (int k = t; k < ngood[im]; k += blockDim.x) {
    const Detection& d = dets[beg + k];
    bool inside = true;
    for (int i = 0; i < 4; ++i)
      inside = inside && (d.p[i][0] >= 0 && d.p[i][0] < (float)W && d.p[i][1] >= 0 && d.p[i][1] < (float)H);
    const int flag = (d.id < nb_tags && inside) ? 1 : 0;
    refine_flag[beg + k] = flag;
    for (int i = 0; i < 4; ++i) {
      // corners that are not refined still occupy a slot (clamped inside so that the kernel may run on them)
      float2 c = make_float2(d.p[i][0], d.p[i][1]);
      if (!flag) c = make_float2(fminf(fmaxf(c.x, 0.f), (float)(W - 1)), fminf(fmaxf(c.y, 0.f), (float)(H - 1)));
      corners[(size_t)(beg + k) * 4 + i] = c;
      cimg[(size_t)(beg + k) * 4 + i] = img0 + im;
    }
  }
}

// aprilgrid_detector.hpp:35-70: slots pre-filled with (-1,-1), detections scattered in order (later overwrite)
__global__ void at_scatter_kernel(const Detection* __restrict__ dets, const int* __restrict__ qstart, const int* __restrict__ ngood,
                                  const float2* __restrict__ refined, const int* __restrict__ refine_flag, int n_img, int nb_tags,
                                  float2* __restrict__ out_corners, int* __restrict__ out_ndet, uint8_t* __restrict__ out_success) {
  const int im = blockIdx.x;
  if (im >= n_img) return;
  float2* oc = out_corners + (size_t)im * 4 * nb_tags;
  for (int k = threadIdx.x; k < 4 * nb_tags; k += blockDim.x) oc[k] = make_float2(-1.f, -1.f);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int beg = qstart[im], ng = ngood[im];
    for (int k = 0; k < ng; ++k) {
      const Detection& d = dets[beg + k];
      if (d.id >= nb_tags) continue;
      for (int i = 0; i < 4; ++i)
        oc[d.id * 4 + i] = refine_flag[beg + k] ? refined[(size_t)(beg + k) * 4 + i] : make_float2(d.p[i][0], d.p[i][1]);
    }
    out_ndet[im] = ng;
    const int threshold = (nb_tags / 2) - 1;
    out_success[im] = ng > threshold ? 1 : 0;
  }
}

__global__ void at_image_starts_kernel(const int* __restrict__ counts, int n_img, int* __restrict__ starts) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < n_img; ++i) { starts[i] = s; s += counts[i]; }
    starts[n_img] = s;
  }
}
__global__ void at_quad_image_count_kernel(const QuadRec* __restrict__ quads, int n, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&counts[quads[i].image], 1);
}

}  // namespace ezc

using namespace ezc;

// ============================================================================================
// Host orchestration
// ============================================================================================
struct ezc_apriltag_debug {
  // optional stage products of the LAST image processed (tests only)
  std::vector<float> theta, mag;
  std::vector<int> edges;      // (cost, pixelA, pixelB) triples in the order the merge sweep consumed them
  std::vector<int> rep, size;  // per pixel (dense), -1 for inactive pixels
  std::vector<float> segments; // x0,y0,x1,y1,theta,length
  std::vector<float> quads;    // 4 corners + perimeter
};

namespace {

struct Pools {
  DevBuf fr, seg, theta, mag, cidx, scan_tmp, total, amask, acount, dbg_theta, dbg_mag;
  DevBuf apix, a_theta, a_mag, nodes, parent, setsize, ecount, label;
  DevBuf eA, eB, ecost, order0, order1, keys0, keys1, flags, starts, work;
  DevBuf rep, keep, plist0, plist1, cstart, segs_raw, seg_valid, segs, img_count, seg_start;
  DevBuf cell_of, cell_count, cell_start, cell_fill, cell_items, ccount, children;
  DevBuf qcount, quads, dets, qimg_count, qstart, ngood, corners, cimg, refine_flag;
  SortTemp sort;
  HostPinned h;
  DevBuf out_corners, out_ndet, out_succ;
};

Pools& get_pools(ezc_context* ctx) {
  if (!ctx->at_pools) {
    ctx->at_pools = new Pools();
    ctx->at_pools_free = [](void* p) { delete static_cast<Pools*>(p); };
  }
  return *static_cast<Pools*>(ctx->at_pools);
}

#define AT_LAUNCH(kernel, n, ...)                                                     \
  do {                                                                                \
    if ((n) > 0) {                                                                    \
      const long long _n = (n);                                                       \
      kernel<<<(unsigned)((_n + 255) / 256), 256, 0, ctx->stream>>>(__VA_ARGS__);     \
      EZC_CUDA(cudaGetLastError());                                                   \
      ctx->kernel_launches++;                                                         \
    }                                                                                 \
  } while (0)

ezc_status read_int(ezc_context* ctx, const int* d, int* h_out, Pools& P) {
  EZC_CUDA(cudaMemcpyAsync(P.h.p, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  *h_out = *P.h.as<int>();
  return EZC_OK;
}

int bits_for(int n) { int b = 1; while ((1LL << b) < n) ++b; return b; }

// Process images [img0, img0 + nb) of the batch.
ezc_status process_sub_batch(ezc_context* ctx, const ezc_images* imgs, int img0, int nb, int nb_tags, Pools& P,
                             float2* d_out_corners, int* d_out_ndet, uint8_t* d_out_success, ezc_apriltag_debug* dbg) {
  const int W = imgs->width, H = imgs->height;
  const long long N = (long long)W * H;
  const long long total = N * nb;
  const uint8_t* d_img = imgs->d + (long long)img0 * imgs->image_stride;
  EZC_CUDA(images_wait(ctx, imgs, img0 + nb - 1));  // asynchronous upload: only this sub-batch has to be resident
  EZC_CUDA(P.theta.reserve(sizeof(float) * total));   // written at active pixels only
  EZC_CUDA(P.mag.reserve(sizeof(float) * total));
  const int wpr = (W + 31) / 32;
  const long long n_words = (long long)nb * H * wpr;
  EZC_CUDA(P.amask.reserve(sizeof(uint32_t) * n_words));
  EZC_CUDA(P.acount.reserve(sizeof(int) * n_words));
  EZC_CUDA(P.total.reserve(sizeof(int) * 16));
  EZC_CUDA(P.h.reserve(4096));
  int* d_tot = P.total.as<int>();

  // ---- 1: preprocessing (fused: blur, gradient, activity mask, theta at active pixels)
  {
    const int tiles_x = (W + kTX - 1) / kTX, tiles_y = (H + kTY - 1) / kTY;
    const long long blocks = (long long)tiles_x * tiles_y * nb;
    at_pre_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_img, imgs->pitch, imgs->image_stride, W, H, tiles_x, tiles_y, wpr,
                                                             P.amask.as<uint32_t>(), P.acount.as<int>(), P.theta.as<float>(), P.mag.as<float>());
    EZC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
  }
  if (dbg && nb == 1) {
    // export path: the three unfused kernels produce the FULL theta / mag images the stage test compares with the
    // reference (theta of inactive pixels included); the pipeline itself continues on the fused products
    EZC_CUDA(P.fr.reserve(sizeof(float) * total)); EZC_CUDA(P.seg.reserve(sizeof(float) * total));
    EZC_CUDA(P.dbg_theta.reserve(sizeof(float) * total)); EZC_CUDA(P.dbg_mag.reserve(sizeof(float) * total));
    EZC_CUDA(P.cidx.reserve(sizeof(int) * total));
    AT_LAUNCH(at_blur_h_kernel, total, d_img, imgs->pitch, imgs->image_stride, W, H, nb, P.fr.as<float>());
    AT_LAUNCH(at_blur_v_kernel, total, P.fr.as<float>(), W, H, nb, P.seg.as<float>());
    AT_LAUNCH(at_gradient_kernel, total, P.seg.as<float>(), W, H, nb, P.dbg_theta.as<float>(), P.dbg_mag.as<float>(), P.cidx.as<int>());
  }
  // ---- 2: compaction (scan of the per-word counts: raster order is preserved)
  if (ezc_status st = exclusive_scan_i32(ctx, P.acount.as<int>(), P.acount.as<int>(), n_words, P.scan_tmp, d_tot)) return st;
  int nA = 0;
  if (ezc_status st = read_int(ctx, d_tot, &nA, P)) return st;
  const size_t cap = (size_t)std::max(nA, 1);
  EZC_CUDA(P.apix.reserve(sizeof(int) * cap)); EZC_CUDA(P.a_theta.reserve(sizeof(float) * cap)); EZC_CUDA(P.a_mag.reserve(sizeof(float) * cap));
  EZC_CUDA(P.nodes.reserve(sizeof(UfNode) * cap)); EZC_CUDA(P.parent.reserve(sizeof(int) * cap));
  EZC_CUDA(P.ecount.reserve(sizeof(int) * (cap + 1))); EZC_CUDA(P.label.reserve(sizeof(int) * cap));
  EZC_CUDA(P.rep.reserve(sizeof(int) * cap)); EZC_CUDA(P.keep.reserve(sizeof(int) * (cap + 1)));
  AT_LAUNCH(at_compact_words_kernel, n_words, P.amask.as<uint32_t>(), P.acount.as<int>(), n_words, wpr, W, H, P.theta.as<float>(),
            P.mag.as<float>(), P.apix.as<int>(), P.a_theta.as<float>(), P.a_mag.as<float>(), P.nodes.as<UfNode>());
  AT_LAUNCH(at_theta_kernel, nA, nA, P.a_theta.as<float>(), P.a_mag.as<float>(), P.nodes.as<UfNode>(), P.parent.as<int>());
  // ---- 3: edges
  AT_LAUNCH(at_edges_words_kernel<false>, nA, P.apix.as<int>(), P.a_theta.as<float>(), P.amask.as<uint32_t>(), P.acount.as<int>(), wpr,
            nA, W, H, P.ecount.as<int>(), nullptr, nullptr, nullptr, nullptr);
  if (ezc_status st = exclusive_scan_i32(ctx, P.ecount.as<int>(), P.ecount.as<int>(), nA, P.scan_tmp, d_tot)) return st;
  int nE = 0;
  if (ezc_status st = read_int(ctx, d_tot, &nE, P)) return st;
  const size_t ecap = (size_t)std::max(nE, 1);
  EZC_CUDA(P.eA.reserve(sizeof(int) * ecap)); EZC_CUDA(P.eB.reserve(sizeof(int) * ecap)); EZC_CUDA(P.ecost.reserve(ecap));
  EZC_CUDA(P.order0.reserve(sizeof(uint32_t) * ecap)); EZC_CUDA(P.order1.reserve(sizeof(uint32_t) * ecap));
  EZC_CUDA(P.flags.reserve(sizeof(int) * (std::max(ecap, cap) + 1)));
  AT_LAUNCH(at_edges_words_kernel<true>, nA, P.apix.as<int>(), P.a_theta.as<float>(), P.amask.as<uint32_t>(), P.acount.as<int>(), wpr,
            nA, W, H, nullptr, P.ecount.as<int>(), P.eA.as<int>(), P.eB.as<int>(), P.ecost.as<uint8_t>());
  // ---- 4: connected components of the edge graph
  AT_LAUNCH(at_ccl_init_runs_kernel, nA, P.label.as<int>(), P.ecount.as<int>(), P.eB.as<int>(), nA, nE);
  AT_LAUNCH(at_ccl_union_kernel, nE, P.label.as<int>(), P.eA.as<int>(), P.eB.as<int>(), nE);
  AT_LAUNCH(at_ccl_flatten_kernel, nA, P.label.as<int>(), nA);
  // ---- 5: stable sort of the edges by (component, cost)
  uint32_t* o0 = P.order0.as<uint32_t>();
  uint32_t* o1 = P.order1.as<uint32_t>();
  // key-value LSD radix: one pass over the cost byte of the unsorted sequence, then the passes over the dense component
  // number (two 8-bit passes up to 65,536 components per sub-batch, 9-bit passes beyond)
  EZC_CUDA(P.keys0.reserve(sizeof(uint32_t) * ecap)); EZC_CUDA(P.keys1.reserve(sizeof(uint32_t) * ecap));
  uint32_t* k0 = P.keys0.as<uint32_t>();
  uint32_t* k1 = P.keys1.as<uint32_t>();
  // dense component numbers (see at_root_flag_kernel); P.rep / P.keep are free until stage 7
  int nComp = 0;
  if (nE > 0) {
    AT_LAUNCH(at_root_flag_kernel, nA, P.label.as<int>(), P.ecount.as<int>(), nA, nE, P.flags.as<int>());
    if (ezc_status st = exclusive_scan_i32(ctx, P.flags.as<int>(), P.rep.as<int>(), nA, P.scan_tmp, d_tot)) return st;
    if (ezc_status st = read_int(ctx, d_tot, &nComp, P)) return st;
    AT_LAUNCH(at_dense_label_kernel, nA, P.label.as<int>(), P.rep.as<int>(), nA, P.keep.as<int>());
  }
  AT_LAUNCH(at_edge_keys_kernel, nE, P.keep.as<int>(), P.eA.as<int>(), nE, k0);
  if (ezc_status st = radix_pass_kv<8, true>(ctx, k0, nullptr, P.ecost.as<uint8_t>(), k1, o1, nE, 0, P.sort)) return st;
  std::swap(o0, o1); std::swap(k0, k1);
  {
    const int bits = bits_for(std::max(nComp, 2));
    if (bits <= 16) {
      for (int shift = 0; shift < bits; shift += 8) {
        if (ezc_status st = radix_pass_kv<8, false>(ctx, k0, o0, nullptr, k1, o1, nE, shift, P.sort)) return st;
        std::swap(o0, o1); std::swap(k0, k1);
      }
    } else {
      for (int shift = 0; shift < bits; shift += 9) {
        if (ezc_status st = radix_pass_kv<9, false>(ctx, k0, o0, nullptr, k1, o1, nE, shift, P.sort)) return st;
        std::swap(o0, o1); std::swap(k0, k1);
      }
    }
  }
  // component starts, written directly: the sorted key is the component index and every index 0..nComp-1 occurs
  if (nE > 0) {
    EZC_CUDA(P.starts.reserve(sizeof(int) * ((size_t)nComp + 2)));
    AT_LAUNCH(at_starts_dense_kernel, nE, k0, nE, nComp, P.starts.as<int>());
    if (getenv("EZC_AT_STATS")) {   // component size distribution (edges per component) of this sub-batch
      std::vector<int> hs((size_t)nComp + 1);
      EZC_CUDA(cudaMemcpyAsync(hs.data(), P.starts.p, sizeof(int) * ((size_t)nComp + 1), cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaStreamSynchronize(ctx->stream));
      std::vector<int> sz((size_t)nComp);
      for (int c = 0; c < nComp; ++c) sz[c] = hs[c + 1] - hs[c];
      std::sort(sz.begin(), sz.end());
      long long tot = 0; for (int v : sz) tot += v;
      long long top = 0; for (int c = std::max(0, nComp - 1000); c < nComp; ++c) top += sz[c];
      fprintf(stderr, "[at] %d views: %d components, %lld edges; max %d, p99.9 %d, p99 %d, median %d; 1000 largest hold %.1f %% of the edges\n", nb, nComp, tot,
              sz.back(), sz[(size_t)(nComp * 0.999)], sz[(size_t)(nComp * 0.99)], sz[nComp / 2], 100.0 * top / tot);
    }
    // ---- 6: merge sweep
    {
      // persistent 4-warp blocks (16 per SM), components handed out through a global counter
      EZC_CUDA(P.work.reserve(64));
      EZC_CUDA(cudaMemsetAsync(P.work.p, 0, sizeof(int), ctx->stream));
      const int grid = (int)std::min<long long>(((long long)nComp + 3) / 4, (long long)ctx->sm_count * 16);
      at_merge_kernel<<<grid, 128, 0, ctx->stream>>>(o0, P.starts.as<int>(), nComp, P.eA.as<int>(), P.eB.as<int>(), P.nodes.as<UfNode>(),
                                                     P.parent.as<int>(), P.work.as<int>());
      EZC_CUDA(cudaGetLastError());
      ctx->kernel_launches++;
    }
  }
  if (dbg && nb == 1) {
    dbg->theta.resize(N); dbg->mag.resize(N);
    EZC_CUDA(cudaMemcpyAsync(dbg->theta.data(), P.dbg_theta.p, sizeof(float) * N, cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaMemcpyAsync(dbg->mag.data(), P.dbg_mag.p, sizeof(float) * N, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> apix(nA), eA(nE), eB(nE), par(nA), ssz(nA);
    std::vector<uint8_t> ec(nE);
    std::vector<uint32_t> ord(nE);
    if (nA) {
      EZC_CUDA(P.setsize.reserve(sizeof(int) * (size_t)nA));
      AT_LAUNCH(at_split_nodes_kernel, nA, P.nodes.as<UfNode>(), nA, P.setsize.as<int>());
      EZC_CUDA(cudaMemcpyAsync(apix.data(), P.apix.p, sizeof(int) * nA, cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaMemcpyAsync(par.data(), P.parent.p, sizeof(int) * nA, cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaMemcpyAsync(ssz.data(), P.setsize.p, sizeof(int) * nA, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (nE) {
      EZC_CUDA(cudaMemcpyAsync(eA.data(), P.eA.p, sizeof(int) * nE, cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaMemcpyAsync(eB.data(), P.eB.p, sizeof(int) * nE, cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaMemcpyAsync(ec.data(), P.ecost.p, nE, cudaMemcpyDeviceToHost, ctx->stream));
      EZC_CUDA(cudaMemcpyAsync(ord.data(), o0, sizeof(uint32_t) * nE, cudaMemcpyDeviceToHost, ctx->stream));
    }
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    dbg->edges.resize((size_t)nE * 3);
    for (int i = 0; i < nE; ++i) { const uint32_t e = ord[i]; dbg->edges[3 * i] = ec[e]; dbg->edges[3 * i + 1] = apix[eA[e]]; dbg->edges[3 * i + 2] = apix[eB[e]]; }
    dbg->rep.assign(N, -1); dbg->size.assign(N, 1);
    for (int i = 0; i < nA; ++i) { int r = i; while (par[r] != r) r = par[r]; dbg->rep[apix[i]] = apix[r]; dbg->size[apix[i]] = ssz[r]; }
  }
  // ---- 7: clusters = pixels grouped by representative (ascending), raster order inside
  AT_LAUNCH(at_rep_kernel, nA, P.nodes.as<UfNode>(), P.parent.as<int>(), nA, P.rep.as<int>(), P.keep.as<int>());
  if (ezc_status st = exclusive_scan_i32(ctx, P.keep.as<int>(), P.keep.as<int>(), nA, P.scan_tmp, d_tot)) return st;
  int nK = 0;
  if (ezc_status st = read_int(ctx, d_tot, &nK, P)) return st;
  const size_t kcap = (size_t)std::max(nK, 1);
  EZC_CUDA(P.plist0.reserve(sizeof(uint32_t) * kcap)); EZC_CUDA(P.plist1.reserve(sizeof(uint32_t) * kcap));
  uint32_t* l0 = P.plist0.as<uint32_t>();
  uint32_t* l1 = P.plist1.as<uint32_t>();
  AT_LAUNCH(at_keep_compact_kernel, nA, P.keep.as<int>(), P.rep.as<int>(), P.nodes.as<UfNode>(), nA, l0);
  for (int shift = 0; shift < bits_for(std::max(nA, 2)); shift += 8) {
    if (ezc_status st = radix_pass(ctx, l0, l1, nK, DigitArray{P.rep.as<int>(), shift}, P.sort)) return st;
    std::swap(l0, l1);
  }
  int nC = 0;
  if (nK > 0) {
    EZC_CUDA(P.flags.reserve(sizeof(int) * (kcap + 1)));
    AT_LAUNCH(at_boundary_kernel, nK, l0, nK, P.rep.as<int>(), (const int*)nullptr, P.flags.as<int>());
    if (ezc_status st = exclusive_scan_i32(ctx, P.flags.as<int>(), P.flags.as<int>(), nK, P.scan_tmp, d_tot)) return st;
    if (ezc_status st = read_int(ctx, d_tot, &nC, P)) return st;
    EZC_CUDA(P.cstart.reserve(sizeof(int) * ((size_t)nC + 2)));
    AT_LAUNCH(at_starts_kernel, nK, P.flags.as<int>(), d_tot, nK, P.cstart.as<int>());
  }
  // ---- 8: line fits -> segments
  const size_t ccap = (size_t)std::max(nC, 1);
  EZC_CUDA(P.segs_raw.reserve(sizeof(Segment) * ccap)); EZC_CUDA(P.seg_valid.reserve(sizeof(int) * (ccap + 1)));
  EZC_CUDA(P.img_count.reserve(sizeof(int) * ((size_t)nb + 1))); EZC_CUDA(P.seg_start.reserve(sizeof(int) * ((size_t)nb + 2)));
  AT_LAUNCH(at_fit_kernel, nC, l0, P.cstart.as<int>(), nC, P.apix.as<int>(), P.a_theta.as<float>(), P.a_mag.as<float>(), W, H,
            P.segs_raw.as<Segment>(), P.seg_valid.as<int>());
  EZC_CUDA(P.keep.reserve(sizeof(int) * (ccap + 1)));
  int* vscan = P.keep.as<int>();
  if (ezc_status st = exclusive_scan_i32(ctx, P.seg_valid.as<int>(), vscan, nC, P.scan_tmp, d_tot)) return st;
  int nS = 0;
  if (ezc_status st = read_int(ctx, d_tot, &nS, P)) return st;
  const size_t scap = (size_t)std::max(nS, 1);
  EZC_CUDA(P.segs.reserve(sizeof(Segment) * scap));
  EZC_CUDA(cudaMemsetAsync(P.img_count.p, 0, sizeof(int) * (nb + 1), ctx->stream));
  AT_LAUNCH(at_seg_compact_kernel, nC, P.segs_raw.as<Segment>(), P.seg_valid.as<int>(), vscan, nC, P.segs.as<Segment>(), P.img_count.as<int>());
  // ---- 9a: spatial hash + children
  const GridDims gd = grid_dims(W, H);
  const int n_cells = gd.gw * gd.gh * nb;
  EZC_CUDA(P.cell_of.reserve(sizeof(int) * scap)); EZC_CUDA(P.cell_count.reserve(sizeof(int) * ((size_t)n_cells + 1)));
  EZC_CUDA(P.cell_start.reserve(sizeof(int) * ((size_t)n_cells + 1))); EZC_CUDA(P.cell_fill.reserve(sizeof(int) * ((size_t)n_cells + 1)));
  EZC_CUDA(P.cell_items.reserve(sizeof(int) * scap)); EZC_CUDA(P.ccount.reserve(sizeof(int) * (scap + 1)));
  EZC_CUDA(P.qcount.reserve(sizeof(int) * (scap + 1)));
  EZC_CUDA(cudaMemsetAsync(P.cell_count.p, 0, sizeof(int) * (n_cells + 1), ctx->stream));
  EZC_CUDA(cudaMemsetAsync(P.cell_fill.p, 0, sizeof(int) * (n_cells + 1), ctx->stream));
  AT_LAUNCH(at_cell_count_kernel, nS, P.segs.as<Segment>(), nS, W, H, P.cell_of.as<int>(), P.cell_count.as<int>());
  if (ezc_status st = exclusive_scan_i32(ctx, P.cell_count.as<int>(), P.cell_start.as<int>(), n_cells, P.scan_tmp, nullptr)) return st;
  AT_LAUNCH(at_cell_fill_kernel, nS, P.cell_of.as<int>(), nS, P.cell_start.as<int>(), P.cell_fill.as<int>(), P.cell_items.as<int>());
  AT_LAUNCH(at_cell_sort_kernel, n_cells, P.cell_start.as<int>(), P.cell_count.as<int>(), n_cells, P.cell_items.as<int>());
  AT_LAUNCH(at_children_kernel<false>, nS, P.segs.as<Segment>(), nS, W, H, P.cell_start.as<int>(), P.cell_count.as<int>(),
            P.cell_items.as<int>(), P.ccount.as<int>(), nullptr, nullptr);
  EZC_CUDA(cudaMemsetAsync(P.ccount.as<int>() + nS, 0, sizeof(int), ctx->stream));
  if (ezc_status st = exclusive_scan_i32(ctx, P.ccount.as<int>(), P.ccount.as<int>(), nS + 1, P.scan_tmp, d_tot)) return st;
  int nCh = 0;
  if (ezc_status st = read_int(ctx, d_tot, &nCh, P)) return st;
  EZC_CUDA(P.children.reserve(sizeof(int) * (size_t)std::max(nCh, 1)));
  AT_LAUNCH(at_children_kernel<true>, nS, P.segs.as<Segment>(), nS, W, H, P.cell_start.as<int>(), P.cell_count.as<int>(),
            P.cell_items.as<int>(), nullptr, P.ccount.as<int>(), P.children.as<int>());
  // ---- 9b: quads
  AT_LAUNCH(at_quads_kernel<false>, nS, P.segs.as<Segment>(), nS, P.ccount.as<int>(), P.children.as<int>(), P.qcount.as<int>(), nullptr, nullptr);
  if (ezc_status st = exclusive_scan_i32(ctx, P.qcount.as<int>(), P.qcount.as<int>(), nS, P.scan_tmp, d_tot)) return st;
  int nQ = 0;
  if (ezc_status st = read_int(ctx, d_tot, &nQ, P)) return st;
  const size_t qcap = (size_t)std::max(nQ, 1);
  EZC_CUDA(P.quads.reserve(sizeof(QuadRec) * qcap)); EZC_CUDA(P.dets.reserve(sizeof(Detection) * qcap));
  EZC_CUDA(P.qimg_count.reserve(sizeof(int) * ((size_t)nb + 1))); EZC_CUDA(P.qstart.reserve(sizeof(int) * ((size_t)nb + 2)));
  EZC_CUDA(P.ngood.reserve(sizeof(int) * ((size_t)nb + 1)));
  EZC_CUDA(P.corners.reserve(sizeof(float2) * qcap * 4)); EZC_CUDA(P.cimg.reserve(sizeof(int) * qcap * 4));
  EZC_CUDA(P.refine_flag.reserve(sizeof(int) * qcap));
  AT_LAUNCH(at_quads_kernel<true>, nS, P.segs.as<Segment>(), nS, P.ccount.as<int>(), P.children.as<int>(), nullptr, P.qcount.as<int>(),
            P.quads.as<QuadRec>());
  // ---- 9c: decode
  AT_LAUNCH(at_decode_kernel, nQ, P.quads.as<QuadRec>(), nQ, imgs->d, imgs->pitch, imgs->image_stride, W, H, img0, P.dets.as<Detection>());
  // ---- 9d: per-image ranges, de-dup, sub-pixel refinement, scatter
  EZC_CUDA(cudaMemsetAsync(P.qimg_count.p, 0, sizeof(int) * (nb + 1), ctx->stream));
  AT_LAUNCH(at_quad_image_count_kernel, nQ, P.quads.as<QuadRec>(), nQ, P.qimg_count.as<int>());
  at_image_starts_kernel<<<1, 1, 0, ctx->stream>>>(P.qimg_count.as<int>(), nb, P.qstart.as<int>());
  ctx->kernel_launches++;
  AT_LAUNCH(at_dedup_kernel, nb, P.dets.as<Detection>(), P.qstart.as<int>(), nb, P.ngood.as<int>());
  EZC_CUDA(cudaMemsetAsync(P.refine_flag.p, 0, sizeof(int) * qcap, ctx->stream));
  if (nQ > 0) {
    // the slots of non-surviving detections are never read, but cornerSubPix runs over the whole array:
    // give them a harmless in-image position
    EZC_CUDA(cudaMemsetAsync(P.corners.p, 0, sizeof(float2) * qcap * 4, ctx->stream));
    EZC_CUDA(cudaMemsetAsync(P.cimg.p, 0, sizeof(int) * qcap * 4, ctx->stream));
    at_gather_corners_kernel<<<nb, 64, 0, ctx->stream>>>(P.dets.as<Detection>(), P.qstart.as<int>(), P.ngood.as<int>(), nb, nb_tags, W, H, img0,
                                                          P.corners.as<float2>(), P.cimg.as<int>(), P.refine_flag.as<int>());
    EZC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    // aprilgrid_detector.hpp:52-53
    if (ezc_status st = subpix_device(ctx, imgs, P.corners.as<float2>(), P.cimg.as<int>(), nQ * 4, 5, 5, 100, 0.05)) return st;
  }
  at_scatter_kernel<<<nb, 64, 0, ctx->stream>>>(P.dets.as<Detection>(), P.qstart.as<int>(), P.ngood.as<int>(), P.corners.as<float2>(),
                                                 P.refine_flag.as<int>(), nb, nb_tags, d_out_corners + (size_t)img0 * 4 * nb_tags,
                                                 d_out_ndet + img0, d_out_success + img0);
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  if (dbg && nb == 1) {
    std::vector<Segment> hs(nS);
    std::vector<QuadRec> hq(nQ);
    if (nS) EZC_CUDA(cudaMemcpyAsync(hs.data(), P.segs.p, sizeof(Segment) * nS, cudaMemcpyDeviceToHost, ctx->stream));
    if (nQ) EZC_CUDA(cudaMemcpyAsync(hq.data(), P.quads.p, sizeof(QuadRec) * nQ, cudaMemcpyDeviceToHost, ctx->stream));
    EZC_CUDA(cudaStreamSynchronize(ctx->stream));
    dbg->segments.resize((size_t)nS * 6);
    for (int i = 0; i < nS; ++i) { float* o = &dbg->segments[(size_t)i * 6]; o[0] = hs[i].x0; o[1] = hs[i].y0; o[2] = hs[i].x1; o[3] = hs[i].y1; o[4] = hs[i].theta; o[5] = hs[i].length; }
    dbg->quads.resize((size_t)nQ * 9);
    for (int i = 0; i < nQ; ++i) { float* o = &dbg->quads[(size_t)i * 9]; for (int k = 0; k < 4; ++k) { o[2 * k] = hq[i].p[k][0]; o[2 * k + 1] = hq[i].p[k][1]; } o[8] = hq[i].perimeter; }
  }
  return EZC_OK;
}

void init_filter() {
  // Gaussian::makeGaussianFilter(0.8f, 3)  (Gaussian.cc:8-31)
  const float sigma = 0.8f;
  const int n = 3;
  float f[3];
  const double inv_variance = 1. / (2 * sigma * sigma);
  float sum = 0;
  for (int i = 0; i < n; i++) {
    const int j = i - n / 2;
    f[i] = (float)std::exp(-j * j * inv_variance);
    sum += f[i];
  }
  for (int i = 0; i < n; i++) f[i] /= sum;
  cudaMemcpyToSymbol(c_filt, f, sizeof(f));
  float lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = (float)((double)v / 255.);
  cudaMemcpyToSymbol(c_u8f, lut, sizeof(lut));
}

}  // namespace

extern "C" {

ezc_status ezc_detect_aprilgrid_batch(ezc_context* ctx, const ezc_images* imgs, int32_t nb_rows, int32_t nb_cols,
                                      float* corners_out, int32_t* n_detections_out, uint8_t* success_out,
                                      int32_t max_images_in_flight) {
  EZC_CHECK(ctx && imgs && corners_out && n_detections_out && success_out, EZC_ERR_INVALID, "ezc_detect_aprilgrid_batch: null argument");
  EZC_CHECK(nb_rows > 0 && nb_cols > 0, EZC_ERR_INVALID, "ezc_detect_aprilgrid_batch: bad grid size");
  EZC_CHECK(imgs->width >= 4 && imgs->height >= 4, EZC_ERR_INVALID, "ezc_detect_aprilgrid_batch: image too small");
  EZC_CUDA(cudaSetDevice(ctx->device));
  const int nb_tags = nb_rows * nb_cols;
  const int n = imgs->n;
  if (n == 0) return EZC_OK;
  init_filter();
  const long long N = (long long)imgs->width * imgs->height;
  // sub-batch: dense indices must stay below 2^31 and the working set (8 B/px sparse theta/mag + ~8 B/px compact pools and
  // edge lists) bounded to ~16 GiB of the 180 GB; the per-component sweeps are latency bound, so more images in flight
  // = more independent components = better fill (35 -> 140 images: +35 % throughput)
  int sub = (int)std::max<long long>(1, std::min<long long>((1LL << 30) / N, 16LL * 1024 * 1024 * 1024 / (N * 16)));
  if (max_images_in_flight > 0) sub = std::min(sub, max_images_in_flight);
  sub = std::min(sub, n);
  Pools& P = get_pools(ctx);
  ezc::DevBuf& d_corners = P.out_corners; ezc::DevBuf& d_ndet = P.out_ndet; ezc::DevBuf& d_succ = P.out_succ;
  EZC_CUDA(d_corners.reserve(sizeof(float2) * (size_t)n * 4 * nb_tags));
  EZC_CUDA(d_ndet.reserve(sizeof(int) * (size_t)n));
  EZC_CUDA(d_succ.reserve((size_t)n));
  // images still crossing PCIe (ezc_images_upload_async): start with a small sub-batch so the pipeline fills quickly
  int cur = imgs->ready.empty() ? sub : std::min(sub, std::max(1, 2 * imgs->chunk));
  for (int i0 = 0; i0 < n;) {
    const int nb = std::min(cur, n - i0);
    if (ezc_status st = process_sub_batch(ctx, imgs, i0, nb, nb_tags, P, d_corners.as<float2>(), d_ndet.as<int>(), d_succ.as<uint8_t>(), nullptr))
      return st;
    i0 += nb;
    cur = std::min(sub, cur * 2);
  }
  EZC_CUDA(cudaMemcpyAsync(corners_out, d_corners.p, sizeof(float2) * (size_t)n * 4 * nb_tags, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(n_detections_out, d_ndet.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(success_out, d_succ.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}

// Test hook: run ONE image and return the stage products (counts first, then data through getters).
ezc_status ezc_apriltag_debug_run(ezc_context* ctx, const ezc_images* imgs, int32_t image, int32_t nb_rows, int32_t nb_cols,
                                  ezc_apriltag_debug** out, float* corners_out, int32_t* n_detections_out, uint8_t* success_out) {
  EZC_CHECK(ctx && imgs && out && image >= 0 && image < imgs->n, EZC_ERR_INVALID, "ezc_apriltag_debug_run: bad argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  init_filter();
  const int nb_tags = nb_rows * nb_cols;
  Pools P;
  ezc::DevBuf d_corners, d_ndet, d_succ;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  EZC_CUDA(d_corners.reserve(sizeof(float2) * (size_t)imgs->n * 4 * nb_tags));
  EZC_CUDA(d_ndet.reserve(sizeof(int) * (size_t)imgs->n));
  EZC_CUDA(d_succ.reserve((size_t)imgs->n));
  ezc_apriltag_debug* d = new ezc_apriltag_debug();
  ezc_status st = process_sub_batch(ctx, imgs, image, 1, nb_tags, P, d_corners.as<float2>(), d_ndet.as<int>(), d_succ.as<uint8_t>(), d);
  if (st != EZC_OK) { delete d; return st; }
  if (corners_out) EZC_CUDA(cudaMemcpyAsync(corners_out, d_corners.as<float2>() + (size_t)image * 4 * nb_tags, sizeof(float2) * 4 * nb_tags, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_detections_out) EZC_CUDA(cudaMemcpyAsync(n_detections_out, d_ndet.as<int>() + image, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (success_out) EZC_CUDA(cudaMemcpyAsync(success_out, d_succ.as<uint8_t>() + image, 1, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = d;
  return EZC_OK;
}
int32_t ezc_apriltag_debug_count(const ezc_apriltag_debug* d, int32_t what) {
  if (!d) return 0;
  switch (what) {
    case 0: return (int32_t)d->theta.size();
    case 1: return (int32_t)(d->edges.size() / 3);
    case 2: return (int32_t)(d->segments.size() / 6);
    case 3: return (int32_t)(d->quads.size() / 9);
    default: return 0;
  }
}
ezc_status ezc_apriltag_debug_get(const ezc_apriltag_debug* d, float* theta, float* mag, int32_t* edges, int32_t* rep, int32_t* size,
                                  float* segments, float* quads) {
  EZC_CHECK(d, EZC_ERR_INVALID, "ezc_apriltag_debug_get: null");
  if (theta) std::memcpy(theta, d->theta.data(), sizeof(float) * d->theta.size());
  if (mag) std::memcpy(mag, d->mag.data(), sizeof(float) * d->mag.size());
  if (edges) std::memcpy(edges, d->edges.data(), sizeof(int) * d->edges.size());
  if (rep) std::memcpy(rep, d->rep.data(), sizeof(int) * d->rep.size());
  if (size) std::memcpy(size, d->size.data(), sizeof(int) * d->size.size());
  if (segments) std::memcpy(segments, d->segments.data(), sizeof(float) * d->segments.size());
  if (quads) std::memcpy(quads, d->quads.data(), sizeof(float) * d->quads.size());
  return EZC_OK;
}
void ezc_apriltag_debug_free(ezc_apriltag_debug* d) { delete d; }

ezc_status ezc_debug_libm(ezc_context* ctx, int32_t which, const float* a, const float* b, int64_t n, float* out) {
  EZC_CHECK(ctx && a && out && n >= 0 && which >= 0 && which <= 2 && (which != 0 || b), EZC_ERR_INVALID, "ezc_debug_libm: bad argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return EZC_OK;
  ezc::DevBuf da, db, dout;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  EZC_CUDA(da.reserve(sizeof(float) * n)); EZC_CUDA(dout.reserve(sizeof(float) * n));
  EZC_CUDA(cudaMemcpyAsync(da.p, a, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (which == 0) {
    EZC_CUDA(db.reserve(sizeof(float) * n));
    EZC_CUDA(cudaMemcpyAsync(db.p, b, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  ezc::at_libm_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 32), 256, 0, ctx->stream>>>(which, da.as<float>(), db.as<float>(), n,
                                                                                                    dout.as<float>());
  EZC_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  EZC_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}

}  // extern "C"
