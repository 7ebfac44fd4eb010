// cb_rows.cuh -- byte-image passes of the chessboard detector on 16-pixel row chunks.
//
// The detector's image-level passes (threshold, 3x3 min / max, run heads of the labelling, run end points of the gate)
// move 1-2 bytes per pixel; with one thread per pixel they issue 32-byte requests per warp and sit at 5 % of the HBM
// rate.  Here one thread owns 16 consecutive pixels of a row: one 128-bit load per row when the row start is 16-byte
// aligned (W, pitch and image stride multiples of 16 -- every camera format of the reference's configs), byte loads
// otherwise (same code, same results), SIMD-in-a-word byte arithmetic, and 16-bit masks for the run logic.
#pragma once
#include <cstdint>

namespace ezc {

struct Chunk16 { uint32_t w[4]; };

// grid.x covers the chunks of one image in raster order, grid.y the image slots
struct ChunkPos { int y, x0; bool ok; };
__device__ __forceinline__ ChunkPos chunk_pos(int W, int H) {
  const int cpr = (W + 15) >> 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  ChunkPos c;
  c.y = idx / cpr;
  c.x0 = (idx - c.y * cpr) << 4;
  c.ok = c.y < H;
  return c;
}
inline unsigned chunk_blocks(int W, int H) { return (unsigned)(((long long)((W + 15) >> 4) * H + 255) / 256); }

// pixels x0 .. x0+15 of `row`; pixels at x >= W read as `fill`
__device__ __forceinline__ Chunk16 ld16(const uint8_t* __restrict__ row, int x0, int W, bool vec, uint32_t fill) {
  Chunk16 c;
  if (vec) {
    const uint4 v = *reinterpret_cast<const uint4*>(row + x0);
    c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int x = x0 + 4 * k + b;
        w |= (x < W ? (uint32_t)row[x] : fill) << (8 * b);
      }
      c.w[k] = w;
    }
  }
  return c;
}
__device__ __forceinline__ void st16(uint8_t* __restrict__ row, int x0, int W, bool vec, const Chunk16& c) {
  if (vec) {
    *reinterpret_cast<uint4*>(row + x0) = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int x = x0 + 4 * k + b;
        if (x < W) row[x] = (uint8_t)(c.w[k] >> (8 * b));
      }
  }
}
// bit k = pixel k of the chunk is non-zero
__device__ __forceinline__ uint32_t nz16(const Chunk16& c) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t b = __vcmpne4(c.w[k], 0u) & 0x01010101u;
    m |= ((b * 0x01020408u) >> 24) << (4 * k);   // gathers the four byte flags into a nibble (no carries: all partial products are distinct powers of two)
  }
  return m;
}
// bits of the chunk's pixels that lie inside the row
__device__ __forceinline__ uint32_t valid16(int x0, int W) { return W - x0 >= 16 ? 0xffffu : ((1u << (W - x0)) - 1u); }

// out = thresh > 0 ? (v >= thresh ? 255 : 0) : v   (icvBinarizationHistogramBased leaves the image alone without a threshold)
__global__ void cb_binarize16_kernel(const uint8_t* __restrict__ img, long long pitch, long long istride, bool vec_src, int W, int H,
                                     const int* __restrict__ thresh, const int* __restrict__ active, uint8_t* __restrict__ out, bool vec_dst) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  Chunk16 v = ld16(img + (long long)im * istride + (long long)c.y * pitch, c.x0, W, vec_src, 0u);
  const int t = thresh[im];
  if (t > 0) {
    const uint32_t t4 = (uint32_t)(t > 255 ? 255 : t) * 0x01010101u;
#pragma unroll
    for (int k = 0; k < 4; ++k) v.w[k] = t > 255 ? 0u : __vcmpgeu4(v.w[k], t4);
  }
  st16(out + ((long long)im * H + c.y) * W, c.x0, W, vec_dst, v);
}

// dst = inv ? (src <= t) : (src > t), 255 / 0   (cv::threshold THRESH_BINARY_INV / THRESH_BINARY), packed images
__global__ void cb_fg16_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H, const int* __restrict__ active,
                               int thresh, int inv, bool vec) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  const long long off = ((long long)im * H + c.y) * W;
  Chunk16 v = ld16(src + off, c.x0, W, vec, 0u);
  const uint32_t t4 = (uint32_t)(thresh < 0 ? 0 : (thresh > 255 ? 255 : thresh)) * 0x01010101u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t above = thresh < 0 ? 0xffffffffu : (thresh >= 255 ? 0u : __vcmpgtu4(v.w[k], t4));
    v.w[k] = inv ? ~above : above;
  }
  st16(dst + off, c.x0, W, vec, v);
}

// cv::erode / cv::dilate(src, dst, Mat(), Point(-1,-1), 1): 3x3 min / max, pixels outside the image are ignored.
// Vertical pass on the chunk and on its two neighbour columns, then the horizontal pass with byte shifts across words.
template <bool IS_MAX>
__global__ void cb_morph16_kernel(const uint8_t* __restrict__ src, long long pitch, long long istride, bool vec_src, uint8_t* __restrict__ dst,
                                  int W, int H, const int* __restrict__ active, bool vec_dst, const int* __restrict__ src_of = nullptr) {
  const ChunkPos c = chunk_pos(W, H);
  if (!c.ok) return;
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  const uint8_t* S = src + (long long)(src_of ? src_of[im] : im) * istride;   // src_of: the slot reads another image's source
  const uint32_t idb = IS_MAX ? 0u : 0xffu;
  auto op = [](uint32_t a, uint32_t b) { return IS_MAX ? __vmaxu4(a, b) : __vminu4(a, b); };
  Chunk16 m;
#pragma unroll
  for (int k = 0; k < 4; ++k) m.w[k] = idb * 0x01010101u;
  uint32_t ml = idb, mr = idb;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = c.y + dy;
    if (yy < 0 || yy >= H) continue;
    const uint8_t* row = S + (long long)yy * pitch;
    const Chunk16 r = ld16(row, c.x0, W, vec_src, idb);
#pragma unroll
    for (int k = 0; k < 4; ++k) m.w[k] = op(m.w[k], r.w[k]);
    if (c.x0 > 0) ml = op(ml, (uint32_t)row[c.x0 - 1]) & 0xffu;
    if (c.x0 + 16 < W) mr = op(mr, (uint32_t)row[c.x0 + 16]) & 0xffu;
  }
  Chunk16 o;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t prev = k == 0 ? (ml << 24) : m.w[k - 1];
    const uint32_t next = k == 3 ? mr : m.w[k + 1];
    const uint32_t left = __funnelshift_l(prev, m.w[k], 8);    // pixel x-1 at the position of pixel x
    const uint32_t right = __funnelshift_r(m.w[k], next, 8);   // pixel x+1 at the position of pixel x
    o.w[k] = op(m.w[k], op(left, right));
  }
  st16(dst + ((long long)im * H + c.y) * W, c.x0, W, vec_dst, o);
}

// rectangle(img, (0,0), (W-1,H-1), 255, 3, LINE_8): a 3-pixel white frame (checked against cv2); border pixels only
__global__ void cb_frame_border_kernel(uint8_t* __restrict__ img, int W, int H, const int* __restrict__ active) {
  const int im = active ? active[blockIdx.y] : (int)blockIdx.y;
  uint8_t* I = img + (long long)im * W * H;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_h = 6 * W;   // three rows at the top and at the bottom
  if (i < n_h) {
    const int r = i / W, x = i - r * W;
    const int y = r < 3 ? r : H - 6 + r;
    if (y >= 0 && y < H) I[(long long)y * W + x] = 255;
  } else if (i < n_h + 6 * H) {
    const int j = i - n_h;
    const int y = j / 6, k = j - y * 6;
    const int x = k < 3 ? k : W - 6 + k;
    if (x >= 0 && x < W) I[(long long)y * W + x] = 255;
  }
}

}  // namespace ezc
