// subpix.cu -- batched cv::cornerSubPix on the GPU.
//
// Replaces the calls at /root/reference/include/target_detectors/chessboard_detector.hpp:39-40 and
// aprilgrid_detector.hpp:52-53:  cv::cornerSubPix(img, pts, Size(5,5), Size(-1,-1), {COUNT+EPS,100,0.05}),
// plus the (2,2)/{15,0.1} refinement inside cv::findChessboardCorners.
//
// One warp per corner.  Every iteration the warp (1) samples the (2w+3)x(2h+3) float32 patch around the
// current estimate with exactly the arithmetic of OpenCV's getRectSubPix_8u32f (its "prev/t" recurrence in
// the fully-inside case, the replicated-border formulas otherwise; no FMA contraction), (2) accumulates the
// five gradient moments in FP64 in the reference's sequential order (no FMA contraction: -fmad=false), (3) solves the 2x2 system and rounds the
// estimate to float32 like cv::Point2f does.
#include <cmath>

#include "images.cuh"

namespace ezc {

constexpr int kMaxWin = 5;                           // window half-size supported
constexpr int kMaxPatch = (2 * kMaxWin + 3);         // 13
constexpr int kSubpixWarps = 4;
constexpr int kMaxTerms = (2 * kMaxWin + 1) * (2 * kMaxWin + 1);  // 121

struct SubpixArgs {
  const uint8_t* imgs;
  long long pitch, image_stride;
  int W, H;
  float2* corners;
  const int32_t* image_index;
  int n;
  int ww, wh;
  int max_iters;
  float eps2;
  float mask[(2 * kMaxWin + 1) * (2 * kMaxWin + 1)];  // exp(-y^2) * exp(-x^2), computed on the host with std::exp
};

__global__ void __launch_bounds__(kSubpixWarps * 32) corner_subpix_kernel(const SubpixArgs a) {
  __shared__ float patch_s[kSubpixWarps][kMaxPatch * kMaxPatch];
  __shared__ double terms_s[kSubpixWarps][5 * kMaxTerms];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * kSubpixWarps + warp;
  if (idx >= a.n) return;
  float* patch = patch_s[warp];
  double* terms = terms_s[warp];
  const uint8_t* img = a.imgs + (long long)a.image_index[idx] * a.image_stride;
  const int W = a.W, H = a.H;
  const long long pitch = a.pitch;
  const int pw = 2 * a.ww + 3, ph = 2 * a.wh + 3;
  const int iw = 2 * a.ww + 1, ih = 2 * a.wh + 1;
  const float2 cT = a.corners[idx];
  float2 cI = cT;
  int iter = 0;
  float err = 0.f;
  do {
    // ---- getRectSubPix(src, Size(pw, ph), cI) -> patch (float32)
    const float centre_x = __fsub_rn(cI.x, (float)(pw - 1) * 0.5f);
    const float centre_y = __fsub_rn(cI.y, (float)(ph - 1) * 0.5f);
    const int ipx = (int)floorf(centre_x), ipy = (int)floorf(centre_y);
    __syncwarp();
    if (0 <= ipx && ipx + pw < W && 0 <= ipy && ipy + ph < H) {
      float fa = __fsub_rn(centre_x, (float)ipx);
      const float fb = __fsub_rn(centre_y, (float)ipy);
      fa = fmaxf(fa, 0.0001f);
      const float b1 = __fsub_rn(1.f, fb), b2 = fb;
      const float a12 = __fmul_rn(fa, b1), a22 = __fmul_rn(fa, fb);
      const double s = (1.0 - (double)fa) / (double)fa;
      const float one_m_a = __fsub_rn(1.f, fa);
      for (int p = lane; p < pw * ph; p += 32) {
        const int i = p / pw, j = p - i * pw;
        const uint8_t* r0 = img + (long long)(ipy + i) * pitch + ipx;
        const uint8_t* r1 = r0 + pitch;
        const float t = __fadd_rn(__fmul_rn(a12, (float)r0[j + 1]), __fmul_rn(a22, (float)r1[j + 1]));
        float prev;
        if (j == 0) {
          prev = __fmul_rn(one_m_a, __fadd_rn(__fmul_rn(b1, (float)r0[0]), __fmul_rn(b2, (float)r1[0])));
        } else {
          const float tp = __fadd_rn(__fmul_rn(a12, (float)r0[j]), __fmul_rn(a22, (float)r1[j]));
          prev = (float)((double)tp * s);
        }
        patch[p] = __fadd_rn(prev, t);
      }
    } else {
      // replicated border: getRectSubPix_Cn_<uchar,float,float> + adjustRect
      const float fa = __fsub_rn(centre_x, (float)ipx);
      const float fb = __fsub_rn(centre_y, (float)ipy);
      const float a11 = __fmul_rn(__fsub_rn(1.f, fa), __fsub_rn(1.f, fb));
      const float a12 = __fmul_rn(fa, __fsub_rn(1.f, fb));
      const float a21 = __fmul_rn(__fsub_rn(1.f, fa), fb);
      const float a22 = __fmul_rn(fa, fb);
      const float b1 = __fsub_rn(1.f, fb), b2 = fb;
      int col0, rx, rw, row0, ry, rh;
      if (ipx >= 0) { col0 = ipx; rx = 0; } else { col0 = 0; rx = min(-ipx, pw); }
      if (ipx < W - pw) rw = pw; else { rw = W - ipx - 1; if (rw < 0) { col0 += rw; rw = 0; } }
      if (ipy >= 0) { row0 = ipy; ry = 0; } else { row0 = 0; ry = -ipy; }
      if (ipy < H - ph) rh = ph; else { rh = H - ipy - 1; if (rh < 0) { row0 += rh; rh = 0; } }
      const int base = col0 - rx;
      for (int p = lane; p < pw * ph; p += 32) {
        const int i = p / pw, j = p - i * pw;
        // row pointer after i steps of the reference loop: advances once per row while i < rh, but not
        // while i < ry (src2 == src there)
        int row = row0 + max(0, min(i, rh) - min(ry, min(i, rh)));
        int row2 = (i < ry || i >= rh) ? row : row + 1;
        row = min(max(row, 0), H - 1);
        row2 = min(max(row2, 0), H - 1);
        const uint8_t* r0 = img + (long long)row * pitch;
        const uint8_t* r1 = img + (long long)row2 * pitch;
        float v;
        if (j < rx) {
          const int c = min(max(base + rx, 0), W - 1);
          v = __fadd_rn(__fmul_rn((float)r0[c], b1), __fmul_rn((float)r1[c], b2));
        } else if (j >= rw) {
          const int c = min(max(base + rw, 0), W - 1);
          v = __fadd_rn(__fmul_rn((float)r0[c], b1), __fmul_rn((float)r1[c], b2));
        } else {
          const int c0 = min(max(base + j, 0), W - 1), c1 = min(max(base + j + 1, 0), W - 1);
          v = __fadd_rn(__fadd_rn(__fmul_rn((float)r0[c0], a11), __fmul_rn((float)r0[c1], a12)), __fmul_rn((float)r1[c0], a21));
          v = __fadd_rn(v, __fmul_rn((float)r1[c1], a22));
        }
        patch[p] = v;
      }
    }
    __syncwarp();
    // ---- gradient moments over the inner iw x ih window
    // products in parallel (one element per lane and pass), sums in the reference's SEQUENTIAL raster order
    // (cornersubpix.cpp: a += gxx; b += gxy; ... inside the i/j loops): lanes 0..4 each own one accumulator.
    // The TU is compiled with -fmad=false, so no product is contracted into the following add.
    for (int q = lane; q < iw * ih; q += 32) {
      const int i = q / iw, j = q - i * iw;
      const float* sp = patch + (i + 1) * pw + (j + 1);
      const double m = (double)a.mask[q];
      const double tgx = (double)__fsub_rn(sp[1], sp[-1]);
      const double tgy = (double)__fsub_rn(sp[pw], sp[-pw]);
      const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
      const double px = (double)(j - a.ww), py = (double)(i - a.wh);
      terms[0 * kMaxTerms + q] = gxx;
      terms[1 * kMaxTerms + q] = gxy;
      terms[2 * kMaxTerms + q] = gyy;
      terms[3 * kMaxTerms + q] = gxx * px + gxy * py;
      terms[4 * kMaxTerms + q] = gxy * px + gyy * py;
    }
    __syncwarp();
    double acc = 0;
    if (lane < 5) {
      const double* tp = terms + lane * kMaxTerms;
      for (int q = 0; q < iw * ih; ++q) acc += tp[q];
    }
    const double sa = __shfl_sync(0xffffffffu, acc, 0), sb = __shfl_sync(0xffffffffu, acc, 1), sc = __shfl_sync(0xffffffffu, acc, 2);
    const double sbb1 = __shfl_sync(0xffffffffu, acc, 3), sbb2 = __shfl_sync(0xffffffffu, acc, 4);
    const double det = sa * sc - sb * sb;
    if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
    const double scale = 1.0 / det;
    float2 cI2;
    cI2.x = (float)((double)cI.x + sc * scale * sbb1 - sb * scale * sbb2);
    cI2.y = (float)((double)cI.y - sb * scale * sbb1 + sa * scale * sbb2);
    const float dx = __fsub_rn(cI2.x, cI.x), dy = __fsub_rn(cI2.y, cI.y);
    err = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    cI = cI2;
    if (cI.x < 0 || cI.x >= (float)W || cI.y < 0 || cI.y >= (float)H) break;
  } while (++iter < a.max_iters && err > a.eps2);
  if (fabsf(__fsub_rn(cI.x, cT.x)) > (float)a.ww || fabsf(__fsub_rn(cI.y, cT.y)) > (float)a.wh) cI = cT;
  if (lane == 0) a.corners[idx] = cI;
}

ezc_status subpix_device(ezc_context* ctx, const ezc_images* imgs, float2* corners, const int32_t* image_index, int n,
                         int win_w, int win_h, int max_iter, double eps) {
  EZC_CHECK(win_w >= 1 && win_h >= 1 && win_w <= kMaxWin && win_h <= kMaxWin, EZC_ERR_INVALID,
            "corner_subpix: window half-size must be in [1,%d]", kMaxWin);
  if (n <= 0) return EZC_OK;
  SubpixArgs a{};
  a.imgs = imgs->d; a.pitch = imgs->pitch; a.image_stride = imgs->image_stride;
  a.W = imgs->width; a.H = imgs->height;
  a.corners = corners; a.image_index = image_index; a.n = n;
  a.ww = win_w; a.wh = win_h;
  a.max_iters = std::min(std::max(max_iter, 1), 100);
  const double e = std::max(eps, 0.0);
  a.eps2 = (float)(e * e);
  const int iw = 2 * win_w + 1, ih = 2 * win_h + 1;
  for (int i = 0; i < ih; ++i) {
    const float y = (float)(i - win_h) / win_h;
    const float vy = std::exp(-y * y);
    for (int j = 0; j < iw; ++j) {
      const float x = (float)(j - win_w) / win_w;
      a.mask[i * iw + j] = (float)(vy * std::exp(-x * x));
    }
  }
  const int grid = (n + kSubpixWarps - 1) / kSubpixWarps;
  corner_subpix_kernel<<<grid, kSubpixWarps * 32, 0, ctx->stream>>>(a);
  cudaError_t e2 = cudaGetLastError();
  if (e2 != cudaSuccess) { set_error("corner_subpix launch failed: %s", cudaGetErrorString(e2)); return EZC_ERR_CUDA; }
  ctx->kernel_launches++;
  return EZC_OK;
}

}  // namespace ezc

extern "C" {

static ezc_status images_upload_impl(ezc_context* ctx, const ezc_image_batch* b, ezc_images** out, bool async) {
  EZC_CHECK(ctx && b && out, EZC_ERR_INVALID, "ezc_images_upload: null argument");
  *out = nullptr;
  EZC_CHECK(b->n_images >= 0 && b->width > 0 && b->height > 0, EZC_ERR_INVALID, "ezc_images_upload: bad sizes");
  EZC_CHECK(b->n_images == 0 || b->data, EZC_ERR_INVALID, "ezc_images_upload: null data");
  EZC_CUDA(cudaSetDevice(ctx->device));
  const int64_t row = b->row_stride > 0 ? b->row_stride : b->width;
  const int64_t ist = b->image_stride > 0 ? b->image_stride : row * b->height;
  ezc_images* im = new ezc_images();
  im->ctx = ctx; im->n = b->n_images; im->width = b->width; im->height = b->height;
  im->pitch = b->width; im->image_stride = (int64_t)b->width * b->height;
  const size_t bytes = (size_t)std::max<int64_t>(1, im->image_stride * b->n_images);
  cudaError_t e = im->own.reserve(bytes + 64);
  cudaStream_t st = ctx->stream;
  if (async && e == cudaSuccess && b->n_images > 0) {
    if (!ctx->copy_stream) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    st = ctx->copy_stream;
    // ~64 MB per readiness event: fine enough to pipeline, coarse enough to keep the event count small
    im->chunk = (int)std::max<int64_t>(1, (int64_t(64) << 20) / std::max<int64_t>(1, im->image_stride));
  }
  const bool packed = (row == b->width && ist == row * b->height);
  for (int i = 0; e == cudaSuccess && i < b->n_images;) {
    const int nb = im->chunk > 0 ? std::min(im->chunk, b->n_images - i) : (packed ? b->n_images : 1);
    if (packed) {
      e = cudaMemcpyAsync(im->own.as<uint8_t>() + (size_t)i * im->image_stride, b->data + (size_t)i * ist, (size_t)nb * ist,
                          cudaMemcpyHostToDevice, st);
    } else {
      for (int j = i; e == cudaSuccess && j < i + nb; ++j)
        e = cudaMemcpy2DAsync(im->own.as<uint8_t>() + (size_t)j * im->image_stride, im->pitch, b->data + (size_t)j * ist, row,
                              b->width, b->height, cudaMemcpyHostToDevice, st);
    }
    if (im->chunk > 0 && e == cudaSuccess) {
      cudaEvent_t ev;
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e == cudaSuccess) { im->ready.push_back(ev); e = cudaEventRecord(ev, st); }
    }
    i += nb;
  }
  if (e == cudaSuccess && !async) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ezc::set_error("ezc_images_upload: %s", cudaGetErrorString(e)); delete im; return EZC_ERR_CUDA; }
  im->d = im->own.as<uint8_t>();
  *out = im;
  return EZC_OK;
}

ezc_status ezc_images_upload(ezc_context* ctx, const ezc_image_batch* b, ezc_images** out) {
  return images_upload_impl(ctx, b, out, false);
}
ezc_status ezc_images_upload_async(ezc_context* ctx, const ezc_image_batch* b, ezc_images** out) {
  return images_upload_impl(ctx, b, out, true);
}

ezc_status ezc_images_wrap_device(ezc_context* ctx, const ezc_image_batch* b, ezc_images** out) {
  EZC_CHECK(ctx && b && out && b->data, EZC_ERR_INVALID, "ezc_images_wrap_device: null argument");
  EZC_CHECK(b->n_images >= 0 && b->width > 0 && b->height > 0, EZC_ERR_INVALID, "ezc_images_wrap_device: bad sizes");
  ezc_images* im = new ezc_images();
  im->ctx = ctx; im->n = b->n_images; im->width = b->width; im->height = b->height;
  im->pitch = b->row_stride > 0 ? b->row_stride : b->width;
  im->image_stride = b->image_stride > 0 ? b->image_stride : im->pitch * b->height;
  im->d = b->data;
  *out = im;
  return EZC_OK;
}

void ezc_images_destroy(ezc_images* im) {
  if (!im) return;
  cudaSetDevice(im->ctx->device);
  cudaStreamSynchronize(im->ctx->stream);
  delete im;
}

ezc_status ezc_corner_subpix(ezc_context* ctx, const ezc_images* imgs, float* corners, const int32_t* image_index,
                             int32_t n, int32_t win_w, int32_t win_h, int32_t max_iter, double eps) {
  EZC_CHECK(ctx && imgs && (n == 0 || (corners && image_index)), EZC_ERR_INVALID, "ezc_corner_subpix: null argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  if (n <= 0) return EZC_OK;
  EZC_CUDA(ezc::images_wait(ctx, imgs, imgs->n - 1));
  for (int i = 0; i < n; ++i)
    EZC_CHECK(image_index[i] >= 0 && image_index[i] < imgs->n, EZC_ERR_INVALID, "ezc_corner_subpix: image index out of range");
  for (int i = 0; i < n; ++i)  // cv::cornerSubPix asserts Rect(0,0,cols,rows).contains(corner)
    EZC_CHECK(corners[2 * i] >= 0 && corners[2 * i] < imgs->width && corners[2 * i + 1] >= 0 && corners[2 * i + 1] < imgs->height,
              EZC_ERR_INVALID, "ezc_corner_subpix: corner %d (%g,%g) outside the image", i, corners[2 * i], corners[2 * i + 1]);
  ezc::DevBuf dc, di;
  ezc::SyncOnExit sync_on_exit(ctx->stream);
  EZC_CUDA(dc.reserve(sizeof(float2) * n));
  EZC_CUDA(di.reserve(sizeof(int32_t) * n));
  EZC_CUDA(cudaMemcpyAsync(dc.p, corners, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(di.p, image_index, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (ezc_status st = ezc::subpix_device(ctx, imgs, dc.as<float2>(), di.as<int32_t>(), n, win_w, win_h, max_iter, eps)) return st;
  EZC_CUDA(cudaMemcpyAsync(corners, dc.p, sizeof(float2) * n, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}

}  // extern "C"
