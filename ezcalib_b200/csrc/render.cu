// render.cu -- synthetic calibration-target renderer on the GPU (data generator for bench.py / large tests;
// NOT part of the reference's hot path).  The reference ships no images, and BASELINE's
// workloads need hundreds to thousands of multi-megapixel frames resident in HBM (SURVEY.md 8d / hard part 7),
// so they are rendered where they are consumed.  Same model as ezcalib_b200/synth.py::render_view:
// ss x ss area sampling of a planar target (z_w = 1) through the camera model, black 20 / white 235,
// Gaussian blur sigma 0.8, additive N(0, sigma_n) noise, u8.
#include "images.cuh"
#include "lm_models.cuh"
#include "tag36h11.cuh"

namespace ezc {

struct RenderArgs {
  int model, W, H, n_img, ss;
  int pattern;            // 0 chessboard, 1 aprilgrid
  int rows, cols;
  double size, space;     // square size | tag size, tag space ratio
  double black, white, noise_sigma;
  unsigned long long seed;
  const double* cams;     // [n_img][12] fx fy cx cy dist[8]
  const double* poses;    // [n_img][7]
};

__device__ void distort_any(int model, const double* k, double x, double y, double& xd, double& yd) {
  double a00, a01, a10, a11, ddx[8], ddy[8];
  switch (model) {
    case RAD1: distort<RAD1, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
    case RAD2: distort<RAD2, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
    case RAD3: distort<RAD3, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
    case RADTAN4: distort<RADTAN4, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
    case RADTAN5: distort<RADTAN5, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
    case RADTAN8: distort<RADTAN8, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
    default: distort<KB4, false>(k, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy); break;
  }
}

// plane coordinates (X,Y on z_w = 1) seen through every pixel CORNER (u = i - 0.5)
__global__ void render_lattice_kernel(const RenderArgs a, float2* __restrict__ lat) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long L = (long long)(a.W + 1) * (a.H + 1);
  if (gid >= L * a.n_img) return;
  const int im = (int)(gid / L);
  const int p = (int)(gid - (long long)im * L);
  const int iy = p / (a.W + 1), ix = p - iy * (a.W + 1);
  const double* cam = a.cams + (size_t)im * 12;
  const double* T = a.poses + (size_t)im * 7;
  const double xdn = ((double)ix - 0.5 - cam[2]) / cam[0], ydn = ((double)iy - 0.5 - cam[3]) / cam[1];
  double x = xdn, y = ydn;
  for (int it = 0; it < 10; ++it) {  // Newton with a central-difference Jacobian
    const double h = 1e-6;
    double f0x, f0y, ax, ay, bx, by, cx, cy, dx, dy;
    distort_any(a.model, cam + 4, x, y, f0x, f0y);
    distort_any(a.model, cam + 4, x + h, y, ax, ay);
    distort_any(a.model, cam + 4, x - h, y, bx, by);
    distort_any(a.model, cam + 4, x, y + h, cx, cy);
    distort_any(a.model, cam + 4, x, y - h, dx, dy);
    const double j00 = (ax - bx) / (2 * h), j10 = (ay - by) / (2 * h), j01 = (cx - dx) / (2 * h), j11 = (cy - dy) / (2 * h);
    const double ex = f0x - xdn, ey = f0y - ydn;
    const double det = j00 * j11 - j01 * j10;
    x -= (j11 * ex - j01 * ey) / det;
    y -= (-j10 * ex + j00 * ey) / det;
  }
  // rotation matrix columns from the quaternion
  const double qx = T[0], qy = T[1], qz = T[2], qw = T[3];
  const double R00 = 1 - 2 * (qy * qy + qz * qz), R01 = 2 * (qx * qy - qz * qw), R02 = 2 * (qx * qz + qy * qw);
  const double R10 = 2 * (qx * qy + qz * qw), R11 = 1 - 2 * (qx * qx + qz * qz), R12 = 2 * (qy * qz - qx * qw);
  const double R20 = 2 * (qx * qz - qy * qw), R21 = 2 * (qy * qz + qx * qw), R22 = 1 - 2 * (qx * qx + qy * qy);
  // lambda*(x,y,1) = R p_w + t, p_w.z = 1  ->  p_w = R^T (lambda ray - t)
  const double num = 1.0 + (R02 * T[4] + R12 * T[5] + R22 * T[6]);
  const double den = R02 * x + R12 * y + R22;
  const double lam = num / den;
  const double px = lam * x - T[4], py = lam * y - T[5], pz = lam - T[6];
  float2 o;
  if (lam > 0 && isfinite(lam)) { o.x = (float)(R00 * px + R10 * py + R20 * pz); o.y = (float)(R01 * px + R11 * py + R21 * pz); }
  else { o.x = 1e9f; o.y = 1e9f; }
  lat[gid] = o;
}

__device__ __forceinline__ float pattern_value(const RenderArgs& a, float X, float Y) {
  if (a.pattern == 0) {
    const float sq = (float)a.size;
    const int j = (int)floorf(-X / sq), i = (int)floorf(Y / sq);
    if (i < 0 || i >= a.rows || j < 0 || j >= a.cols) return 1.f;
    return (((i + (a.cols - 1 - j)) & 1) == 0) ? 0.f : 1.f;
  }
  const float s = (float)a.size, off = (float)(a.size * (1.0 + a.space)), cs = s / 10.f;
  const int c = (int)floorf(X / off), r = (int)floorf(-Y / off);
  if (c < 0 || c >= a.cols || r < 0 || r >= a.rows) return 1.f;
  const float lx = X - c * off, ly = Y + r * off;
  if (!(lx >= 0 && lx < s && ly > -s && ly <= 0)) return 1.f;
  const int cx = min(9, max(0, (int)floorf(lx / cs))), ry = min(9, max(0, (int)floorf((ly + s) / cs)));
  if (cx < 2 || cx > 7 || ry < 2 || ry > 7) return 0.f;
  const int idx = (ry - 2) * 6 + (cx - 2);
  return (float)((c_tag36h11[r * a.cols + c] >> (35 - idx)) & 1ULL);
}

__global__ void render_sample_kernel(const RenderArgs a, const float2* __restrict__ lat, float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long N = (long long)a.W * a.H;
  if (gid >= N * a.n_img) return;
  const int im = (int)(gid / N);
  const int p = (int)(gid - (long long)im * N);
  const int y = p / a.W, x = p - y * a.W;
  const float2* L = lat + (long long)im * (a.W + 1) * (a.H + 1);
  const float2 c00 = L[(long long)y * (a.W + 1) + x], c01 = L[(long long)y * (a.W + 1) + x + 1];
  const float2 c10 = L[(long long)(y + 1) * (a.W + 1) + x], c11 = L[(long long)(y + 1) * (a.W + 1) + x + 1];
  float acc = 0.f;
  for (int sb = 0; sb < a.ss; ++sb) {
    const float wb = (sb + 0.5f) / a.ss;
    for (int sa = 0; sa < a.ss; ++sa) {
      const float wa = (sa + 0.5f) / a.ss;
      const float X = (1 - wb) * ((1 - wa) * c00.x + wa * c01.x) + wb * ((1 - wa) * c10.x + wa * c11.x);
      const float Y = (1 - wb) * ((1 - wa) * c00.y + wa * c01.y) + wb * ((1 - wa) * c10.y + wa * c11.y);
      acc += pattern_value(a, X, Y);
    }
  }
  out[gid] = (float)a.black + (float)(a.white - a.black) * acc / (float)(a.ss * a.ss);
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__global__ void render_finish_kernel(const RenderArgs a, const float* __restrict__ in, uint8_t* __restrict__ out, long long pitch,
                                     long long istride) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long N = (long long)a.W * a.H;
  if (gid >= N * a.n_img) return;
  const int im = (int)(gid / N);
  const int p = (int)(gid - (long long)im * N);
  const int y = p / a.W, x = p - y * a.W;
  const float* I = in + (long long)im * N;
  // 7x7 Gaussian, sigma 0.8, replicated border
  const float k[7] = {0.00044f, 0.02191f, 0.22854f, 0.49822f, 0.22854f, 0.02191f, 0.00044f};
  float v = 0.f;
  for (int dy = -3; dy <= 3; ++dy) {
    const int yy = min(a.H - 1, max(0, y + dy));
    float row = 0.f;
    for (int dx = -3; dx <= 3; ++dx) row += k[dx + 3] * I[(long long)yy * a.W + min(a.W - 1, max(0, x + dx))];
    v += k[dy + 3] * row;
  }
  if (a.noise_sigma > 0) {
    const unsigned long long h = splitmix64(a.seed ^ splitmix64((unsigned long long)gid));
    const float u1 = ((h >> 40) + 1.0f) * (1.0f / 16777217.0f);
    const float u2 = ((h >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
    v += (float)a.noise_sigma * sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
  }
  out[(long long)im * istride + (long long)y * pitch + x] = (uint8_t)min(255.f, max(0.f, rintf(v)));
}

}  // namespace ezc

extern "C" ezc_status ezc_render_targets(ezc_context* ctx, int32_t pattern, int32_t rows, int32_t cols, double size, double space,
                                         int32_t model, int32_t width, int32_t height, int32_t n_images, const double* cams12,
                                         const double* poses7, double noise_sigma, uint64_t seed, ezc_images** out) {
  using namespace ezc;
  EZC_CHECK(ctx && cams12 && poses7 && out && n_images > 0 && width > 0 && height > 0, EZC_ERR_INVALID, "ezc_render_targets: bad argument");
  EZC_CHECK(pattern == 0 || (rows * cols <= kNumTag36h11), EZC_ERR_INVALID, "ezc_render_targets: too many tags");
  EZC_CUDA(cudaSetDevice(ctx->device));
  *out = nullptr;
  ezc_images* im = new ezc_images();
  im->ctx = ctx; im->n = n_images; im->width = width; im->height = height; im->pitch = width; im->image_stride = (int64_t)width * height;
  if (im->own.reserve((size_t)im->image_stride * n_images + 64) != cudaSuccess) { delete im; set_error("ezc_render_targets: out of device memory"); return EZC_ERR_CUDA; }
  im->d = im->own.as<uint8_t>();
  DevBuf d_cams, d_poses, lat, tmp;
  const int chunk = std::max<int>(1, (int)std::min<long long>(n_images, (1LL << 28) / ((long long)width * height)));
  EZC_CUDA(d_cams.reserve(sizeof(double) * 12 * n_images));
  EZC_CUDA(d_poses.reserve(sizeof(double) * 7 * n_images));
  EZC_CUDA(cudaMemcpyAsync(d_cams.p, cams12, sizeof(double) * 12 * n_images, cudaMemcpyHostToDevice, ctx->stream));
  EZC_CUDA(cudaMemcpyAsync(d_poses.p, poses7, sizeof(double) * 7 * n_images, cudaMemcpyHostToDevice, ctx->stream));
  EZC_CUDA(lat.reserve(sizeof(float2) * (size_t)(width + 1) * (height + 1) * chunk));
  EZC_CUDA(tmp.reserve(sizeof(float) * (size_t)width * height * chunk));
  for (int i0 = 0; i0 < n_images; i0 += chunk) {
    RenderArgs a{};
    a.model = model; a.W = width; a.H = height; a.n_img = std::min(chunk, n_images - i0); a.ss = 4;
    a.pattern = pattern; a.rows = rows; a.cols = cols; a.size = size; a.space = space;
    a.black = 20; a.white = 235; a.noise_sigma = noise_sigma; a.seed = seed + 0x1000003ULL * (unsigned long long)i0;
    a.cams = d_cams.as<double>() + (size_t)i0 * 12; a.poses = d_poses.as<double>() + (size_t)i0 * 7;
    const long long L = (long long)(width + 1) * (height + 1) * a.n_img, N = (long long)width * height * a.n_img;
    render_lattice_kernel<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(a, lat.as<float2>());
    render_sample_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(a, lat.as<float2>(), tmp.as<float>());
    render_finish_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(a, tmp.as<float>(), im->own.as<uint8_t>() + (size_t)i0 * im->image_stride,
                                                                                  im->pitch, im->image_stride);
    EZC_CUDA(cudaGetLastError());
    ctx->kernel_launches += 3;
  }
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = im;
  return EZC_OK;
}

extern "C" ezc_status ezc_images_download(ezc_context* ctx, const ezc_images* imgs, int32_t first, int32_t count, uint8_t* host_out) {
  EZC_CHECK(ctx && imgs && host_out && first >= 0 && count >= 0 && first + count <= imgs->n, EZC_ERR_INVALID, "ezc_images_download: bad argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  EZC_CUDA(ezc::images_wait(ctx, imgs, first + count - 1));
  for (int i = 0; i < count; ++i)
    EZC_CUDA(cudaMemcpy2DAsync(host_out + (size_t)i * imgs->width * imgs->height, imgs->width, imgs->d + (size_t)(first + i) * imgs->image_stride,
                               imgs->pitch, imgs->width, imgs->height, cudaMemcpyDeviceToHost, ctx->stream));
  EZC_CUDA(cudaStreamSynchronize(ctx->stream));
  return EZC_OK;
}
