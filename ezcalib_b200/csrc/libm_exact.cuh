// libm_exact.cuh -- bit-exact device restatements of the host libm functions the reference's AprilTag code
// relies on for DISCRETE decisions (edge costs, merge predicates): glibc 2.39's atan2f (fdlibm e_atan2f.c /
// s_atanf.c) and sinf/cosf (ARM optimized routines, s_sincosf_data.c).  CUDA's own atan2f/sinf/cosf differ in
// the last ulp, which flips integer edge costs (Edge.cc:22-23).  Verified bit-identical to the host glibc on
// 4e7 random arguments each when written; tests/test_libm_exact.py repeats the check (8e5 arguments per function) through the C ABI on the GPU box.
// Compile the including translation unit with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ezc {

__device__ __forceinline__ float fd_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  const int32_t hx = __float_as_int(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {
    if (ix > 0x7f800000) return x + x;
    return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = __int_as_float(ix);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else { id = 3; x = -1.0f / x; }
    }
  }
  float z = x * x;
  const float w = z * z;
  const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

__device__ __forceinline__ float fd_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = __float_as_int(x), ix = hx & 0x7fffffff;
  const int32_t hy = __float_as_int(y), iy = hy & 0x7fffffff;
  if ((ix > 0x7f800000) || (iy > 0x7f800000)) return x + y;
  if (hx == 0x3f800000) return fd_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = fd_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

__device__ __forceinline__ float gl_sincos_poly(double x, double x2, int n, bool flip) {
  const double sgn = flip ? -1.0 : 1.0;
  const double c0 = sgn * 0x1p0, c1 = sgn * -0x1.ffffffd0c621cp-2, c2 = sgn * 0x1.55553e1068f19p-5,
               c3 = sgn * -0x1.6c087e89a359dp-10, c4 = sgn * 0x1.99343027bf8c3p-16;
  const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double s1 = s2c + x2 * s3c;
    const double x7 = x3 * x2;
    const double s = x + x3 * s1c;
    return (float)(s + x7 * s1);
  }
  const double x4 = x2 * x2;
  const double cc2 = c3 + x2 * c4;
  const double cc1 = c1 + x2 * c2;
  const double x6 = x4 * x2;
  const double c = c0 + x2 * cc1;
  return (float)(c + x6 * cc2);
}

// valid for |y| < 120 (the reference only passes angles in [-2pi, 2pi])
__device__ __forceinline__ float gl_sinf_or_cosf(float y, bool is_cos) {
  double x = (double)y;
  const float ay = fabsf(y);
  if (ay < 0x1.921fb6p-1f) {
    const double x2 = x * x;
    if (ay < 0x1p-12f) return is_cos ? 1.0f : y;
    return gl_sincos_poly(x, x2, is_cos ? 1 : 0, false);
  }
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const double r = x * hpi_inv;
  const int n = ((int32_t)r + 0x800000) >> 24;
  x = x - (double)n * hpi;
  const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return gl_sincos_poly(x * s, x * x, is_cos ? (n ^ 1) : n, (n & 2) != 0);
}
__device__ __forceinline__ float gl_sinf(float y) { return gl_sinf_or_cosf(y, false); }
__device__ __forceinline__ float gl_cosf(float y) { return gl_sinf_or_cosf(y, true); }

}  // namespace ezc
