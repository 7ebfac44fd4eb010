// lm_models.cuh -- per-observation reprojection residual and ANALYTIC Jacobian for the seven
// distortion models of the reference (one device function, templated on the model).
//
// Reference definitions (the residual functors that Ceres differentiates with Jets):
//   rad1     /root/reference/include/dist_models/dist_model_k1.hpp:84-189
//   rad2     /root/reference/include/dist_models/dist_model_k1k2.hpp:86-193
//   rad3     /root/reference/include/dist_models/dist_model_k1k2k3.hpp:88-199
//   radtan4  /root/reference/include/dist_models/dist_model_k1k2p1p2.hpp:101-220
//   radtan5  /root/reference/include/dist_models/dist_model_k1k2k3p1p2.hpp:103-226   [k1,k2,k3,p1,p2]
//   radtan8  /root/reference/include/dist_models/dist_model_radtan_full.hpp:113-242  [k1..k6,p1,p2]
//   kb4      /root/reference/include/dist_models/dist_model_kb.hpp:105-232           [k1..k4]
// Pose block: left perturbation T <- exp(delta) T, delta = (upsilon, omega)
//   /root/reference/include/ceres_params/se3_param.hpp:12-41  =>  d p_c / d delta = [ I | -[p_c]x ].
#pragma once
#include <cuda_runtime.h>

namespace ezc {

enum : int { RAD1 = 0, RAD2 = 1, RAD3 = 2, RADTAN4 = 3, RADTAN5 = 4, RADTAN8 = 5, KB4 = 6, NUM_MODELS = 7 };

__host__ __device__ constexpr int num_dist(int model) {
  return model == RAD1 ? 1 : model == RAD2 ? 2 : model == RAD3 ? 3 : model == RADTAN4 ? 4 : model == RADTAN5 ? 5
       : model == RADTAN8 ? 8 : 4;
}

// Shared (intrinsic) parameters of one camera, packed [focal(2) pp(2) dist(8)].
// For mono focal, focal[1] is ignored.
struct CamParams {
  double focal[2];
  double pp[2];
  double dist[8];
};

// p_c = R(q) p + t with the Sophus point-rotation formula (unit quaternion [x y z w]).
__device__ __forceinline__ void se3_act(const double* __restrict__ T, double px, double py, double pz,
                                        double& X, double& Y, double& Z) {
  const double qx = T[0], qy = T[1], qz = T[2], qw = T[3];
  double uvx = qy * pz - qz * py;
  double uvy = qz * px - qx * pz;
  double uvz = qx * py - qy * px;
  uvx += uvx; uvy += uvy; uvz += uvz;
  X = px + qw * uvx + (qy * uvz - qz * uvy) + T[4];
  Y = py + qw * uvy + (qz * uvx - qx * uvz) + T[5];
  Z = pz + qw * uvz + (qx * uvy - qy * uvx) + T[6];
}

// Distortion value + partials.  dd[i][0..1] = d(xd,yd)/d dist_i.
template <int MODEL, bool WITH_JAC>
__device__ __forceinline__ void distort(const double* __restrict__ k, double x, double y, double& xd, double& yd,
                                        double& a00, double& a01, double& a10, double& a11,
                                        double (&ddx)[8], double (&ddy)[8]) {
  if constexpr (MODEL == KB4) {
    const double r2 = x * x + y * y;
    const double r = sqrt(r2);
    const double th = atan(r);
    const double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2;
    const double poly = 1.0 + t2 * (k[0] + t2 * k[1] + t4 * k[2] + t6 * k[3]);
    const double P = th * poly;
    const double ir = 1.0 / r;          // r == 0 -> inf/NaN exactly like the reference (no guard)
    const double g = P * ir;
    xd = g * x; yd = g * y;
    if constexpr (WITH_JAC) {
      const double t8 = t4 * t4;
      const double dP = 1.0 + 3.0 * k[0] * t2 + 5.0 * k[1] * t4 + 7.0 * k[2] * t6 + 9.0 * k[3] * t8;
      // dg/dr / r
      const double gp_r = (dP * r / (1.0 + r2) - P) * ir * ir * ir;
      a00 = g + x * x * gp_r; a01 = x * y * gp_r; a10 = a01; a11 = g + y * y * gp_r;
      const double t3 = t2 * th;
      const double b0 = t3 * ir, b1 = b0 * t2, b2 = b1 * t2, b3 = b2 * t2;
      ddx[0] = x * b0; ddy[0] = y * b0;
      ddx[1] = x * b1; ddy[1] = y * b1;
      ddx[2] = x * b2; ddy[2] = y * b2;
      ddx[3] = x * b3; ddy[3] = y * b3;
    }
  } else {
    constexpr bool TAN = (MODEL == RADTAN4 || MODEL == RADTAN5 || MODEL == RADTAN8);
    constexpr int NK = (MODEL == RAD1) ? 1 : (MODEL == RAD2 || MODEL == RADTAN4) ? 2 : 3;  // numerator order
    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, r4 = r2 * r2;
    const double k1 = k[0];
    const double k2 = NK >= 2 ? k[1] : 0.0;
    const double k3 = NK >= 3 ? k[2] : 0.0;
    double D, Dp;  // D and dD/d(r2)
    double dk[6];  // dD/dk_i
    if constexpr (MODEL == RADTAN8) {
      const double N = 1.0 + r2 * (k1 + k2 * r2 + k3 * r4);
      const double M = 1.0 + r2 * (k[3] + k[4] * r2 + k[5] * r4);
      const double iM = 1.0 / M;
      D = N * iM;
      if constexpr (WITH_JAC) {
        const double Np = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;
        const double Mp = k[3] + 2.0 * k[4] * r2 + 3.0 * k[5] * r4;
        Dp = (Np - D * Mp) * iM;
        dk[0] = r2 * iM; dk[1] = r4 * iM; dk[2] = r4 * r2 * iM;
        const double m = -D * iM;
        dk[3] = m * r2; dk[4] = m * r4; dk[5] = m * r4 * r2;
      }
    } else {
      D = 1.0 + r2 * (k1 + k2 * r2 + k3 * r4);
      if constexpr (WITH_JAC) {
        Dp = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;
        dk[0] = r2; dk[1] = r4; dk[2] = r4 * r2;
      }
    }
    constexpr int NKTOT = (MODEL == RADTAN8) ? 6 : NK;
    double p1 = 0.0, p2 = 0.0;
    if constexpr (TAN) { p1 = k[NKTOT]; p2 = k[NKTOT + 1]; }
    const double xy = x * y;
    xd = x * D; yd = y * D;
    if constexpr (TAN) {
      xd += 2.0 * p1 * xy + p2 * (r2 + 2.0 * x2);
      yd += 2.0 * p2 * xy + p1 * (r2 + 2.0 * y2);
    }
    if constexpr (WITH_JAC) {
      a00 = D + 2.0 * x2 * Dp; a01 = 2.0 * xy * Dp; a11 = D + 2.0 * y2 * Dp;
      if constexpr (TAN) {
        a00 += 2.0 * p1 * y + 6.0 * p2 * x;
        a01 += 2.0 * p1 * x + 2.0 * p2 * y;
        a11 += 2.0 * p2 * x + 6.0 * p1 * y;
      }
      a10 = a01;
#pragma unroll
      for (int i = 0; i < NKTOT; ++i) { ddx[i] = x * dk[i]; ddy[i] = y * dk[i]; }
      if constexpr (TAN) {
        ddx[NKTOT] = 2.0 * xy;           ddy[NKTOT] = r2 + 2.0 * y2;      // d/dp1
        ddx[NKTOT + 1] = r2 + 2.0 * x2;  ddy[NKTOT + 1] = 2.0 * xy;       // d/dp2
      }
    }
  }
}

// Residual only (candidate-cost evaluation, per-frame RMSE).
template <int MODEL, bool MONO>
__device__ __forceinline__ void residual(const CamParams& c, const double* __restrict__ pose, double px, double py,
                                         double pz, double u, double v, double& rx, double& ry) {
  double X, Y, Z;
  se3_act(pose, px, py, pz, X, Y, Z);
  const double iz = 1.0 / Z;
  const double x = X * iz, y = Y * iz;
  double xd, yd, a00, a01, a10, a11;
  double ddx[8], ddy[8];
  distort<MODEL, false>(c.dist, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy);
  const double fx = c.focal[0], fy = MONO ? c.focal[0] : c.focal[1];
  rx = fx * xd + c.pp[0] - u;
  ry = fy * yd + c.pp[1] - v;
}

// Residual + Jacobian.  Jp[2][6]: pose tangent (upsilon, omega).  Jc[2][NF+2+ND]: [focal | pp | dist].
// `wgt` scales every Jacobian entry (the Cauchy corrector sqrt(rho') is applied by the caller through it when WEIGHTED:
// the residual is evaluated first, the weight derived from it, and the 2x2 / focal factors are scaled once instead of
// scaling all 2 x (6 + k) entries).
template <int MODEL, bool MONO, bool WEIGHTED>
__device__ __forceinline__ void residual_jacobian_w(const CamParams& c, const double* __restrict__ pose, double px,
                                                    double py, double pz, double u, double v, double& rx, double& ry,
                                                    double& wgt, double& half_rho,
                                                    double (&Jpx)[6], double (&Jpy)[6], double (&Jcx)[12], double (&Jcy)[12]) {
  constexpr int NF = MONO ? 1 : 2;
  constexpr int ND = num_dist(MODEL);
  double X, Y, Z;
  se3_act(pose, px, py, pz, X, Y, Z);
  const double iz = 1.0 / Z;
  const double x = X * iz, y = Y * iz;
  double xd, yd, a00, a01, a10, a11;
  double ddx[8], ddy[8];
  distort<MODEL, true>(c.dist, x, y, xd, yd, a00, a01, a10, a11, ddx, ddy);
  const double fx = c.focal[0], fy = MONO ? c.focal[0] : c.focal[1];
  rx = fx * xd + c.pp[0] - u;
  ry = fy * yd + c.pp[1] - v;
  double w = 1.0;
  if constexpr (WEIGHTED) {
    // ceres::CauchyLoss(1): rho = log(1+s), rho' = 1/(1+s), rho'' < 0 -> Corrector alpha = 0 -> scale by sqrt(rho')
    // The cost term 0.5 log(1+s) is NOT taken here: a double-precision log is ~80 instructions per observation on the
    // pipe this kernel is bound by.  The caller multiplies the (1+s) of its observations (exponent peeled off with frexp)
    // and takes one log per lane at the end: sum log = log prod.  `half_rho` carries 1+s.
    const double ss = rx * rx + ry * ry;
    w = rsqrt(1.0 + ss);
    half_rho = 1.0 + ss;
  }
  wgt = w;
  const double wfx = w * fx, wfy = w * fy;
  const double m00 = wfx * a00, m01 = wfx * a01, m10 = wfy * a10, m11 = wfy * a11;
  // d(x,y)/d delta = [[iz, 0, -x iz, -xy, 1+x^2, -y], [0, iz, -y iz, -(1+y^2), xy, x]]
  const double xy = x * y;
  Jpx[0] = m00 * iz;               Jpy[0] = m10 * iz;
  Jpx[1] = m01 * iz;               Jpy[1] = m11 * iz;
  Jpx[2] = -(Jpx[0] * x + Jpx[1] * y);  Jpy[2] = -(Jpy[0] * x + Jpy[1] * y);
  Jpx[3] = -(m00 * xy + m01 * (1.0 + y * y));  Jpy[3] = -(m10 * xy + m11 * (1.0 + y * y));
  Jpx[4] = m00 * (1.0 + x * x) + m01 * xy;     Jpy[4] = m10 * (1.0 + x * x) + m11 * xy;
  Jpx[5] = m01 * x - m00 * y;      Jpy[5] = m11 * x - m10 * y;
  if constexpr (MONO) {
    Jcx[0] = w * xd; Jcy[0] = w * yd;
  } else {
    Jcx[0] = w * xd; Jcy[0] = 0.0;
    Jcx[1] = 0.0; Jcy[1] = w * yd;
  }
  Jcx[NF] = w; Jcy[NF] = 0.0;
  Jcx[NF + 1] = 0.0; Jcy[NF + 1] = w;
#pragma unroll
  for (int i = 0; i < ND; ++i) { Jcx[NF + 2 + i] = wfx * ddx[i]; Jcy[NF + 2 + i] = wfy * ddy[i]; }
}

// unweighted form (tests, debug entry)
template <int MODEL, bool MONO>
__device__ __forceinline__ void residual_jacobian(const CamParams& c, const double* __restrict__ pose, double px,
                                                  double py, double pz, double u, double v, double& rx, double& ry,
                                                  double (&Jpx)[6], double (&Jpy)[6], double (&Jcx)[12], double (&Jcy)[12]) {
  double w, hr;
  residual_jacobian_w<MODEL, MONO, false>(c, pose, px, py, pz, u, v, rx, ry, w, hr, Jpx, Jpy, Jcx, Jcy);
}

// ---- SE3 exp(delta) * T  (Sophus SE3::exp + group product with normalisation) -------------
__device__ __forceinline__ void se3_plus(const double* __restrict__ x, const double* __restrict__ d, double* __restrict__ out) {
  const double ux = d[0], uy = d[1], uz = d[2], wx = d[3], wy = d[4], wz = d[5];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double imag, real, c1, c2;
  if (th2 < 1e-20) {  // Sophus::Constants<double>::epsilon()^2
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
    // V = R(exp(omega)) in Sophus for theta < eps; to first order in omega that is I + [omega]x,
    // exact to < 1e-20 here.
    c1 = 1.0; c2 = 0.0;
  } else {
    const double th = sqrt(th2);
    const double half = 0.5 * th;
    double sh, ch;
    sincos(half, &sh, &ch);
    imag = sh / th; real = ch;
    double st, ct;
    sincos(th, &st, &ct);
    c1 = (1.0 - ct) / th2;
    c2 = (th - st) / (th2 * th);
  }
  const double qdx = imag * wx, qdy = imag * wy, qdz = imag * wz, qdw = real;
  // td = V * upsilon, V = I + c1 [w]x + c2 [w]x^2
  const double cx_ = wy * uz - wz * uy, cy_ = wz * ux - wx * uz, cz_ = wx * uy - wy * ux;          // w x u
  const double ccx = wy * cz_ - wz * cy_, ccy = wz * cx_ - wx * cz_, ccz = wx * cy_ - wy * cx_;    // w x (w x u)
  const double tdx = ux + c1 * cx_ + c2 * ccx, tdy = uy + c1 * cy_ + c2 * ccy, tdz = uz + c1 * cz_ + c2 * ccz;
  const double ax = x[0], ay = x[1], az = x[2], aw = x[3];
  double qw = qdw * aw - qdx * ax - qdy * ay - qdz * az;
  double qx = qdw * ax + qdx * aw + qdy * az - qdz * ay;
  double qy = qdw * ay + qdy * aw + qdz * ax - qdx * az;
  double qz = qdw * az + qdz * aw + qdx * ay - qdy * ax;
  const double inv = 1.0 / sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
  out[0] = qx * inv; out[1] = qy * inv; out[2] = qz * inv; out[3] = qw * inv;
  // t' = td + R(qd) t
  const double tx = x[4], ty = x[5], tz = x[6];
  double uvx = qdy * tz - qdz * ty, uvy = qdz * tx - qdx * tz, uvz = qdx * ty - qdy * tx;
  uvx += uvx; uvy += uvy; uvz += uvz;
  out[4] = tdx + tx + qdw * uvx + (qdy * uvz - qdz * uvy);
  out[5] = tdy + ty + qdw * uvy + (qdz * uvx - qdx * uvz);
  out[6] = tdz + tz + qdw * uvz + (qdx * uvy - qdy * uvx);
}

}  // namespace ezc
