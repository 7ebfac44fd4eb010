// oracle/lm_ref/ref_functors.cc -- TEST INFRASTRUCTURE ONLY.
//
// Thin C driver around the reference's OWN, UNMODIFIED headers, included from where they lie:
//   /root/reference/include/dist_models/dist_models.hpp   (7 DistParam classes, 14 autodiff functors,
//                                                          createCeresCostFunction, distortCamPoint)
//   /root/reference/include/ceres_params/se3_param.hpp    (AutoDiffLocalLeftSE3)
// compiled against the stand-in headers in shim/ (ceres::Jet / AutoDiffCostFunction / AutoDiffManifold,
// Sophus::SE3, Eigen::Matrix / Map) because Ceres, Sophus and Eigen are not in this image.  What this pins:
// the residual DEFINITIONS and the manifold update exactly as the reference spells them.  What it does not
// pin: Ceres' trust-region rules (oracle/lm_oracle.cpp restates those).
#include <cstring>
#include <memory>
#include <vector>
#include <iostream>
#include <cassert>
#include <limits>

#include "dist_models/dist_models.hpp"

namespace {

// Camera::setupInitialDistortion, /root/reference/src/camera.cpp:67-94 (same order as EZC_RAD1..EZC_KB4)
std::unique_ptr<DistParam> make_dist(int model, bool mono) {
  switch (model) {
    case 0: return std::unique_ptr<DistParam>(new Rad1DistParam(mono));
    case 1: return std::unique_ptr<DistParam>(new Rad2DistParam(mono));
    case 2: return std::unique_ptr<DistParam>(new Rad3DistParam(mono));
    case 3: return std::unique_ptr<DistParam>(new RadTan4DistParam(mono));
    case 4: return std::unique_ptr<DistParam>(new RadTan5DistParam(mono));
    case 5: return std::unique_ptr<DistParam>(new RadTan8DistParam(mono));
    case 6: return std::unique_ptr<DistParam>(new KB4DistParam(mono));
  }
  return nullptr;
}

}  // namespace

extern "C" {

int lmref_num_dist(int model) {
  auto d = make_dist(model, true);
  return d ? d->getNumberOfParameters() : -1;
}

// One residual block through DistParam::createCeresCostFunction(u, v, wpt)->Evaluate.
// r[2]; J_ambient[2][nf+2+nd+7] (columns focal | pp | dist | pose7, what Ceres' autodiff returns);
// J_local[2][nf+2+nd+6] (pose columns multiplied by AutoDiffLocalLeftSE3::PlusJacobian, what the solver sees).
int lmref_eval(int model, int mono, const double* focal, const double* pp, const double* dist, const double* pose7,
               const double* wpt3, double u, double v, double* r, double* J_ambient, double* J_local) {
  auto d = make_dist(model, mono != 0);
  if (!d) return -1;
  const int nf = mono ? 1 : 2, nd = d->getNumberOfParameters();
  std::unique_ptr<ceres::CostFunction> cf(d->createCeresCostFunction(u, v, Eigen::Vector3d(wpt3[0], wpt3[1], wpt3[2])));
  const std::vector<int>& sz = cf->parameter_block_sizes();
  if (sz.size() != 4 || sz[0] != nf || sz[1] != 2 || sz[2] != nd || sz[3] != 7 || cf->num_residuals() != 2) return -2;
  const double* params[4] = {focal, pp, dist, pose7};
  double jf[2 * 2], jp[2 * 2], jd[2 * 8], jT[2 * 7];
  double* jac[4] = {jf, jp, jd, jT};
  if (!cf->Evaluate(params, r, jac)) return -3;
  // value-only path (jacobians == nullptr evaluates in plain doubles; Jet division rounds differently by <= 1 ulp)
  double r2[2];
  if (!cf->Evaluate(params, r2, nullptr)) return -4;
  for (int i = 0; i < 2; ++i)
    if (r2[i] == r2[i] && r[i] == r[i] && std::fabs(r2[i] - r[i]) > 1e-12 * (1.0 + std::fabs(r[i]))) return -4;
  AutoDiffLocalLeftSE3 manifold;
  double PJ[7 * 6];
  if (!manifold.PlusJacobian(pose7, PJ)) return -5;
  const int ka = nf + 2 + nd + 7, kl = nf + 2 + nd + 6;
  for (int row = 0; row < 2; ++row) {
    int c = 0;
    for (int i = 0; i < nf; ++i, ++c) J_ambient[row * ka + c] = J_local[row * kl + c] = jf[row * nf + i];
    for (int i = 0; i < 2; ++i, ++c) J_ambient[row * ka + c] = J_local[row * kl + c] = jp[row * 2 + i];
    for (int i = 0; i < nd; ++i, ++c) J_ambient[row * ka + c] = J_local[row * kl + c] = jd[row * nd + i];
    for (int i = 0; i < 7; ++i) J_ambient[row * ka + c + i] = jT[row * 7 + i];
    for (int j = 0; j < 6; ++j) {
      double s = 0.0;
      for (int a = 0; a < 7; ++a) s += jT[row * 7 + a] * PJ[a * 6 + j];
      J_local[row * kl + c + j] = s;
    }
  }
  return 0;
}

// AutoDiffLocalLeftSE3::Plus / PlusJacobian / Minus (se3_param.hpp:12-41)
int lmref_plus(const double* x7, const double* delta6, double* out7) {
  AutoDiffLocalLeftSE3 m;
  return m.Plus(x7, delta6, out7) ? 0 : -1;
}
int lmref_plus_jacobian(const double* x7, double* J42) {
  AutoDiffLocalLeftSE3 m;
  return m.PlusJacobian(x7, J42) ? 0 : -1;
}
int lmref_minus(const double* y7, const double* x7, double* out6) {
  AutoDiffLocalLeftSE3 m;
  return m.Minus(y7, x7, out6) ? 0 : -1;
}

// DistParam::resetParameters + getDistParameters + distortCamPoint(x, y) (the plain-double path used by
// IntrinsicsParam::projCamToImage for the per-frame RMSE, /root/reference/src/ezcalibrator.cpp:245-253)
int lmref_distort(int model, const double* dist, double x, double y, double* out2) {
  auto d = make_dist(model, true);
  if (!d) return -1;
  std::vector<double> v(dist, dist + d->getNumberOfParameters());
  d->resetParameters(v);
  const std::vector<double> back = d->getDistParameters();
  for (size_t i = 0; i < v.size(); ++i) if (back[i] != v[i]) return -2;
  const Eigen::Vector2d o = d->distortCamPoint(x, y);
  const Eigen::Vector2d o3 = d->distortCamPoint(Eigen::Vector3d(2.0 * x, 2.0 * y, 2.0));
  out2[0] = o[0]; out2[1] = o[1];
  (void)o3;
  return 0;
}

// DistParam::undistortCamPoint (dist_model_base.hpp:33-71) -- off-path in the reference, exported for the host mirror test
int lmref_undistort(int model, const double* dist, double xd, double yd, double* out2) {
  auto d = make_dist(model, true);
  if (!d) return -1;
  d->resetParameters(std::vector<double>(dist, dist + d->getNumberOfParameters()));
  const Eigen::Vector2d o = d->undistortCamPoint(xd, yd);
  out2[0] = o[0]; out2[1] = o[1];
  return 0;
}

// The std-dev quirk of DistParam::setDistParamsStd (dist_model_base.hpp:90-102): returns the vector length
int lmref_dist_std_quirk(int model, const double* cov, double* out, int cap) {
  auto d = make_dist(model, true);
  if (!d) return -1;
  const int nd = d->getNumberOfParameters();
  d->setDistParamsStd(std::vector<double>(cov, cov + nd * nd));
  const std::vector<double> s = d->getDistParamsStd();
  for (int i = 0; i < (int)s.size() && i < cap; ++i) out[i] = s[i];
  return (int)s.size();
}

}  // extern "C"
