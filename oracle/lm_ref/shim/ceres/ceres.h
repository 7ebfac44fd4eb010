// Stand-in for <ceres/ceres.h> -- TEST INFRASTRUCTURE ONLY (oracle/lm_ref).
// Ceres 2.2.0 is pinned by /root/reference/build.sh:45 but is neither in the reference tree nor in
// this image.  This header provides the pieces the reference's functor headers need to compile
// UNMODIFIED: ceres::Jet (forward-mode duals with the derivative rules of ceres/jet.h),
// ceres::CostFunction, ceres::AutoDiffCostFunction<Functor, kNumResiduals, Ns...>::Evaluate
// (row-major jacobians per parameter block) and ceres::AutoDiffManifold<Functor, Ambient, Tangent>
// (Plus, PlusJacobian at delta = 0, Minus).
#pragma once
#include <cmath>
#include <utility>
#include <vector>

#define CERES_VERSION_MAJOR 2
#define CERES_VERSION_MINOR 2
#define CERES_VERSION_REVISION 0

namespace ceres {

template <class T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = T(0); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(0); }  // NOLINT
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(0); v[k] = T(1); }
};

#define EZC_JET template <class T, int N> inline
EZC_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
EZC_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
EZC_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
EZC_JET Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
EZC_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
// jet.h: h = f / g ; dh = (df - h dg) / g
EZC_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; const T gi = T(1.0) / g.a; h.a = f.a * gi; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi; return h; }
EZC_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a += s; return h; }
EZC_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a += s; return h; }
EZC_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a -= s; return h; }
EZC_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h = -f; h.a += s; return h; }
EZC_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
EZC_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
EZC_JET Jet<T, N> operator/(const Jet<T, N>& f, T s) { return f * (T(1.0) / s); }
EZC_JET Jet<T, N> operator/(T s, const Jet<T, N>& g) { Jet<T, N> h; const T gi = T(1.0) / g.a; h.a = s * gi; const T m = -s * gi * gi; for (int i = 0; i < N; ++i) h.v[i] = m * g.v[i]; return h; }
EZC_JET Jet<T, N>& operator+=(Jet<T, N>& f, const Jet<T, N>& g) { f = f + g; return f; }
EZC_JET Jet<T, N>& operator-=(Jet<T, N>& f, const Jet<T, N>& g) { f = f - g; return f; }
EZC_JET Jet<T, N>& operator*=(Jet<T, N>& f, const Jet<T, N>& g) { f = f * g; return f; }
EZC_JET Jet<T, N>& operator/=(Jet<T, N>& f, const Jet<T, N>& g) { f = f / g; return f; }
EZC_JET bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
EZC_JET bool operator<(const Jet<T, N>& f, T g) { return f.a < g; }
EZC_JET bool operator>(const Jet<T, N>& f, T g) { return f.a > g; }
EZC_JET Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sqrt(f.a); const T m = T(1.0) / (T(2.0) * h.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
EZC_JET Jet<T, N> atan(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::atan(f.a); const T m = T(1.0) / (T(1.0) + f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
EZC_JET Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T m = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
EZC_JET Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T m = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
EZC_JET Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
// jet.h: atan2(g, f): d = (f dg - g df) / (f^2 + g^2)
EZC_JET Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) { Jet<T, N> h; h.a = std::atan2(g.a, f.a); const T t = T(1.0) / (f.a * f.a + g.a * g.a); for (int i = 0; i < N; ++i) h.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]); return h; }
#undef EZC_JET
using std::abs;
using std::atan;
using std::atan2;
using std::cos;
using std::sin;
using std::sqrt;

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int> sizes_;
  int num_residuals_ = 0;
};

template <class Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  static constexpr int kBlocks = sizeof...(Ns);
  static constexpr int kTotal = (Ns + ...);
 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) { sizes_ = {Ns...}; num_residuals_ = kNumResiduals; }
  ~AutoDiffCostFunction() override { delete f_; }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (jacobians == nullptr) return call(parameters, residuals, std::make_index_sequence<kBlocks>());
    typedef Jet<double, kTotal> J;
    const int sizes[kBlocks] = {Ns...};
    J x[kTotal];
    const J* ptr[kBlocks];
    int k = 0;
    for (int b = 0; b < kBlocks; ++b) {
      ptr[b] = x + k;
      for (int i = 0; i < sizes[b]; ++i, ++k) x[k] = J(parameters[b][i], k);
    }
    J out[kNumResiduals];
    if (!call(ptr, out, std::make_index_sequence<kBlocks>())) return false;
    for (int r = 0; r < kNumResiduals; ++r) residuals[r] = out[r].a;
    k = 0;
    for (int b = 0; b < kBlocks; ++b) {
      if (jacobians[b] != nullptr)
        for (int r = 0; r < kNumResiduals; ++r)
          for (int i = 0; i < sizes[b]; ++i) jacobians[b][r * sizes[b] + i] = out[r].v[k + i];
      k += sizes[b];
    }
    return true;
  }
 private:
  template <class T, size_t... I>
  bool call(T const* const* p, T* out, std::index_sequence<I...>) const { return (*f_)(p[I]..., out); }
  Functor* f_;
};

template <class Functor, int kAmbient, int kTangent>
class AutoDiffManifold {
 public:
  AutoDiffManifold() {}
  int AmbientSize() const { return kAmbient; }
  int TangentSize() const { return kTangent; }
  bool Plus(const double* x, const double* delta, double* x_plus_delta) const { return f_.Plus(x, delta, x_plus_delta); }
  bool Minus(const double* y, const double* x, double* y_minus_x) const { return f_.Minus(y, x, y_minus_x); }
  // d Plus(x, delta) / d delta at delta = 0, row-major kAmbient x kTangent
  bool PlusJacobian(const double* x, double* jacobian) const {
    typedef Jet<double, kTangent> J;
    J jx[kAmbient], jd[kTangent], out[kAmbient];
    for (int i = 0; i < kAmbient; ++i) jx[i] = J(x[i]);
    for (int i = 0; i < kTangent; ++i) jd[i] = J(0.0, i);
    if (!f_.Plus(jx, jd, out)) return false;
    for (int r = 0; r < kAmbient; ++r) for (int c = 0; c < kTangent; ++c) jacobian[r * kTangent + c] = out[r].v[c];
    return true;
  }
 private:
  Functor f_;
};

}  // namespace ceres
