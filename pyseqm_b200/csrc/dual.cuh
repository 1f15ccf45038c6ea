// dual.cuh -- forward-mode differentiation scalar with three directional derivatives.
// The pair-integral device functions are templates over the scalar type: `double` for the SCF kernels,
// `Dual3` for the force kernel, where the three derivative slots are d/dR_i(x,y,z) of the pair.  This is
// how "analytic gradients reuse the pair kernels" (north star): the same code path is differentiated
// exactly instead of transcribing the reference's 900 lines of expanded derivative formulas
// (seqm/seqm_functions/anal_grad.py:718-1605).
#pragma once
#include "seqm_rt.h"

struct Dual3 {
  double v, d0, d1, d2;
  SEQM_HD Dual3() : v(0.0), d0(0.0), d1(0.0), d2(0.0) {}
  SEQM_HD Dual3(double a) : v(a), d0(0.0), d1(0.0), d2(0.0) {}
  SEQM_HD Dual3(double a, double x, double y, double z) : v(a), d0(x), d1(y), d2(z) {}
};

SEQM_HD Dual3 operator+(const Dual3& a, const Dual3& b) { return Dual3(a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2); }
SEQM_HD Dual3 operator-(const Dual3& a, const Dual3& b) { return Dual3(a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2); }
SEQM_HD Dual3 operator-(const Dual3& a) { return Dual3(-a.v, -a.d0, -a.d1, -a.d2); }
SEQM_HD Dual3 operator*(const Dual3& a, const Dual3& b) {
  return Dual3(a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2);
}
SEQM_HD Dual3 operator/(const Dual3& a, const Dual3& b) {
  double inv = 1.0 / b.v;
  double q = a.v * inv;
  return Dual3(q, (a.d0 - q * b.d0) * inv, (a.d1 - q * b.d1) * inv, (a.d2 - q * b.d2) * inv);
}
SEQM_HD Dual3 operator+(const Dual3& a, double b) { return Dual3(a.v + b, a.d0, a.d1, a.d2); }
SEQM_HD Dual3 operator+(double b, const Dual3& a) { return Dual3(a.v + b, a.d0, a.d1, a.d2); }
SEQM_HD Dual3 operator-(const Dual3& a, double b) { return Dual3(a.v - b, a.d0, a.d1, a.d2); }
SEQM_HD Dual3 operator-(double b, const Dual3& a) { return Dual3(b - a.v, -a.d0, -a.d1, -a.d2); }
SEQM_HD Dual3 operator*(const Dual3& a, double b) { return Dual3(a.v * b, a.d0 * b, a.d1 * b, a.d2 * b); }
SEQM_HD Dual3 operator*(double b, const Dual3& a) { return Dual3(a.v * b, a.d0 * b, a.d1 * b, a.d2 * b); }
SEQM_HD Dual3 operator/(const Dual3& a, double b) { double i = 1.0 / b; return Dual3(a.v * i, a.d0 * i, a.d1 * i, a.d2 * i); }
SEQM_HD Dual3 operator/(double a, const Dual3& b) {
  double inv = 1.0 / b.v;
  double q = a * inv;
  double f = -q * inv;
  return Dual3(q, f * b.d0, f * b.d1, f * b.d2);
}
SEQM_HD Dual3& operator+=(Dual3& a, const Dual3& b) { a.v += b.v; a.d0 += b.d0; a.d1 += b.d1; a.d2 += b.d2; return a; }
SEQM_HD Dual3& operator-=(Dual3& a, const Dual3& b) { a.v -= b.v; a.d0 -= b.d0; a.d1 -= b.d1; a.d2 -= b.d2; return a; }

// one-derivative dual: value + d/dr (used by the adjoint force kernel for the radial derivatives)
struct Dual1 {
  double v, d;
  SEQM_HD Dual1() : v(0.0), d(0.0) {}
  SEQM_HD Dual1(double a) : v(a), d(0.0) {}
  SEQM_HD Dual1(double a, double b) : v(a), d(b) {}
};
SEQM_HD Dual1 operator+(const Dual1& a, const Dual1& b) { return Dual1(a.v + b.v, a.d + b.d); }
SEQM_HD Dual1 operator-(const Dual1& a, const Dual1& b) { return Dual1(a.v - b.v, a.d - b.d); }
SEQM_HD Dual1 operator-(const Dual1& a) { return Dual1(-a.v, -a.d); }
SEQM_HD Dual1 operator*(const Dual1& a, const Dual1& b) { return Dual1(a.v * b.v, a.d * b.v + a.v * b.d); }
SEQM_HD Dual1 operator/(const Dual1& a, const Dual1& b) {
  double inv = 1.0 / b.v, q = a.v * inv;
  return Dual1(q, (a.d - q * b.d) * inv);
}
SEQM_HD Dual1 operator+(const Dual1& a, double b) { return Dual1(a.v + b, a.d); }
SEQM_HD Dual1 operator+(double b, const Dual1& a) { return Dual1(a.v + b, a.d); }
SEQM_HD Dual1 operator-(const Dual1& a, double b) { return Dual1(a.v - b, a.d); }
SEQM_HD Dual1 operator-(double b, const Dual1& a) { return Dual1(b - a.v, -a.d); }
SEQM_HD Dual1 operator*(const Dual1& a, double b) { return Dual1(a.v * b, a.d * b); }
SEQM_HD Dual1 operator*(double b, const Dual1& a) { return Dual1(a.v * b, a.d * b); }
SEQM_HD Dual1 operator/(const Dual1& a, double b) { double i = 1.0 / b; return Dual1(a.v * i, a.d * i); }
SEQM_HD Dual1 operator/(double a, const Dual1& b) {
  double inv = 1.0 / b.v, q = a * inv;
  return Dual1(q, -q * inv * b.d);
}
SEQM_HD Dual1& operator+=(Dual1& a, const Dual1& b) { a.v += b.v; a.d += b.d; return a; }
SEQM_HD Dual1 sq_root(const Dual1& x) { double s = sqrt(x.v); return Dual1(s, 0.5 / s * x.d); }
SEQM_HD Dual1 inv_sqrt(const Dual1& x) { double s = 1.0 / sqrt(x.v); return Dual1(s, -0.5 * s / x.v * x.d); }
SEQM_HD Dual1 e_xp(const Dual1& x) { double e = exp(x.v); return Dual1(e, e * x.d); }
SEQM_HD double val(const Dual1& x) { return x.v; }

// scalar helpers, overloaded for double and Dual3
SEQM_HD double sq_root(double x) { return sqrt(x); }
SEQM_HD Dual3 sq_root(const Dual3& x) {
  double s = sqrt(x.v);
  double f = 0.5 / s;
  return Dual3(s, f * x.d0, f * x.d1, f * x.d2);
}
SEQM_HD double inv_sqrt(double x) { return 1.0 / sqrt(x); }
SEQM_HD Dual3 inv_sqrt(const Dual3& x) {
  double s = 1.0 / sqrt(x.v);
  double f = -0.5 * s / x.v;
  return Dual3(s, f * x.d0, f * x.d1, f * x.d2);
}
SEQM_HD double e_xp(double x) { return exp(x); }
SEQM_HD Dual3 e_xp(const Dual3& x) {
  double e = exp(x.v);
  return Dual3(e, e * x.d0, e * x.d1, e * x.d2);
}
SEQM_HD double val(double x) { return x; }
SEQM_HD double val(const Dual3& x) { return x.v; }
// integer power by repeated multiplication (n small, >= 0)
template <class T>
SEQM_HD T ipow(T x, int n) {
  T r = T(1.0);
  for (int i = 0; i < n; ++i) r = r * x;
  return r;
}
