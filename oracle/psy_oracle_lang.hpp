// psy_oracle_lang.hpp — host-side stand-in for the Armadillo subset psydiff's user equations are written in.
// TEST INFRASTRUCTURE ONLY (part of the CPU oracle; see oracle/psy_oracle.cpp).
//
// The reference compiles `dx = DRIFT*x;`-style statements against arma::mat / arma::colvec with every referenced
// parameter pulled out of an Rcpp::List (R/prepareModel.R:737-770).  The oracle compiles the SAME statements
// against these small dynamic-size types, with parameters read from the person's flat parameter list.
#pragma once
#include <cmath>

#ifndef PSY_REAL
#define PSY_REAL double
#endif
typedef PSY_REAL real;

namespace ol {

struct Mat;
constexpr int VCAP = 16, MCAP = 256;  // inline storage: no heap traffic in the hot loop

struct Vec {
  real v[VCAP];
  int len = 0;
  Vec() {}
  explicit Vec(int n) : len(n) { for (int i = 0; i < n; i++) v[i] = 0.0; }
  Vec(const real* p, int n) : len(n) { for (int i = 0; i < n; i++) v[i] = p[i]; }
  real& operator()(int i) { return v[i]; }
  const real& operator()(int i) const { return v[i]; }
  real& operator[](int i) { return v[i]; }
  const real& operator[](int i) const { return v[i]; }
  int n() const { return len; }
  Vec& operator=(real s) { len = 1; v[0] = s; return *this; }  // arma: scalar assignment makes a 1x1 object
};

struct Mat {
  int r = 0, c = 0;
  real a[MCAP];  // column-major
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_) { for (int i = 0; i < r_ * c_; i++) a[i] = 0.0; }
  Mat(const real* p, int r_, int c_) : r(r_), c(c_) { for (int i = 0; i < r_ * c_; i++) a[i] = p[i]; }
  int size() const { return r * c; }
  real& operator()(int i, int j) { return a[i + (size_t)j * r]; }
  const real& operator()(int i, int j) const { return a[i + (size_t)j * r]; }
  real& operator()(int i) { return a[i]; }
  const real& operator()(int i) const { return a[i]; }
  real& operator[](int i) { return a[i]; }
  const real& operator[](int i) const { return a[i]; }
};

inline Vec operator+(const Vec& a, const Vec& b) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] + b.v[i]; return r; }
inline Vec operator-(const Vec& a, const Vec& b) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] - b.v[i]; return r; }
inline Vec operator%(const Vec& a, const Vec& b) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] * b.v[i]; return r; }
inline Vec operator/(const Vec& a, const Vec& b) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] / b.v[i]; return r; }
inline Vec operator+(const Mat& a, const Vec& b) { Vec r(b.n()); for (int i = 0; i < b.n(); i++) r.v[i] = a.a[i] + b.v[i]; return r; }
inline Vec operator+(const Vec& a, const Mat& b) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] + b.a[i]; return r; }
inline Vec operator-(const Mat& a, const Vec& b) { Vec r(b.n()); for (int i = 0; i < b.n(); i++) r.v[i] = a.a[i] - b.v[i]; return r; }
inline Vec operator-(const Vec& a, const Mat& b) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] - b.a[i]; return r; }
inline Vec operator*(real s, const Vec& a) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = s * a.v[i]; return r; }
inline Vec operator*(const Vec& a, real s) { return s * a; }
inline Vec operator/(const Vec& a, real s) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] / s; return r; }
inline Vec operator+(const Vec& a, real s) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] + s; return r; }
inline Vec operator+(real s, const Vec& a) { return a + s; }
inline Vec operator-(const Vec& a, real s) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = a.v[i] - s; return r; }
inline Vec operator-(const Vec& a) { Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = -a.v[i]; return r; }
inline Mat operator+(const Mat& a, const Mat& b) { Mat r(a.r, a.c); for (int i = 0; i < a.size(); i++) r.a[i] = a.a[i] + b.a[i]; return r; }
inline Mat operator-(const Mat& a, const Mat& b) { Mat r(a.r, a.c); for (int i = 0; i < a.size(); i++) r.a[i] = a.a[i] - b.a[i]; return r; }
inline Mat operator*(real s, const Mat& a) { Mat r(a.r, a.c); for (int i = 0; i < a.size(); i++) r.a[i] = s * a.a[i]; return r; }
inline Mat operator*(const Mat& a, real s) { return s * a; }
inline Mat operator-(const Mat& a) { return -1.0 * a; }
inline Vec operator*(const Mat& A, const Vec& x) {
  Vec r(A.r);
  for (int i = 0; i < A.r; i++) { real acc = 0; for (int k = 0; k < A.c; k++) acc += A(i, k) * x.v[k]; r.v[i] = acc; }
  return r;
}
inline Mat operator*(const Mat& A, const Mat& B) {
  Mat r(A.r, B.c);
  for (int j = 0; j < B.c; j++) for (int i = 0; i < A.r; i++) { real acc = 0; for (int k = 0; k < A.c; k++) acc += A(i, k) * B(k, j); r(i, j) = acc; }
  return r;
}
}  // namespace ol

namespace arma {
inline ol::Vec exp(const ol::Vec& a) { ol::Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = std::exp(a.v[i]); return r; }
inline ol::Vec log(const ol::Vec& a) { ol::Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = std::log(a.v[i]); return r; }
#define OL_ELEMENTWISE(NAME, EXPR) \
  inline ol::Vec NAME(const ol::Vec& a) { ol::Vec r(a.n()); for (int i = 0; i < a.n(); i++) { const real v = a.v[i]; r.v[i] = (EXPR); } return r; }
OL_ELEMENTWISE(sin, std::sin(v))
OL_ELEMENTWISE(cos, std::cos(v))
OL_ELEMENTWISE(tanh, std::tanh(v))
OL_ELEMENTWISE(sqrt, std::sqrt(v))
OL_ELEMENTWISE(abs, std::fabs(v))
OL_ELEMENTWISE(square, v * v)
#undef OL_ELEMENTWISE
inline ol::Vec pow(const ol::Vec& a, real p) { ol::Vec r(a.n()); for (int i = 0; i < a.n(); i++) r.v[i] = std::pow(a.v[i], p); return r; }
inline real norm(const ol::Vec& a) { real s = 0; for (int i = 0; i < a.n(); i++) s += a.v[i] * a.v[i]; return std::sqrt(s); }
inline real max(const ol::Vec& a) { real m = a.v[0]; for (int i = 1; i < a.n(); i++) m = a.v[i] > m ? a.v[i] : m; return m; }
inline real min(const ol::Vec& a) { real m = a.v[0]; for (int i = 1; i < a.n(); i++) m = a.v[i] < m ? a.v[i] : m; return m; }
inline real mean(const ol::Vec& a) { real s = 0; for (int i = 0; i < a.n(); i++) s += a.v[i]; return s / (real)a.n(); }
inline real as_scalar(const ol::Vec& a) { return a.v[0]; }
inline real as_scalar(const ol::Mat& a) { return a.a[0]; }
inline real dot(const ol::Vec& a, const ol::Vec& b) { real s = 0; for (int i = 0; i < a.n(); i++) s += a.v[i] * b.v[i]; return s; }
inline real accu(const ol::Vec& a) { real s = 0; for (int i = 0; i < a.n(); i++) s += a.v[i]; return s; }
inline real sum(const ol::Vec& a) { return accu(a); }
typedef ol::Mat mat;       // initialised temporaries the user declares: `arma::mat T = arma::expm(A);`
typedef ol::Vec colvec;
typedef ol::Vec vec;
// arma::expm: scaling and squaring around a degree-12 Taylor polynomial (third-party arithmetic, restated; Armadillo uses Pade)
inline ol::Mat expm(const ol::Mat& A) {
  const int n = A.r;
  real nrm = 0;
  for (int j = 0; j < n; j++) { real cs = 0; for (int i = 0; i < n; i++) cs += std::fabs(A(i, j)); nrm = cs > nrm ? cs : nrm; }
  int s = 0;
  real sc = 1.0;
  while (nrm * sc > 0.5 && s < 64) { sc *= 0.5; s++; }
  const ol::Mat B = sc * A;
  ol::Mat E(n, n);
  for (int i = 0; i < n; i++) E(i, i) = 1.0;
  for (int k = 12; k >= 1; k--) { E = ((real)1.0 / (real)k) * (B * E); for (int i = 0; i < n; i++) E(i, i) += 1.0; }
  for (int q = 0; q < s; q++) E = E * E;
  return E;
}
inline ol::Mat trans(const ol::Mat& a) { ol::Mat r(a.c, a.r); for (int j = 0; j < a.c; j++) for (int i = 0; i < a.r; i++) r(j, i) = a(i, j); return r; }
}  // namespace arma
