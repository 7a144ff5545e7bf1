// oracle/refbuild/include/RcppArmadillo.h — TEST INFRASTRUCTURE (oracle/_ref build): stand-in for <RcppArmadillo.h>.
//
// Purpose: compile the REFERENCE'S OWN C++ — inst/include/psydiffBasic.h, inst/include/psydiffInternals.h and the
// fitModel / getGradients / getGradient translation unit that R/modelTemplate.R holds as a string — unmodified, where it
// lies under /root/reference, on a machine that has neither R, Rcpp, Armadillo nor Boost.  Everything in this
// header is a restatement of THIRD-PARTY interfaces, written from their published semantics, covering exactly the surface
// the reference touches:
//   Rcpp  >= 1.0.5 : List / DataFrame / NumericVector / NumericMatrix / LogicalVector / StringVector with R's reference
//                    semantics (handles onto shared objects), name proxies, clone, as<>, sugar (==, unique, any, is_false),
//                    stop / warning.                       Rcpp sugar `unique()` is hash ordered; here: first appearance.
//   Armadillo      : dense double matrices, column-major; mat / colvec / rowvec / uvec; col / row / rows / cols / submat /
//                    elem / diag / span views; each_col; trans, inv, inv(trimatl), trimatl, chol("lower"), qr_econ, det,
//                    eye, diagmat, log, exp, sum, find, find_finite, find_nonfinite, is_finite, min; and, for user equations,
//                    % and element-wise /, dot, accu, mean, norm, max, as_scalar, sin / cos / tanh / sqrt / abs / square / pow,
//                    expm (scaling and squaring around a degree-12 Taylor polynomial).
//                    LAPACK-backed routines (inv, chol, qr_econ, det) are textbook algorithms here: unblocked Householder
//                    QR, Cholesky, LU with partial pivoting, forward substitution for triangular inverses.
// Not a general implementation of either library.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define R_NaN (std::numeric_limits<double>::quiet_NaN())

// =====================================================================================================================
namespace arma {

typedef unsigned long long uword;
struct span { uword a, b; span(uword a_, uword b_) : a(a_), b(b_) {} };
namespace fill { struct fill_zeros {}; static const fill_zeros zeros = fill_zeros(); }

struct uvec {
  std::vector<uword> v;
  uword n_elem = 0;
  uvec() {}
  explicit uvec(std::vector<uword> x) : v(std::move(x)) { n_elem = v.size(); }
  uword size() const { return v.size(); }
  uword operator()(uword i) const { return v.at(i); }
  uword operator[](uword i) const { return v.at(i); }
};
inline uword min(const uvec& u) {
  if (u.v.empty()) throw std::logic_error("min(): object has no elements");
  return *std::min_element(u.v.begin(), u.v.end());
}

class Mat;
struct View;
struct EachCol;
template <class T> struct is_dense : std::integral_constant<bool, std::is_base_of<Mat, T>::value || std::is_same<T, View>::value> {};

class Mat {
 public:
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<double> mem;
  Mat() {}
  Mat(uword r, uword c) { init(r, c); }
  Mat(uword r, uword c, const fill::fill_zeros&) { init(r, c); }
  void init(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem.assign(n_elem, 0.0); }
  double& operator()(uword i, uword j) { return mem.at(i + j * n_rows); }
  const double& operator()(uword i, uword j) const { return mem.at(i + j * n_rows); }
  double& operator()(uword i) { return mem.at(i); }
  const double& operator()(uword i) const { return mem.at(i); }
  double& operator[](uword i) { return mem.at(i); }
  const double& operator[](uword i) const { return mem.at(i); }
  double* begin() { return mem.data(); }
  double* end() { return mem.data() + n_elem; }
  const double* begin() const { return mem.data(); }
  const double* end() const { return mem.data() + n_elem; }
  uword size() const { return n_elem; }
  void fill(double v) { std::fill(mem.begin(), mem.end(), v); }
  void reshape(uword r, uword c) {  // column-major element order kept, missing elements zero
    std::vector<double> old = mem;
    init(r, c);
    for (uword k = 0; k < std::min<uword>(n_elem, old.size()); k++) mem[k] = old[k];
  }
  Mat& operator=(const View& v);
  Mat& operator+=(const Mat& o) { same(o); for (uword k = 0; k < n_elem; k++) mem[k] += o.mem[k]; return *this; }
  Mat& operator-=(const Mat& o) { same(o); for (uword k = 0; k < n_elem; k++) mem[k] -= o.mem[k]; return *this; }
  void same(const Mat& o) const { if (o.n_rows != n_rows || o.n_cols != n_cols) throw std::logic_error("arma stand-in: incompatible matrix dimensions"); }
  // views
  View col(uword j) const; View row(uword i) const; View rows(const uvec& r) const; View cols(const span& s) const;
  View submat(uword r0, uword c0, uword r1, uword c1) const; View submat(const uvec& r, const uvec& c) const;
  View elem(const uvec& i) const; View diag() const;
  View operator()(const uvec& i) const; View operator()(const span& s) const; View operator()(const span& s, uword c) const;
  EachCol each_col();
};

struct View {  // a set of elements of a parent matrix, shaped r x c (column-major order of idx)
  Mat* p; std::vector<uword> idx; uword r, c;
  operator Mat() const { Mat m(r, c); for (uword k = 0; k < idx.size(); k++) m.mem[k] = p->mem.at(idx[k]); return m; }
  void check(const Mat& m) const { if (m.n_elem != idx.size()) throw std::logic_error("arma stand-in: view assignment of wrong size"); }
  View& operator=(const Mat& m) { check(m); for (uword k = 0; k < idx.size(); k++) p->mem.at(idx[k]) = m.mem[k]; return *this; }
  View& operator=(const View& o) { return *this = Mat(o); }
  View& operator+=(const Mat& m) { check(m); for (uword k = 0; k < idx.size(); k++) p->mem.at(idx[k]) += m.mem[k]; return *this; }
  View& operator-=(const Mat& m) { check(m); for (uword k = 0; k < idx.size(); k++) p->mem.at(idx[k]) -= m.mem[k]; return *this; }
  View& operator+=(double s) { for (uword k : idx) p->mem.at(k) += s; return *this; }
  View& operator-=(double s) { for (uword k : idx) p->mem.at(k) -= s; return *this; }
};
inline Mat& Mat::operator=(const View& v) { Mat t = v; *this = t; return *this; }

struct EachCol {
  Mat* p;
  void operator+=(const Mat& v) { if (v.n_elem != p->n_rows) throw std::logic_error("each_col(): wrong size"); for (uword j = 0; j < p->n_cols; j++) for (uword i = 0; i < p->n_rows; i++) (*p)(i, j) += v.mem[i]; }
  void operator-=(const Mat& v) { if (v.n_elem != p->n_rows) throw std::logic_error("each_col(): wrong size"); for (uword j = 0; j < p->n_cols; j++) for (uword i = 0; i < p->n_rows; i++) (*p)(i, j) -= v.mem[i]; }
};
inline EachCol Mat::each_col() { return EachCol{this}; }

inline View make_view(const Mat* m, std::vector<uword> idx, uword r, uword c) {
  for (uword k : idx) if (k >= m->n_elem) throw std::out_of_range("arma stand-in: index out of bounds");
  return View{const_cast<Mat*>(m), std::move(idx), r, c};
}
inline View Mat::col(uword j) const { if (j >= n_cols) throw std::out_of_range("col(): index out of bounds"); std::vector<uword> x; for (uword i = 0; i < n_rows; i++) x.push_back(i + j * n_rows); return make_view(this, x, n_rows, 1); }
inline View Mat::row(uword i) const { if (i >= n_rows) throw std::out_of_range("row(): index out of bounds"); std::vector<uword> x; for (uword j = 0; j < n_cols; j++) x.push_back(i + j * n_rows); return make_view(this, x, 1, n_cols); }
inline View Mat::rows(const uvec& r) const { std::vector<uword> x; for (uword j = 0; j < n_cols; j++) for (uword i : r.v) { if (i >= n_rows) throw std::out_of_range("rows(): index out of bounds"); x.push_back(i + j * n_rows); } return make_view(this, x, r.size(), n_cols); }
inline View Mat::cols(const span& s) const { if (s.b >= n_cols || s.a > s.b) throw std::out_of_range("cols(): index out of bounds"); std::vector<uword> x; for (uword j = s.a; j <= s.b; j++) for (uword i = 0; i < n_rows; i++) x.push_back(i + j * n_rows); return make_view(this, x, n_rows, s.b - s.a + 1); }
inline View Mat::submat(uword r0, uword c0, uword r1, uword c1) const { if (r1 >= n_rows || c1 >= n_cols || r0 > r1 || c0 > c1) throw std::out_of_range("submat(): indices out of bounds"); std::vector<uword> x; for (uword j = c0; j <= c1; j++) for (uword i = r0; i <= r1; i++) x.push_back(i + j * n_rows); return make_view(this, x, r1 - r0 + 1, c1 - c0 + 1); }
inline View Mat::submat(const uvec& r, const uvec& c) const { std::vector<uword> x; for (uword j : c.v) for (uword i : r.v) { if (i >= n_rows || j >= n_cols) throw std::out_of_range("submat(): indices out of bounds"); x.push_back(i + j * n_rows); } return make_view(this, x, r.size(), c.size()); }
inline View Mat::elem(const uvec& i) const { return make_view(this, i.v, i.size(), 1); }
inline View Mat::diag() const { std::vector<uword> x; for (uword i = 0; i < std::min(n_rows, n_cols); i++) x.push_back(i + i * n_rows); return make_view(this, x, x.size(), 1); }
inline View Mat::operator()(const uvec& i) const { return elem(i); }
inline View Mat::operator()(const span& s) const { if (s.b >= n_elem || s.a > s.b) throw std::out_of_range("span: index out of bounds"); std::vector<uword> x; for (uword k = s.a; k <= s.b; k++) x.push_back(k); return make_view(this, x, x.size(), 1); }
inline View Mat::operator()(const span& s, uword cc) const { if (s.b >= n_rows || s.a > s.b || cc >= n_cols) throw std::out_of_range("span: index out of bounds"); std::vector<uword> x; for (uword i = s.a; i <= s.b; i++) x.push_back(i + cc * n_rows); return make_view(this, x, x.size(), 1); }

class Col : public Mat {
 public:
  Col() {}
  explicit Col(uword n) : Mat(n, 1) {}
  Col(uword n, const fill::fill_zeros&) : Mat(n, 1) {}
  Col(uword r, uword c, const fill::fill_zeros&) : Mat(r, c) { if (c != 1) throw std::logic_error("Col(): not a column"); }
  template <class T, class = typename std::enable_if<is_dense<T>::value>::type>
  Col(const T& o) { assign(Mat(o)); }
  template <class T, class = typename std::enable_if<is_dense<T>::value>::type>
  Col& operator=(const T& o) { assign(Mat(o)); return *this; }
  void assign(const Mat& m) { if (m.n_cols != 1 && m.n_elem != 0) throw std::logic_error("arma stand-in: assigning a non-column to a colvec"); n_rows = m.n_rows; n_cols = 1; n_elem = m.n_elem; mem = m.mem; }
  using Mat::operator();
};
class Row : public Mat {
 public:
  Row() {}
  explicit Row(uword n) : Mat(1, n) {}
  template <class T, class = typename std::enable_if<is_dense<T>::value>::type>
  Row(const T& o) { assign(Mat(o)); }
  template <class T, class = typename std::enable_if<is_dense<T>::value>::type>
  Row& operator=(const T& o) { assign(Mat(o)); return *this; }
  void assign(const Mat& m) { if (m.n_rows != 1 && m.n_elem != 0) throw std::logic_error("arma stand-in: assigning a non-row to a rowvec"); n_rows = 1; n_cols = m.n_cols; n_elem = m.n_elem; mem = m.mem; }
};
typedef Mat mat;
typedef Col colvec;
typedef Col vec;
typedef Row rowvec;

// ---- arithmetic (eager) ----------------------------------------------------------------------------------------------
inline Mat operator+(const Mat& a, const Mat& b) { a.same(b); Mat r = a; r += b; return r; }
inline Mat operator-(const Mat& a, const Mat& b) { a.same(b); Mat r = a; r -= b; return r; }
inline Mat operator*(const Mat& a, const Mat& b) {
  if (a.n_cols != b.n_rows) throw std::logic_error("arma stand-in: matrix multiplication: incompatible matrix dimensions");
  Mat r(a.n_rows, b.n_cols);
  for (uword j = 0; j < b.n_cols; j++)
    for (uword i = 0; i < a.n_rows; i++) {
      double acc = 0.0;
      for (uword k = 0; k < a.n_cols; k++) acc += a(i, k) * b(k, j);
      r(i, j) = acc;
    }
  return r;
}
inline Mat operator*(double s, const Mat& a) { Mat r = a; for (auto& x : r.mem) x *= s; return r; }
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator/(const Mat& a, double s) { Mat r = a; for (auto& x : r.mem) x /= s; return r; }
inline Mat operator+(const Mat& a, double s) { Mat r = a; for (auto& x : r.mem) x += s; return r; }
inline Mat operator-(const Mat& a) { return -1.0 * a; }
inline Mat operator%(const Mat& a, const Mat& b) { a.same(b); Mat r = a; for (uword k = 0; k < r.n_elem; k++) r.mem[k] *= b.mem[k]; return r; }
inline Mat operator/(const Mat& a, const Mat& b) { a.same(b); Mat r = a; for (uword k = 0; k < r.n_elem; k++) r.mem[k] /= b.mem[k]; return r; }
inline Mat operator-(const Mat& a, double s) { return a + (-s); }
inline Mat operator+(double s, const Mat& a) { return a + s; }
inline uvec operator==(const Mat& a, double s) { uvec u; for (double x : a.mem) u.v.push_back(x == s ? 1 : 0); u.n_elem = u.v.size(); return u; }

inline Mat trans(const Mat& a) { Mat r(a.n_cols, a.n_rows); for (uword j = 0; j < a.n_cols; j++) for (uword i = 0; i < a.n_rows; i++) r(j, i) = a(i, j); return r; }
inline Mat eye(uword r, uword c) { Mat m(r, c); for (uword i = 0; i < std::min(r, c); i++) m(i, i) = 1.0; return m; }
inline Mat expm(const Mat& A) {  // scaling and squaring, degree-12 Taylor (Armadillo: Pade; both accurate to rounding)
  if (A.n_rows != A.n_cols) throw std::logic_error("expm(): given matrix must be square sized");
  const uword n = A.n_rows; double nrm = 0.0;
  for (uword j = 0; j < n; j++) { double cs = 0.0; for (uword i = 0; i < n; i++) cs += std::fabs(A(i, j)); nrm = std::max(nrm, cs); }
  int s = 0; double sc = 1.0;
  while (nrm * sc > 0.5 && s < 64) { sc *= 0.5; s++; }
  Mat B = sc * A, E = eye(n, n);
  for (int k = 12; k >= 1; k--) { E = (1.0 / k) * (B * E); for (uword i = 0; i < n; i++) E(i, i) += 1.0; }
  for (int q = 0; q < s; q++) E = E * E;
  return E;
}
inline Mat diagmat(const Mat& v) { Mat m(v.n_elem, v.n_elem); for (uword i = 0; i < v.n_elem; i++) m(i, i) = v.mem[i]; return m; }
inline Mat log(const Mat& a) { Mat r = a; for (auto& x : r.mem) x = std::log(x); return r; }
inline Mat exp(const Mat& a) { Mat r = a; for (auto& x : r.mem) x = std::exp(x); return r; }
#define PSY_REF_ELEMENTWISE(NAME, EXPR) inline Mat NAME(const Mat& a) { Mat r = a; for (auto& v : r.mem) v = (EXPR); return r; }
PSY_REF_ELEMENTWISE(sin, std::sin(v))
PSY_REF_ELEMENTWISE(cos, std::cos(v))
PSY_REF_ELEMENTWISE(tanh, std::tanh(v))
PSY_REF_ELEMENTWISE(sqrt, std::sqrt(v))
PSY_REF_ELEMENTWISE(abs, std::fabs(v))
PSY_REF_ELEMENTWISE(square, v * v)
#undef PSY_REF_ELEMENTWISE
inline Mat pow(const Mat& a, double p) { Mat r = a; for (auto& v : r.mem) v = std::pow(v, p); return r; }
inline double norm(const Mat& a) { double s = 0.0; for (double x : a.mem) s += x * x; return std::sqrt(s); }   // vectors (2-norm)
inline double max(const Mat& a) { if (a.mem.empty()) throw std::logic_error("max(): object has no elements"); return *std::max_element(a.mem.begin(), a.mem.end()); }
inline double min(const Mat& a) { if (a.mem.empty()) throw std::logic_error("min(): object has no elements"); return *std::min_element(a.mem.begin(), a.mem.end()); }
inline double mean(const Mat& a) { double s = 0.0; for (double x : a.mem) s += x; return s / (double)a.n_elem; }   // vectors
inline double as_scalar(const Mat& a) { if (a.n_elem != 1) throw std::logic_error("as_scalar(): expression must evaluate to exactly one element"); return a.mem[0]; }
inline double accu(const Mat& a) { double s = 0.0; for (double x : a.mem) s += x; return s; }
inline double dot(const Mat& a, const Mat& b) { if (a.n_elem != b.n_elem) throw std::logic_error("dot(): objects must have the same number of elements"); double s = 0.0; for (uword k = 0; k < a.n_elem; k++) s += a.mem[k] * b.mem[k]; return s; }
inline double sum(const Mat& a) { double s = 0.0; for (double x : a.mem) s += x; return s; }  // vectors only (all the reference uses)
inline bool is_finite(const Mat& a) { for (double x : a.mem) if (!std::isfinite(x)) return false; return true; }
inline bool is_finite(double x) { return std::isfinite(x); }
inline uvec find(const uvec& mask) { uvec u; for (uword k = 0; k < mask.v.size(); k++) if (mask.v[k]) u.v.push_back(k); u.n_elem = u.v.size(); return u; }
inline uvec find_finite(const Mat& a) { uvec u; for (uword k = 0; k < a.n_elem; k++) if (std::isfinite(a.mem[k])) u.v.push_back(k); u.n_elem = u.v.size(); return u; }
inline uvec find_nonfinite(const Mat& a) { uvec u; for (uword k = 0; k < a.n_elem; k++) if (!std::isfinite(a.mem[k])) u.v.push_back(k); u.n_elem = u.v.size(); return u; }

struct TriL { Mat m; operator Mat() const { return m; } };   // trimatl(X): lower triangle, and the tag inv() dispatches on
inline TriL trimatl(const Mat& a) { if (a.n_rows != a.n_cols) throw std::logic_error("trimatl(): given matrix must be square sized"); Mat r = a; for (uword j = 0; j < a.n_cols; j++) for (uword i = 0; i < j; i++) r(i, j) = 0.0; return TriL{r}; }
inline Mat inv(const TriL& t) {  // inverse of a lower-triangular matrix by forward substitution (LAPACK dtrtri)
  const Mat& L = t.m; const uword n = L.n_rows; Mat X(n, n);
  for (uword j = 0; j < n; j++) {
    if (L(j, j) == 0.0) throw std::runtime_error("inv(): matrix is singular");
    X(j, j) = 1.0 / L(j, j);
    for (uword i = j + 1; i < n; i++) { double acc = 0.0; for (uword k = j; k < i; k++) acc += L(i, k) * X(k, j); X(i, j) = -acc / L(i, i); }
  }
  return X;
}
struct LU { Mat a; std::vector<uword> piv; int sign; bool singular; };
inline LU lu_decompose(const Mat& A) {  // LU with partial pivoting (LAPACK dgetrf)
  if (A.n_rows != A.n_cols) throw std::logic_error("inv()/det(): given matrix must be square sized");
  LU f{A, {}, 1, false}; const uword n = A.n_rows; f.piv.resize(n);
  for (uword k = 0; k < n; k++) {
    uword p = k; double best = std::fabs(f.a(k, k));
    for (uword i = k + 1; i < n; i++) if (std::fabs(f.a(i, k)) > best) { best = std::fabs(f.a(i, k)); p = i; }
    f.piv[k] = p;
    if (best == 0.0 || !(best == best)) { f.singular = true; if (!(best == best)) return f; continue; }
    if (p != k) { for (uword j = 0; j < n; j++) std::swap(f.a(k, j), f.a(p, j)); f.sign = -f.sign; }
    for (uword i = k + 1; i < n; i++) { f.a(i, k) /= f.a(k, k); for (uword j = k + 1; j < n; j++) f.a(i, j) -= f.a(i, k) * f.a(k, j); }
  }
  return f;
}
inline double det(const Mat& A) { LU f = lu_decompose(A); double d = f.sign; for (uword i = 0; i < A.n_rows; i++) d *= f.a(i, i); return d; }
inline Mat inv(const Mat& A) {
  LU f = lu_decompose(A); const uword n = A.n_rows;
  if (f.singular) throw std::runtime_error("inv(): matrix is singular");
  Mat X(n, n);
  for (uword c = 0; c < n; c++) {
    std::vector<double> b(n, 0.0); b[c] = 1.0;
    for (uword k = 0; k < n; k++) std::swap(b[k], b[f.piv[k]]);
    for (uword i = 0; i < n; i++) for (uword k = 0; k < i; k++) b[i] -= f.a(i, k) * b[k];
    for (uword ii = n; ii-- > 0;) { for (uword k = ii + 1; k < n; k++) b[ii] -= f.a(ii, k) * b[k]; b[ii] /= f.a(ii, ii); }
    for (uword i = 0; i < n; i++) X(i, c) = b[i];
  }
  return X;
}
inline Mat chol(const Mat& A, const char* layout) {  // LAPACK dpotrf; throws like arma::chol when A is not positive definite
  if (A.n_rows != A.n_cols) throw std::logic_error("chol(): given matrix must be square sized");
  const uword n = A.n_rows; Mat L(n, n);
  for (uword j = 0; j < n; j++) {
    double d = A(j, j); for (uword k = 0; k < j; k++) d -= L(j, k) * L(j, k);
    if (!(d > 0.0)) throw std::runtime_error("chol(): decomposition failed");
    L(j, j) = std::sqrt(d);
    for (uword i = j + 1; i < n; i++) { double s = A(i, j); for (uword k = 0; k < j; k++) s -= L(i, k) * L(j, k); L(i, j) = s / L(j, j); }
  }
  return std::string(layout) == "lower" ? L : trans(L);
}
inline bool qr_econ(Mat& Q, Mat& R, const Mat& X) {  // Householder QR (LAPACK dgeqrf + dorgqr), X is rows x cols with rows >= cols
  const uword mr = X.n_rows, nc = X.n_cols; Mat A = X; std::vector<std::vector<double>> vs; std::vector<double> betas;
  for (uword k = 0; k < std::min(mr, nc); k++) {
    double norm = 0.0; for (uword i = k; i < mr; i++) norm += A(i, k) * A(i, k); norm = std::sqrt(norm);
    std::vector<double> v(mr, 0.0); double beta = 0.0;
    if (norm != 0.0) {
      const double alpha = A(k, k) >= 0 ? -norm : norm;   // LAPACK sign choice: R(k,k) has the sign opposite to A(k,k)
      for (uword i = k; i < mr; i++) v[i] = A(i, k); v[k] -= alpha;
      double vv = 0.0; for (uword i = k; i < mr; i++) vv += v[i] * v[i];
      beta = vv != 0.0 ? 2.0 / vv : 0.0;
      for (uword j = k; j < nc; j++) { double s = 0.0; for (uword i = k; i < mr; i++) s += v[i] * A(i, j); s *= beta; for (uword i = k; i < mr; i++) A(i, j) -= s * v[i]; }
    }
    vs.push_back(v); betas.push_back(beta);
  }
  const uword kk = std::min(mr, nc);
  R = Mat(kk, nc); for (uword j = 0; j < nc; j++) for (uword i = 0; i <= std::min(j, kk - 1); i++) R(i, j) = A(i, j);
  Q = Mat(mr, kk); for (uword i = 0; i < kk; i++) Q(i, i) = 1.0;
  for (uword k = kk; k-- > 0;) for (uword j = 0; j < kk; j++) { double s = 0.0; for (uword i = k; i < mr; i++) s += vs[k][i] * Q(i, j); s *= betas[k]; for (uword i = k; i < mr; i++) Q(i, j) -= s * vs[k][i]; }
  return true;
}

}  // namespace arma

// =====================================================================================================================
namespace Rcpp {

struct Obj;
typedef std::shared_ptr<Obj> SEXP;
struct Obj {
  enum Type { NIL, REAL, LGL, STR, LIST } type = NIL;
  std::vector<double> real; std::vector<int> lgl; std::vector<std::string> str; std::vector<SEXP> list;
  std::vector<std::string> names; int nrow = -1, ncol = -1;
};
inline SEXP make(Obj::Type t) { SEXP s = std::make_shared<Obj>(); s->type = t; return s; }
inline SEXP deep_copy(const SEXP& s) { if (!s) return s; SEXP r = std::make_shared<Obj>(*s); for (auto& e : r->list) e = deep_copy(e); return r; }

[[noreturn]] inline void stop(const std::string& m) { throw std::runtime_error(m); }
inline std::vector<std::string>& warnings() { static std::vector<std::string> w; return w; }
inline void warning(const std::string& m) { warnings().push_back(m); }
inline void checkUserInterrupt() {}
struct NullStream { template <class T> NullStream& operator<<(const T&) { return *this; } NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } };
static NullStream Rcout;
struct Placeholder {}; static const Placeholder _ = Placeholder();

struct String {  // one element of a character vector
  std::string s;
  String() {} String(const std::string& x) : s(x) {} String(const char* x) : s(x) {}
  bool operator==(const String& o) const { return s == o.s; }
  bool operator==(const char* o) const { return s == o; }
  operator std::string() const { return s; }
};
inline std::string operator+(const std::string& a, const String& b) { return a + b.s; }
inline std::string operator+(const char* a, const String& b) { return std::string(a) + b.s; }

class LogicalVector; class StringVector; class NumericVector;

class LogicalVector {
 public:
  SEXP x;
  LogicalVector() : x(make(Obj::LGL)) {} explicit LogicalVector(SEXP s) : x(s) { if (!x || x->type != Obj::LGL) stop("not a logical vector"); }
  int length() const { return (int)x->lgl.size(); }
  int& operator()(int i) { return x->lgl.at(i); } int operator()(int i) const { return x->lgl.at(i); }
  struct Subset { SEXP x; std::vector<int> idx;
    Subset& operator=(bool v) { for (int i : idx) x->lgl.at(i) = v; return *this; }
    operator LogicalVector() const { LogicalVector r; for (int i : idx) r.x->lgl.push_back(x->lgl.at(i)); return r; } };
  Subset operator[](const LogicalVector& mask) { if (mask.length() != length()) stop("logical subsetting requires vectors of identical size"); Subset s{x, {}}; for (int i = 0; i < length(); i++) if (mask(i)) s.idx.push_back(i); return s; }
};
struct AnyResult { bool v; };
inline AnyResult any(const LogicalVector& l) { for (int i = 0; i < l.length(); i++) if (l(i)) return AnyResult{true}; return AnyResult{false}; }
inline bool is_false(const AnyResult& a) { return !a.v; }
inline bool is_true(const AnyResult& a) { return a.v; }

class StringVector {
 public:
  SEXP x;
  StringVector() : x(make(Obj::STR)) {} explicit StringVector(SEXP s) : x(s) { if (!x || x->type != Obj::STR) stop("not a character vector"); }
  int length() const { return (int)x->str.size(); } int size() const { return length(); }
  String operator()(int i) const { return String(x->str.at(i)); }
  StringVector operator[](const LogicalVector& mask) const { if (mask.length() != length()) stop("logical subsetting requires vectors of identical size"); StringVector r; for (int i = 0; i < length(); i++) if (mask(i)) r.x->str.push_back(x->str[i]); return r; }
};
typedef StringVector CharacterVector;
inline StringVector unique(const StringVector& v) { StringVector r; for (auto& s : v.x->str) if (std::find(r.x->str.begin(), r.x->str.end(), s) == r.x->str.end()) r.x->str.push_back(s); return r; }

struct NamesProxy {  // x.names() as an lvalue and an rvalue
  SEXP x;
  operator StringVector() const { StringVector r; r.x->str = x->names; return r; }
  NamesProxy& operator=(const StringVector& v) { x->names = v.x->str; return *this; }
  NamesProxy& operator=(const String& s) { x->names.assign(1, s.s); return *this; }
  NamesProxy& operator=(const NamesProxy& o) { x->names = o.x->names; return *this; }
};

class NumericVector {
 public:
  SEXP x;
  NumericVector() : x(make(Obj::REAL)) {}
  explicit NumericVector(SEXP s) : x(s) { if (!x || x->type != Obj::REAL) stop("not a numeric vector"); }
  NumericVector(int n) : x(make(Obj::REAL)) { x->real.assign(n, 0.0); }
  NumericVector(int n, double v) : x(make(Obj::REAL)) { x->real.assign(n, v); }
  int length() const { return (int)x->real.size(); } int size() const { return length(); }
  double& operator()(int i) { return x->real.at(i); } double operator()(int i) const { return x->real.at(i); }
  double& operator[](int i) { return x->real.at(i); } double operator[](int i) const { return x->real.at(i); }
  NamesProxy names() const { return NamesProxy{x}; }
  NumericVector operator[](const LogicalVector& mask) const { if (mask.length() != length()) stop("logical subsetting requires vectors of identical size"); NumericVector r; for (int i = 0; i < length(); i++) if (mask(i)) r.x->real.push_back(x->real[i]); return r; }
};
inline LogicalVector operator==(const NumericVector& v, double s) { LogicalVector r; for (double e : v.x->real) r.x->lgl.push_back(e == s); return r; }
inline NumericVector unique(const NumericVector& v) { NumericVector r; for (double e : v.x->real) if (std::find(r.x->real.begin(), r.x->real.end(), e) == r.x->real.end()) r.x->real.push_back(e); return r; }
inline NumericVector operator-(const NumericVector& a, const NumericVector& b) { if (a.length() != b.length()) stop("sugar: size mismatch"); NumericVector r(a.length()); for (int i = 0; i < a.length(); i++) r(i) = a(i) - b(i); return r; }
inline NumericVector operator/(const NumericVector& a, const NumericVector& b) { if (a.length() != b.length()) stop("sugar: size mismatch"); NumericVector r(a.length()); for (int i = 0; i < a.length(); i++) r(i) = a(i) / b(i); return r; }

class NumericMatrix {
 public:
  SEXP x;
  NumericMatrix() : x(make(Obj::REAL)) { x->nrow = x->ncol = 0; }
  explicit NumericMatrix(SEXP s) : x(s) { if (!x || x->type != Obj::REAL || x->nrow < 0) stop("not a numeric matrix"); }
  NumericMatrix(int r, int c) : x(make(Obj::REAL)) { x->real.assign((size_t)r * c, 0.0); x->nrow = r; x->ncol = c; }
  void fill(double v) { std::fill(x->real.begin(), x->real.end(), v); }
  double& operator()(int i, int j) { if (i < 0 || i >= x->nrow || j < 0 || j >= x->ncol) stop("subscript out of bounds"); return x->real[i + (size_t)j * x->nrow]; }
  NumericVector operator()(const Placeholder&, int j) const { NumericVector r(x->nrow); for (int i = 0; i < x->nrow; i++) r(i) = x->real.at(i + (size_t)j * x->nrow); return r; }
};

// ---- conversions between R objects and C++ / Armadillo values -----------------------------------------------------------
template <class T, class E = void> struct Conv;
template <> struct Conv<double> { static double from(const SEXP& s) { if (!s) stop("NULL"); if (s->type == Obj::REAL && s->real.size() == 1) return s->real[0]; if (s->type == Obj::LGL && s->lgl.size() == 1) return s->lgl[0]; stop("Expecting a single value"); }
                                   static SEXP to(double v) { SEXP s = make(Obj::REAL); s->real.assign(1, v); return s; } };
template <> struct Conv<int> { static int from(const SEXP& s) { return (int)Conv<double>::from(s); } static SEXP to(int v) { return Conv<double>::to(v); } };
template <> struct Conv<bool> { static bool from(const SEXP& s) { return Conv<double>::from(s) != 0.0; } static SEXP to(bool v) { SEXP s = make(Obj::LGL); s->lgl.assign(1, v); return s; } };
template <> struct Conv<std::string> { static std::string from(const SEXP& s) { if (!s || s->type != Obj::STR || s->str.size() != 1) stop("Expecting a single string value"); return s->str[0]; }
                                        static SEXP to(const std::string& v) { SEXP s = make(Obj::STR); s->str.assign(1, v); return s; } };
template <> struct Conv<arma::Mat> {
  static arma::Mat from(const SEXP& s) { if (!s || s->type != Obj::REAL) stop("Not a matrix."); const bool m = s->nrow >= 0; arma::Mat r(m ? s->nrow : s->real.size(), m ? s->ncol : 1); r.mem = s->real; return r; }
  static SEXP to(const arma::Mat& v) { SEXP s = make(Obj::REAL); s->real = v.mem; s->nrow = (int)v.n_rows; s->ncol = (int)v.n_cols; return s; } };
template <> struct Conv<arma::Col> { static arma::Col from(const SEXP& s) { if (!s || s->type != Obj::REAL) stop("Not a vector."); arma::Col r(s->real.size()); r.mem = s->real; return r; }
                                      static SEXP to(const arma::Col& v) { return Conv<arma::Mat>::to(v); } };   // RcppArmadillo wraps a colvec as an n x 1 matrix
template <> struct Conv<arma::Row> { static arma::Row from(const SEXP& s) { if (!s || s->type != Obj::REAL) stop("Not a vector."); arma::Row r(s->real.size()); r.mem = s->real; return r; }
                                      static SEXP to(const arma::Row& v) { return Conv<arma::Mat>::to(v); } };
template <> struct Conv<NumericVector> { static NumericVector from(const SEXP& s) { return NumericVector(s); } static SEXP to(const NumericVector& v) { return v.x; } };
template <> struct Conv<NumericMatrix> { static NumericMatrix from(const SEXP& s) { return NumericMatrix(s); } static SEXP to(const NumericMatrix& v) { return v.x; } };
template <> struct Conv<LogicalVector> { static LogicalVector from(const SEXP& s) { return LogicalVector(s); } static SEXP to(const LogicalVector& v) { return v.x; } };
template <> struct Conv<StringVector> { static StringVector from(const SEXP& s) { return StringVector(s); } static SEXP to(const StringVector& v) { return v.x; } };

class List; class DataFrame;
struct ElemProxy {  // list["name"] as an lvalue and an rvalue
  SEXP parent; std::string name;
  SEXP get() const { for (size_t i = 0; i < parent->names.size(); i++) if (parent->names[i] == name) return parent->list.at(i); stop("Index out of bounds: [index='" + name + "']."); }
  void set(const SEXP& v) { for (size_t i = 0; i < parent->names.size(); i++) if (parent->names[i] == name) { parent->list.at(i) = v; return; } parent->names.push_back(name); parent->list.push_back(v); }
  template <class T> operator T() const { return Conv<T>::from(get()); }
  template <class T> ElemProxy& operator=(const T& v) { set(Conv<T>::to(v)); return *this; }
};
class List {
 public:
  SEXP x;
  List() : x(make(Obj::LIST)) {}
  explicit List(SEXP s) : x(s) { if (!x || x->type == Obj::NIL) x = make(Obj::LIST); else if (x->type != Obj::LIST) stop("not a list"); }
  ElemProxy operator[](const std::string& n) const { return ElemProxy{x, n}; }
  ElemProxy operator[](const char* n) const { return ElemProxy{x, n}; }
  ElemProxy operator[](const String& n) const { return ElemProxy{x, n.s}; }
  NamesProxy names() const { return NamesProxy{x}; }
  int length() const { return (int)x->list.size(); }
  struct NamedValue { std::string name; SEXP v; };
  static List create(const NamedValue& a, const NamedValue& b, const NamedValue& c) { List l; for (auto* e : {&a, &b, &c}) { l.x->names.push_back(e->name); l.x->list.push_back(e->v); } return l; }
};
class DataFrame : public List { public: DataFrame() {} explicit DataFrame(SEXP s) : List(s) {} };
template <> struct Conv<List> { static List from(const SEXP& s) { return List(s); } static SEXP to(const List& v) { return v.x; } };
template <> struct Conv<DataFrame> { static DataFrame from(const SEXP& s) { return DataFrame(s); } static SEXP to(const DataFrame& v) { return v.x; } };
struct Named { std::string n; explicit Named(const std::string& s) : n(s) {} template <class T> List::NamedValue operator=(const T& v) const { return List::NamedValue{n, Conv<T>::to(v)}; } };

template <class T> T as(const ElemProxy& p) { return Conv<T>::from(p.get()); }
template <class T> T as(const NumericVector& v) { return Conv<T>::from(v.x); }
template <class T> T as(const List& v) { return Conv<T>::from(v.x); }
template <class T> T clone(const T& v) { T r = v; r.x = deep_copy(v.x); return r; }

}  // namespace Rcpp
