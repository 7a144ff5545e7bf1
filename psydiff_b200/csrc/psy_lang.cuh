// psy_lang.cuh — device-side micro-library the user's drift / measurement statements are compiled against.
//
// psydiff lets users write Armadillo statements such as `dx = DRIFT*x;`, `y = MANIFESTMEANS + LAMBDA*x;`,
// `dx(1) = cint + a*x(0) + b*x(1);`, `dx[0] = x[0]*(alpha - beta*x[1]);`  (R/prepareModel.R:80-86, 172-179,
// 257-260, 345-355; binding rules R/prepareModel.R:737-770).  On the GPU these become fixed-size, fully unrolled
// register arithmetic: every Vec/Mat here has compile-time extents, every index the generated code uses is a
// literal, so after inlining nothing lives in memory.
//
// Supported surface: x(i) / x[i] element access on vectors, M(i,j) on matrices, + - * between matrices, vectors
// and scalars (matrix*vector, matrix*matrix, scalar scaling), element-wise % and /, unary minus, assignment of a
// scalar to a 1-vector (`y = x(0);`), arma::exp/log/sin/cos/tanh/sqrt/abs/square/pow on vectors, arma::dot/accu/sum/mean/norm/max/
// min/as_scalar/trans, arma::expm of a square matrix (the vignette's example,
// vignettes/psydiff-basics.Rmd:94-98), and every CUDA <cmath> function on scalars.  Initialised temporaries
// `arma::mat T = <expr>;` / `arma::colvec v = <expr>;` are accepted: the code generator turns the declared type into `auto`,
// so T gets the expression's compile-time extents.  Anything else (size-only declarations such as `arma::mat T(2,2);`,
// other Armadillo functions) does not compile and is reported by compileModel as "unsupported on device".
#pragma once

namespace psy {

template <int N>
struct Vec {
  double v[N];
  __device__ __forceinline__ double& operator()(int i) { return v[i]; }
  __device__ __forceinline__ const double& operator()(int i) const { return v[i]; }
  __device__ __forceinline__ double& operator[](int i) { return v[i]; }
  __device__ __forceinline__ const double& operator[](int i) const { return v[i]; }
  // Armadillo: assigning a scalar to a colvec makes it 1x1 with that value (`y = x(0);`, R/prepareModel.R:253-255)
  __device__ __forceinline__ Vec& operator=(double s) {
    static_assert(N == 1, "scalar assignment to a vector is only defined for length-1 vectors");
    v[0] = s;
    return *this;
  }
  __device__ __forceinline__ void fill(double s) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = s;
  }
  __device__ __forceinline__ void zeros() { fill(0.0); }
};

template <int R, int C>
struct Mat {
  double a[R * C];  // column-major, like arma::mat
  __device__ __forceinline__ double& operator()(int i, int j) { return a[i + j * R]; }
  __device__ __forceinline__ const double& operator()(int i, int j) const { return a[i + j * R]; }
  __device__ __forceinline__ double& operator()(int i) { return a[i]; }
  __device__ __forceinline__ const double& operator()(int i) const { return a[i]; }
  __device__ __forceinline__ double& operator[](int i) { return a[i]; }
  __device__ __forceinline__ const double& operator[](int i) const { return a[i]; }
};

// a C x 1 matrix container used as a vector in `MANIFESTMEANS + LAMBDA*x`
template <int R>
__device__ __forceinline__ Vec<R> as_vec(const Mat<R, 1>& m) {
  Vec<R> r;
#pragma unroll
  for (int i = 0; i < R; i++) r.v[i] = m.a[i];
  return r;
}

#define PSY_VEC_BINOP(OP)                                                                     \
  template <int N>                                                                            \
  __device__ __forceinline__ Vec<N> operator OP(const Vec<N>& a, const Vec<N>& b) {           \
    Vec<N> r;                                                                                 \
    _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = a.v[i] OP b.v[i];                  \
    return r;                                                                                 \
  }                                                                                           \
  template <int N>                                                                            \
  __device__ __forceinline__ Vec<N> operator OP(const Mat<N, 1>& a, const Vec<N>& b) {        \
    Vec<N> r;                                                                                 \
    _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = a.a[i] OP b.v[i];                  \
    return r;                                                                                 \
  }                                                                                           \
  template <int N>                                                                            \
  __device__ __forceinline__ Vec<N> operator OP(const Vec<N>& a, const Mat<N, 1>& b) {        \
    Vec<N> r;                                                                                 \
    _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = a.v[i] OP b.a[i];                  \
    return r;                                                                                 \
  }                                                                                           \
  template <int R, int C>                                                                     \
  __device__ __forceinline__ Mat<R, C> operator OP(const Mat<R, C>& a, const Mat<R, C>& b) {  \
    Mat<R, C> r;                                                                              \
    _Pragma("unroll") for (int i = 0; i < R * C; i++) r.a[i] = a.a[i] OP b.a[i];              \
    return r;                                                                                 \
  }
PSY_VEC_BINOP(+)
PSY_VEC_BINOP(-)
#undef PSY_VEC_BINOP

// element-wise product / quotient (arma `%` and `/`)
template <int N>
__device__ __forceinline__ Vec<N> operator%(const Vec<N>& a, const Vec<N>& b) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = a.v[i] * b.v[i];
  return r;
}
template <int N>
__device__ __forceinline__ Vec<N> operator/(const Vec<N>& a, const Vec<N>& b) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = a.v[i] / b.v[i];
  return r;
}

// scalar scaling / shifting
template <int N>
__device__ __forceinline__ Vec<N> operator*(double s, const Vec<N>& a) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = s * a.v[i];
  return r;
}
template <int N>
__device__ __forceinline__ Vec<N> operator*(const Vec<N>& a, double s) { return s * a; }
template <int N>
__device__ __forceinline__ Vec<N> operator/(const Vec<N>& a, double s) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = a.v[i] / s;
  return r;
}
template <int N>
__device__ __forceinline__ Vec<N> operator+(const Vec<N>& a, double s) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = a.v[i] + s;
  return r;
}
template <int N>
__device__ __forceinline__ Vec<N> operator+(double s, const Vec<N>& a) { return a + s; }
template <int N>
__device__ __forceinline__ Vec<N> operator-(const Vec<N>& a, double s) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = a.v[i] - s;
  return r;
}
template <int N>
__device__ __forceinline__ Vec<N> operator-(const Vec<N>& a) {
  Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = -a.v[i];
  return r;
}
template <int R, int C>
__device__ __forceinline__ Mat<R, C> operator*(double s, const Mat<R, C>& a) {
  Mat<R, C> r;
#pragma unroll
  for (int i = 0; i < R * C; i++) r.a[i] = s * a.a[i];
  return r;
}
template <int R, int C>
__device__ __forceinline__ Mat<R, C> operator*(const Mat<R, C>& a, double s) { return s * a; }
template <int R, int C>
__device__ __forceinline__ Mat<R, C> operator-(const Mat<R, C>& a) {
  Mat<R, C> r;
#pragma unroll
  for (int i = 0; i < R * C; i++) r.a[i] = -a.a[i];
  return r;
}

// matrix * vector, matrix * matrix.  Accumulation order k = 0..C-1 like a textbook gemv.
template <int R, int C>
__device__ __forceinline__ Vec<R> operator*(const Mat<R, C>& A, const Vec<C>& x) {
  Vec<R> r;
#pragma unroll
  for (int i = 0; i < R; i++) {
    double acc = A.a[i] * x.v[0];
#pragma unroll
    for (int k = 1; k < C; k++) acc += A.a[i + k * R] * x.v[k];
    r.v[i] = acc;
  }
  return r;
}
template <int R, int K, int C>
__device__ __forceinline__ Mat<R, C> operator*(const Mat<R, K>& A, const Mat<K, C>& B) {
  Mat<R, C> r;
#pragma unroll
  for (int j = 0; j < C; j++)
#pragma unroll
    for (int i = 0; i < R; i++) {
      double acc = A.a[i] * B.a[j * K];
#pragma unroll
      for (int k = 1; k < K; k++) acc += A.a[i + k * R] * B.a[k + j * K];
      r.a[i + j * R] = acc;
    }
  return r;
}

}  // namespace psy

// the handful of arma:: free functions that show up in model code
namespace arma {
template <int N>
__device__ __forceinline__ psy::Vec<N> exp(const psy::Vec<N>& a) {
  psy::Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = ::exp(a.v[i]);
  return r;
}
template <int N>
__device__ __forceinline__ psy::Vec<N> log(const psy::Vec<N>& a) {
  psy::Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = ::log(a.v[i]);
  return r;
}
// element-wise <cmath> functions on vectors: arma::sin(x), arma::tanh(x), arma::sqrt(x), arma::abs(x), arma::square(x), arma::pow(x, p)
#define PSY_ARMA_ELEMENTWISE(NAME, EXPR)                                       \
  template <int N>                                                             \
  __device__ __forceinline__ psy::Vec<N> NAME(const psy::Vec<N>& a) {          \
    psy::Vec<N> r;                                                             \
    _Pragma("unroll") for (int i = 0; i < N; i++) { const double v = a.v[i]; r.v[i] = (EXPR); } \
    return r;                                                                  \
  }
PSY_ARMA_ELEMENTWISE(sin, ::sin(v))
PSY_ARMA_ELEMENTWISE(cos, ::cos(v))
PSY_ARMA_ELEMENTWISE(tanh, ::tanh(v))
PSY_ARMA_ELEMENTWISE(sqrt, ::sqrt(v))
PSY_ARMA_ELEMENTWISE(abs, ::fabs(v))
PSY_ARMA_ELEMENTWISE(square, v * v)
#undef PSY_ARMA_ELEMENTWISE
template <int N>
__device__ __forceinline__ psy::Vec<N> pow(const psy::Vec<N>& a, double p) {
  psy::Vec<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = ::pow(a.v[i], p);
  return r;
}
// reductions: arma::norm (2-norm), max, min, mean, as_scalar of a one-element object
template <int N>
__device__ __forceinline__ double norm(const psy::Vec<N>& a) {
  double acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) acc += a.v[i] * a.v[i];
  return ::sqrt(acc);
}
template <int N>
__device__ __forceinline__ double max(const psy::Vec<N>& a) {
  double m = a.v[0];
#pragma unroll
  for (int i = 1; i < N; i++) m = (a.v[i] > m) ? a.v[i] : m;
  return m;
}
template <int N>
__device__ __forceinline__ double min(const psy::Vec<N>& a) {
  double m = a.v[0];
#pragma unroll
  for (int i = 1; i < N; i++) m = (a.v[i] < m) ? a.v[i] : m;
  return m;
}
template <int N>
__device__ __forceinline__ double mean(const psy::Vec<N>& a) {
  double acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) acc += a.v[i];
  return acc / (double)N;
}
__device__ __forceinline__ double as_scalar(const psy::Vec<1>& a) { return a.v[0]; }
__device__ __forceinline__ double as_scalar(const psy::Mat<1, 1>& a) { return a.a[0]; }
template <int N>
__device__ __forceinline__ double dot(const psy::Vec<N>& a, const psy::Vec<N>& b) {
  double acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) acc += a.v[i] * b.v[i];
  return acc;
}
template <int N>
__device__ __forceinline__ double accu(const psy::Vec<N>& a) {
  double acc = 0;
#pragma unroll
  for (int i = 0; i < N; i++) acc += a.v[i];
  return acc;
}
template <int N>
__device__ __forceinline__ double sum(const psy::Vec<N>& a) { return accu(a); }
// arma::expm: scaling and squaring around a degree-12 Taylor polynomial in Horner form.  ||A / 2^s||_1 <= 1/2 makes the
// truncation error < 2e-14 relative; Armadillo's Pade-based expm agrees with it to rounding.
template <int N>
__device__ __forceinline__ psy::Mat<N, N> expm(const psy::Mat<N, N>& A) {
  double nrm = 0.0;
#pragma unroll
  for (int j = 0; j < N; j++) {
    double cs = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) cs += fabs(A.a[i + j * N]);
    nrm = fmax(nrm, cs);
  }
  int s = 0;
  double sc = 1.0;
  while (nrm * sc > 0.5 && s < 64) { sc *= 0.5; s++; }
  const psy::Mat<N, N> B = sc * A;
  psy::Mat<N, N> E;
#pragma unroll
  for (int i = 0; i < N * N; i++) E.a[i] = (i % (N + 1) == 0) ? 1.0 : 0.0;
#pragma unroll 1
  for (int k = 12; k >= 1; k--) {
    E = (1.0 / (double)k) * (B * E);
#pragma unroll
    for (int i = 0; i < N; i++) E.a[i + i * N] += 1.0;
  }
#pragma unroll 1
  for (int q = 0; q < s; q++) E = E * E;
  return E;
}
template <int R, int C>
__device__ __forceinline__ psy::Mat<C, R> trans(const psy::Mat<R, C>& a) {
  psy::Mat<C, R> r;
#pragma unroll
  for (int j = 0; j < C; j++)
#pragma unroll
    for (int i = 0; i < R; i++) r.a[j + i * C] = a.a[i + j * R];
  return r;
}
}  // namespace arma
