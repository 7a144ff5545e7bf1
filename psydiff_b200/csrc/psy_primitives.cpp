// psy_primitives.cpp — the filter primitives psydiff exports to R one by one (src/exposeInlineFunctions.cpp:13-308,
// registered in src/RcppExports.cpp:292-316), as plain C functions on column-major buffers.
//
// These are small host utilities kept so R code that calls e.g. psydiff::cholupdate() or computeIndividualM2LL()
// keeps working against the drop-in library.  They are NOT part of the fit / gradient path: that path is the CUDA
// kernel in psy_filter.cuh and never calls into this file.
#include "../../include/psydiff_b200.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {
inline size_t ix(int i, int j, int ld) { return (size_t)i + (size_t)j * ld; }

// C (r x c) = A (r x k) * B (k x c), optionally with B transposed (B given as c x k)
void gemm(int r, int k, int c, const double* A, const double* B, bool bt, double* C) {
  for (int j = 0; j < c; j++)
    for (int i = 0; i < r; i++) {
      double acc = 0;
      for (int q = 0; q < k; q++) acc += A[ix(i, q, r)] * (bt ? B[ix(j, q, c)] : B[ix(q, j, k)]);
      C[ix(i, j, r)] = acc;
    }
}

// Gauss-Jordan with partial pivoting: inverse and determinant of a general n x n matrix
bool inv_det(int n, const double* Ain, double* inv, double* det) {
  std::vector<double> A(Ain, Ain + (size_t)n * n), B((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) B[ix(i, i, n)] = 1.0;
  double d = 1.0;
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++) if (std::fabs(A[ix(i, k, n)]) > std::fabs(A[ix(p, k, n)])) p = i;
    if (A[ix(p, k, n)] == 0.0) { if (det) *det = 0.0; return false; }
    if (p != k) {
      d = -d;
      for (int j = 0; j < n; j++) { std::swap(A[ix(k, j, n)], A[ix(p, j, n)]); std::swap(B[ix(k, j, n)], B[ix(p, j, n)]); }
    }
    const double piv = A[ix(k, k, n)];
    d *= piv;
    for (int j = 0; j < n; j++) { A[ix(k, j, n)] /= piv; B[ix(k, j, n)] /= piv; }
    for (int i = 0; i < n; i++) {
      if (i == k) continue;
      const double f = A[ix(i, k, n)];
      if (f == 0.0) continue;
      for (int j = 0; j < n; j++) { A[ix(i, j, n)] -= f * A[ix(k, j, n)]; B[ix(i, j, n)] -= f * B[ix(k, j, n)]; }
    }
  }
  if (inv) std::memcpy(inv, B.data(), sizeof(double) * n * n);
  if (det) *det = d;
  return true;
}

// inverse of the lower triangle of L (arma::inv(arma::trimatl(L)))
void tril_inv(int n, const double* L, double* Li) {
  std::fill(Li, Li + (size_t)n * n, 0.0);
  for (int j = 0; j < n; j++) {
    Li[ix(j, j, n)] = 1.0 / L[ix(j, j, n)];
    for (int i = j + 1; i < n; i++) {
      double acc = 0;
      for (int k = j; k < i; k++) acc += L[ix(i, k, n)] * Li[ix(k, j, n)];
      Li[ix(i, j, n)] = -acc / L[ix(i, i, n)];
    }
  }
}
}  // namespace

extern "C" {

// getMeanWeights_C / getCovWeights_C (inst/include/psydiffInternals.h:18-33)
void psy_getMeanWeights(int n, double alpha, double beta, double kappa, double* w) {
  (void)beta;
  const double lambda = alpha * alpha * (n + kappa) - n;
  for (int i = 0; i < 2 * n + 1; i++) w[i] = 1 / (2 * (n + lambda));
  w[0] = lambda / (n + lambda);
}
void psy_getCovWeights(int n, double alpha, double beta, double kappa, double* w) {
  const double lambda = alpha * alpha * (n + kappa) - n;
  for (int i = 0; i < 2 * n + 1; i++) w[i] = 1 / (2 * (n + lambda));
  w[0] = lambda / (n + lambda) + (1 - alpha * alpha + beta);
}
// getWMatrix_C (:35-44)
void psy_getWMatrix(int s, const double* wm, const double* wc, double* W) {
  for (int j = 0; j < s; j++)
    for (int i = 0; i < s; i++) {
      double acc = 0;
      for (int k = 0; k < s; k++) acc += ((i == k ? 1.0 : 0.0) - wm[i]) * wc[k] * ((j == k ? 1.0 : 0.0) - wm[j]);
      W[ix(i, j, s)] = acc;
    }
}
// getSigmaPoints_C (:4-16).  Returns 1 if chol(A) fails (covariance_is_root == 0 and A not positive definite).
int psy_getSigmaPoints(int n, const double* m, const double* Ain, int is_root, double c, double* X) {
  std::vector<double> A(Ain, Ain + (size_t)n * n);
  if (!is_root) {
    std::vector<double> Lc((size_t)n * n, 0.0);
    for (int j = 0; j < n; j++) {
      double d = Ain[ix(j, j, n)];
      for (int k = 0; k < j; k++) d -= Lc[ix(j, k, n)] * Lc[ix(j, k, n)];
      if (!(d > 0)) return PSY_ERR_ARG;
      Lc[ix(j, j, n)] = std::sqrt(d);
      for (int i = j + 1; i < n; i++) {
        double a = Ain[ix(i, j, n)];
        for (int k = 0; k < j; k++) a -= Lc[ix(i, k, n)] * Lc[ix(j, k, n)];
        Lc[ix(i, j, n)] = a / Lc[ix(j, j, n)];
      }
    }
    A = Lc;
  }
  const double sc = std::sqrt(c);
  for (int i = 0; i < n; i++) X[ix(i, 0, n)] = m[i];
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) {
      X[ix(i, 1 + j, n)] = m[i] + sc * A[ix(i, j, n)];
      X[ix(i, 1 + n + j, n)] = m[i] - sc * A[ix(i, j, n)];
    }
  return PSY_OK;
}
// computeMeanFromSigmaPoints_C (:46-49)
void psy_computeMeanFromSigmaPoints(int n, int s, const double* X, const double* wm, double* mean) {
  for (int i = 0; i < n; i++) {
    double a = 0;
    for (int k = 0; k < s; k++) a += X[ix(i, k, n)] * wm[k];
    mean[i] = a;
  }
}
// getAFromSigmaPoints_C (:51-60)
void psy_getAFromSigmaPoints(int n, const double* X, const double* wm, double c, double* A) {
  std::vector<double> mean(n);
  psy_computeMeanFromSigmaPoints(n, 2 * n + 1, X, wm, mean.data());
  const double sc = std::sqrt(c);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) A[ix(i, j, n)] = (X[ix(i, 1 + j, n)] - mean[i]) / sc;
}
// computePFromSigmaPoints_C (:62-67)
void psy_computePFromSigmaPoints(int n, const double* X, const double* wm, double c, double* P) {
  std::vector<double> A((size_t)n * n);
  psy_getAFromSigmaPoints(n, X, wm, c, A.data());
  gemm(n, n, n, A.data(), A.data(), true, P);
}
// getPhi_C (:69-75)
void psy_getPhi(int n, const double* Mx, double* Phi) {
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) Phi[ix(i, j, n)] = (i > j) ? Mx[ix(i, j, n)] : (i == j ? .5 * Mx[ix(i, j, n)] : 0.0);
}
// predictMu_C (:77-79)
void psy_predictMu(int m, int s, const double* Y, const double* wm, double* mu) { psy_computeMeanFromSigmaPoints(m, s, Y, wm, mu); }
// predictS_C (:81-83)
void psy_predictS(int m, int s, const double* Y, const double* W, const double* R, double* S) {
  std::vector<double> YW((size_t)m * s);
  gemm(m, s, s, Y, W, false, YW.data());
  gemm(m, s, m, YW.data(), Y, true, S);
  for (int i = 0; i < m * m; i++) S[i] += R[i];
}
// predictC_C (:85-88)
void psy_predictC(int n, int m, int s, const double* X, const double* Y, const double* W, double* Cm) {
  std::vector<double> XW((size_t)n * s);
  gemm(n, s, s, X, W, false, XW.data());
  gemm(n, s, m, XW.data(), Y, true, Cm);
}
// computeK_C (:90-92)
int psy_computeK(int n, int m, const double* Cm, const double* S, double* K) {
  std::vector<double> Si((size_t)m * m);
  if (!inv_det(m, S, Si.data(), nullptr)) return PSY_ERR_ARG;
  gemm(n, m, m, Cm, Si.data(), false, K);
  return PSY_OK;
}
// updateM_C (:94-97)
void psy_updateM(int n, int m, const double* mp, const double* K, const double* res, double* mu) {
  for (int i = 0; i < n; i++) {
    double a = 0;
    for (int k = 0; k < m; k++) a += K[ix(i, k, n)] * res[k];
    mu[i] = mp[i] + a;
  }
}
// updateP_C (:99-101)
void psy_updateP(int n, int m, const double* Pp, const double* K, const double* S, double* Pu) {
  std::vector<double> KS((size_t)n * m), KSK((size_t)n * n);
  gemm(n, m, m, K, S, false, KS.data());
  gemm(n, m, n, KS.data(), K, true, KSK.data());
  for (int i = 0; i < n * n; i++) Pu[i] = Pp[i] - KSK[i];
}
// computeIndividualM2LL_C (:103-115)
int psy_computeIndividualM2LL(int o, const double* raw, const double* mean, const double* cov, double* out) {
  std::vector<double> Si((size_t)o * o), r(o);
  double det = 0;
  const bool ok = inv_det(o, cov, Si.data(), &det);
  for (int a = 0; a < o; a++) r[a] = raw[a] - mean[a];
  double dist = 0;
  if (ok)
    for (int a = 0; a < o; a++) {
      double z = 0;
      for (int k = 0; k < o; k++) z += Si[ix(a, k, o)] * r[k];
      dist += r[a] * z;
    }
  else dist = NAN;
  *out = o * std::log(2 * M_PI) + std::log(det) + dist;
  return PSY_OK;
}
// computeIndividualM2LLChol_C (:117-131)
int psy_computeIndividualM2LLChol(int o, const double* raw, const double* mean, const double* chol, double* out) {
  std::vector<double> Li((size_t)o * o);
  tril_inv(o, chol, Li.data());
  double ld = 0, dist = 0;
  for (int a = 0; a < o; a++) ld += std::log(chol[ix(a, a, o)]);
  for (int a = 0; a < o; a++) {
    double z = 0;
    for (int k = 0; k <= a; k++) z += Li[ix(a, k, o)] * (raw[k] - mean[k]);
    dist += z * z;
  }
  *out = o * std::log(2 * M_PI) + 2 * ld + dist;
  return PSY_OK;
}
// cholupdate_C (:133-165): v enters as v^(1/4), |v| if negative; unknown direction is an error (Rcpp::stop, :145)
int psy_cholupdate(int n, double* L, const double* xin, double v, const char* direction) {
  int dir;
  const std::string d(direction ? direction : "");
  if (d == "update") dir = 1;
  else if (d == "downdate") dir = -1;
  else return PSY_ERR_ARG;
  if (v < 0) v = -v;
  const double sc = std::pow(v, .25);
  std::vector<double> x(n);
  for (int i = 0; i < n; i++) x[i] = sc * xin[i];
  for (int k = 0; k < n; k++) {
    const double lkk = L[ix(k, k, n)];
    const double r = std::sqrt(lkk * lkk + dir * x[k] * x[k]);
    const double c = r / lkk, s = x[k] / lkk;
    L[ix(k, k, n)] = r;
    for (int i = k + 1; i < n; i++) {
      L[ix(i, k, n)] = (L[ix(i, k, n)] + dir * s * x[i]) / c;
      x[i] = c * x[i] - s * L[ix(i, k, n)];
    }
  }
  return PSY_OK;
}
// qr_C (:167-171): R of the economy QR, by Givens rotations (sign convention: non-negative diagonal)
void psy_qr(int rows, int cols, const double* X, double* R) {
  std::fill(R, R + (size_t)cols * cols, 0.0);
  std::vector<double> t(cols);
  for (int r = 0; r < rows; r++) {
    for (int j = 0; j < cols; j++) t[j] = X[ix(r, j, rows)];
    for (int k = 0; k < cols; k++) {
      const double a = R[ix(k, k, cols)], b = t[k];
      const double h = std::sqrt(a * a + b * b);
      if (h == 0.0) continue;
      const double c = a / h, s = b / h;
      R[ix(k, k, cols)] = h;
      for (int j = k + 1; j < cols; j++) {
        const double rkj = R[ix(k, j, cols)];
        R[ix(k, j, cols)] = c * rkj + s * t[j];
        t[j] = c * t[j] - s * rkj;
      }
    }
  }
}
// predictSquareRootOfS_C (:174-198)
void psy_predictSquareRootOfS(int m, int s, const double* Y, const double* mu, const double* Rchol, double wc0, double wc1, double* srS) {
  const int rows = s + m;
  std::vector<double> Tt((size_t)rows * m), R((size_t)m * m), d0(m);
  const double sq = std::sqrt(wc1);
  for (int co = 0; co < s; co++)
    for (int a = 0; a < m; a++) Tt[ix(co, a, rows)] = sq * (Y[ix(a, co, m)] - mu[a]);
  for (int b = 0; b < m; b++)
    for (int a = 0; a < m; a++) Tt[ix(s + b, a, rows)] = Rchol[ix(a, b, m)];
  psy_qr(rows, m, Tt.data(), R.data());
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) srS[ix(i, j, m)] = R[ix(j, i, m)];
  for (int a = 0; a < m; a++) d0[a] = Y[ix(a, 0, m)] - mu[a];
  psy_cholupdate(m, srS, d0.data(), wc0 < 0 ? -wc0 : wc0, wc0 < 0 ? "downdate" : "update");
  for (int j = 0; j < m; j++)
    for (int i = 0; i < j; i++) srS[ix(i, j, m)] = 0.0;
}
// computeK_SR_C (:200-203)
int psy_computeK_SR(int n, int m, const double* Cm, const double* srS, double* K) {
  std::vector<double> Li((size_t)m * m), T((size_t)n * m);
  tril_inv(m, srS, Li.data());
  gemm(n, m, m, Cm, Li.data(), true, T.data());
  gemm(n, m, m, T.data(), Li.data(), false, K);
  return PSY_OK;
}
// logChol2Chol_C (:205-209)
void psy_logChol2Chol(int n, const double* lc, double* out) {
  for (int i = 0; i < n * n; i++) out[i] = lc[i];
  for (int i = 0; i < n; i++) out[ix(i, i, n)] = std::exp(lc[ix(i, i, n)]);
}

}  // extern "C"
