// psy_oracle.cpp — CPU ORACLE for the psydiff SR-UKF likelihood / FD-gradient path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under psydiff_b200/ may include, link, import or call this file.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
// only as the checker or as the timed CPU arm.
//
// PARITY STATUS: the reference (jhorzek/psydiff, R + Rcpp + RcppArmadillo + Boost.odeint) holds no golden vectors for this
// path (inst/tinytest/test_psydiff.R:3-4 is `expect_equal(1 + 1, 2)`) and cannot be built the way its authors build it
// in this image (no R, Rcpp, Armadillo, Boost, LAPACK).  This file is a dependency-free restatement of the reference's
// algorithm.  It is PINNED TO THE REFERENCE'S OWN C++ run here: oracle/reference_build.py compiles the fitModel /
// getGradients text of R/modelTemplate.R and inst/include/psydiff*.h from /root/reference against stand-in headers for the
// third-party interfaces only (oracle/refbuild/include) and tests/test_reference_build.py compares this file with it
// (-2LL to 3e-14 on the linear model, 6e-12 on the nonlinear C4 model; states; gradients; every primitive); the same
// outputs are committed as tests/golden/*_reference.json.  What stays a restatement — there and here — is third-party
// arithmetic: Boost.odeint (unpinned CRAN `BH`) integrate_const<runge_kutta4> and integrate (controlled dopri5);
// Armadillo/LAPACK qr_econ, inv(trimatl), inv, det, chol.  Further anchors: (i) the exact discrete-time
// Kalman filter for linear models, (ii) SR == covariance variant for linear measurement equations, (iii) primitive
// identities, (iv) a numpy/LAPACK restatement, (v) this file's own 80-bit build.
//
// Every function cites the reference file:line (relative to /root/reference) it follows.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (see oracle/build.py).  All matrices column-major.

#include <cmath>
#include <cstring>
#include <cstdint>
#include <cfloat>
#include <thread>
#include <vector>
#include <algorithm>

#ifdef PSY_COUNT_FLOPS
static thread_local unsigned long long g_flops = 0;
#define FL(k) (g_flops += (k))
#else
#define FL(k) ((void)0)
#endif

namespace {

#ifndef PSY_REAL
#define PSY_REAL double
#endif
typedef PSY_REAL real;  // long double in the extended-precision build used to bound the oracle's own rounding noise

constexpr int NMAX = 12;            // max latent dimension
constexpr int SMAX = 2 * NMAX + 1;  // max sigma points
constexpr int MMAX = 16;            // max manifest dimension

#define IX(i, j, ld) ((i) + (j) * (ld))

// user equation signature emitted by oracle/codegen_host.py (mirrors R/prepareModel.R:479-505:
// latentEquation(x, pars, person, t) -> dx ; measurementEquation(x, pars, person, t) -> y)
typedef void (*user_fn)(const real* x, real* out, const real* plist, int person, real t);

struct Settings {
  int n, m;
  real alpha, beta, kappa;
  int stepper;       // 0 = rk4 (integrate_const), 1 = runge_kutta_dopri5 (integrate), 2 = "default"
  real h;          // timeStep[0]  (Q2: later entries can never repair a NaN state, R/modelTemplate.R:324-346)
  int sr;            // 1 = SR variant (R/modelTemplate.R:165-420), 0 = covariance variant (:583-847)
  int skip_update;   // fitModel(skipUpdate)
  // offsets of the special containers inside the flat parameter list
  int off_m0, off_A0, off_L, off_R;
  int plist_len;
};

struct Ctx {
  const Settings* s;
  user_fn drift, meas;
  const real* plist;
  int person;
  real c, sqrtc;
  real wm[SMAX], wc[SMAX];
  real L[NMAX * NMAX];   // logChol2Chol(LLogChol) (computeL, R/modelTemplate.R:77-83) — hoisted, constant per person
  real LLt[NMAX * NMAX]; // L*Qc*L' with Qc = I (R/modelTemplate.R:104, 132)
};

// inst/include/psydiffInternals.h:18-33  getMeanWeights_C / getCovWeights_C
void ut_weights(int n, real alpha, real beta, real kappa, real* wm, real* wc) {
  real lambda = alpha * alpha * (n + kappa) - n;
  int s = 2 * n + 1;
  for (int i = 0; i < s; i++) { wm[i] = 1 / (2 * (n + lambda)); wc[i] = 1 / (2 * (n + lambda)); }
  wm[0] = lambda / (n + lambda);
  wc[0] = lambda / (n + lambda) + (1 - alpha * alpha + beta);
}

// inst/include/psydiffInternals.h:35-44  getWMatrix_C : W = (I - wm 1')diag(wc)(I - wm 1')'
void w_matrix(int s, const real* wm, const real* wc, real* W) {
  std::vector<real> D(s * s);
  for (int j = 0; j < s; j++) for (int i = 0; i < s; i++) D[IX(i, j, s)] = (i == j ? 1.0 : 0.0) - wm[i];
  for (int j = 0; j < s; j++) for (int i = 0; i < s; i++) {
    real acc = 0; for (int k = 0; k < s; k++) acc += D[IX(i, k, s)] * wc[k] * D[IX(j, k, s)];
    W[IX(i, j, s)] = acc;
  }
}

// inst/include/psydiffInternals.h:4-16  getSigmaPoints_C (covarianceIsRoot = true)
void sigma_points(int n, const real* m, const real* A, real sqrtc, real* X) {
  for (int i = 0; i < n; i++) X[IX(i, 0, n)] = 0.0 + m[i];
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    X[IX(i, 1 + j, n)] = sqrtc * A[IX(i, j, n)] + m[i];
    X[IX(i, 1 + n + j, n)] = -1 * sqrtc * A[IX(i, j, n)] + m[i];
  }
  FL(4 * n * n);
}

// inst/include/psydiffInternals.h:46-60  computeMeanFromSigmaPoints_C + getAFromSigmaPoints_C (full n x n)
void get_A(int n, const real* X, const real* wm, real sqrtc, real* A) {
  int s = 2 * n + 1; real mean[NMAX];
  for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < s; k++) a += X[IX(i, k, n)] * wm[k]; mean[i] = a; }
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[IX(i, j, n)] = (X[IX(i, 1 + j, n)] - mean[i]) / sqrtc;
  FL(2 * n * s + 2 * n * n);
}

// arma::inv(arma::trimatl(A)) (R/modelTemplate.R:97, psydiffInternals.h:125,201): forward substitution on the
// lower triangle only (LAPACK dtrtri equivalent); strict upper of the result is exactly zero.
void tri_inv_lower(int n, const real* A, int ld, real* Ai) {
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) Ai[IX(i, j, n)] = 0.0;
    Ai[IX(j, j, n)] = 1.0 / A[IX(j, j, ld)];
    for (int i = j + 1; i < n; i++) {
      real acc = 0; for (int k = j; k < i; k++) acc += A[IX(i, k, ld)] * Ai[IX(k, j, n)];
      Ai[IX(i, j, n)] = -acc / A[IX(i, i, ld)];
    }
  }
  FL(n * n * n / 3 + n);
}

// odeintModel::operator() (R/modelTemplate.R:138-163) with getMMatrix (:85-135) and getPhi_C (Internals.h:69-75).
// The drift is evaluated ONCE per column (the reference evaluates it twice with identical results, :106-109, :155-157).
void rhs(const Ctx& cx, const real* X, real* dX, real t) {
  const int n = cx.s->n, s = 2 * n + 1;
  real A[NMAX * NMAX], Ai[NMAX * NMAX], F[NMAX * SMAX];
  get_A(n, X, cx.wm, cx.sqrtc, A);                      // :94 / :143 (same value both times)
  tri_inv_lower(n, A, n, Ai);                           // :97
  for (int co = 0; co < s; co++) cx.drift(X + co * n, F + co * n, cx.plist, cx.person, t);  // :106-109
  real mx[NMAX], mf[NMAX];
  for (int i = 0; i < n; i++) { mx[i] = 0; mf[i] = 0; }
  for (int co = 0; co < s; co++) for (int i = 0; i < n; i++) {   // :113-116
    mx[i] += cx.wm[co] * X[IX(i, co, n)]; mf[i] += cx.wm[co] * F[IX(i, co, n)];
  }
  real S1[NMAX * NMAX];  // sum1 + sum2 + L L'   (:119-133)
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    real s1 = 0, s2 = 0;
    for (int co = 0; co < s; co++) {
      s1 += cx.wc[co] * (X[IX(i, co, n)] - mx[i]) * (F[IX(j, co, n)] - mf[j]);
      s2 += cx.wc[co] * (F[IX(i, co, n)] - mf[i]) * (X[IX(j, co, n)] - mx[j]);
    }
    S1[IX(i, j, n)] = s1 + s2 + cx.LLt[IX(i, j, n)];
  }
  // M = invCov * S1 * invCov'
  real T1[NMAX * NMAX], M[NMAX * NMAX];
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    real a = 0; for (int k = 0; k < n; k++) a += Ai[IX(i, k, n)] * S1[IX(k, j, n)]; T1[IX(i, j, n)] = a;
  }
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    real a = 0; for (int k = 0; k < n; k++) a += T1[IX(i, k, n)] * Ai[IX(j, k, n)]; M[IX(i, j, n)] = a;
  }
  // Phi(M): trimatl with halved diagonal (Internals.h:69-75); APhi = A * Phi (full A as returned by getA, :150)
  real APhi[NMAX * NMAX];
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    real a = 0;
    for (int k = j; k < n; k++) { real ph = (k == j) ? 0.5 * M[IX(k, j, n)] : M[IX(k, j, n)]; a += A[IX(i, k, n)] * ph; }
    APhi[IX(i, j, n)] = a;
  }
  // dxdt.col(co) = sigmaPredict * meanWeights + sqrt(c) * augmented.col(co)  (:159-161)
  for (int i = 0; i < n; i++) dX[IX(i, 0, n)] = mf[i] + cx.sqrtc * 0.0;
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    dX[IX(i, 1 + j, n)] = mf[i] + cx.sqrtc * APhi[IX(i, j, n)];
    dX[IX(i, 1 + n + j, n)] = mf[i] + cx.sqrtc * (-1 * APhi[IX(i, j, n)]);
  }
  // op count of the minimal algorithm: 7ns + 2n^2 s + 6n^2 + 3n^3 on top of the s drift calls
  FL(7 * n * s + 2 * n * n * s + 6 * n * n + 3 * n * n * n);
}

// Boost.odeint integrate_const(runge_kutta4, sys, x, t0, t1, h) — call sites R/modelTemplate.R:327-329, 335-337.
// Rule T1: steps while (time + h) - t1 <= DBL_EPSILON, time = t0 + step*h, remainder dropped.
int rk4_steps(real t0, real t1, real h) {
  int step = 0; real time = t0;
  if (!(h > 0)) return 0;
  while ((time + h) - t1 <= DBL_EPSILON) { ++step; time = t0 + (real)step * h; if (step > 100000000) break; }
  return step;
}

int integrate_const_rk4(const Ctx& cx, real* X, real t0, real t1, real h) {
  const int n = cx.s->n, len = n * (2 * n + 1);
  real k1[NMAX * SMAX], k2[NMAX * SMAX], k3[NMAX * SMAX], k4[NMAX * SMAX], xt[NMAX * SMAX];
  int step = 0; real time = t0;
  while ((time + h) - t1 <= DBL_EPSILON) {
    // runge_kutta4::do_step (generic RK form: explicit zero coefficients are multiplied in too)
    rhs(cx, X, k1, time);
    for (int i = 0; i < len; i++) xt[i] = 1.0 * X[i] + (h * 0.5) * k1[i];
    rhs(cx, xt, k2, time + 0.5 * h);
    for (int i = 0; i < len; i++) xt[i] = 1.0 * X[i] + (h * 0.0) * k1[i] + (h * 0.5) * k2[i];
    rhs(cx, xt, k3, time + 0.5 * h);
    for (int i = 0; i < len; i++) xt[i] = 1.0 * X[i] + (h * 0.0) * k1[i] + (h * 0.0) * k2[i] + (h * 1.0) * k3[i];
    rhs(cx, xt, k4, time + h);
    for (int i = 0; i < len; i++)
      X[i] = 1.0 * X[i] + (h * (1.0 / 6.0)) * k1[i] + (h * (1.0 / 3.0)) * k2[i] + (h * (1.0 / 3.0)) * k3[i] + (h * (1.0 / 6.0)) * k4[i];
    FL(14 * len);
    ++step; time = t0 + (real)step * h;
  }
  return step;
}

// Boost.odeint integrate(sys, x, t0, t1, dt) = integrate_adaptive(controlled_runge_kutta<runge_kutta_dopri5>(1e-6,1e-6))
// — call sites R/modelTemplate.R:331-333, 340-342.  Rule T2.  Returns -1 on step-size underflow
// (odeint throws after too many failed tries) after poisoning X with NaN.
int integrate_dopri5(const Ctx& cx, real* X, real t0, real t1, real dt) {
  const int n = cx.s->n, len = n * (2 * n + 1);
  const real eps_abs = 1e-6, eps_rel = 1e-6;
  static const real a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9,
      a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729,
      a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
      c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  static const real dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
      dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  real k1[NMAX * SMAX], k2[NMAX * SMAX], k3[NMAX * SMAX], k4[NMAX * SMAX], k5[NMAX * SMAX], k6[NMAX * SMAX], k7[NMAX * SMAX];
  real xt[NMAX * SMAX], xn[NMAX * SMAX];
  real t = t0; int count = 0;
  rhs(cx, X, k1, t);  // fresh controlled stepper per call: FSAL derivative initialised here
  while (t1 - t > DBL_EPSILON) {
    if ((t + dt) - t1 > DBL_EPSILON) dt = t1 - t;
    int tries = 0;
    for (;;) {
      if (++tries > 500) { for (int i = 0; i < len; i++) X[i] = NAN; return -1; }
      for (int i = 0; i < len; i++) xt[i] = X[i] + dt * a21 * k1[i];
      rhs(cx, xt, k2, t + dt * (1.0 / 5));
      for (int i = 0; i < len; i++) xt[i] = X[i] + dt * a31 * k1[i] + dt * a32 * k2[i];
      rhs(cx, xt, k3, t + dt * (3.0 / 10));
      for (int i = 0; i < len; i++) xt[i] = X[i] + dt * a41 * k1[i] + dt * a42 * k2[i] + dt * a43 * k3[i];
      rhs(cx, xt, k4, t + dt * (4.0 / 5));
      for (int i = 0; i < len; i++) xt[i] = X[i] + dt * a51 * k1[i] + dt * a52 * k2[i] + dt * a53 * k3[i] + dt * a54 * k4[i];
      rhs(cx, xt, k5, t + dt * (8.0 / 9));
      for (int i = 0; i < len; i++) xt[i] = X[i] + dt * a61 * k1[i] + dt * a62 * k2[i] + dt * a63 * k3[i] + dt * a64 * k4[i] + dt * a65 * k5[i];
      rhs(cx, xt, k6, t + dt);
      for (int i = 0; i < len; i++) xn[i] = X[i] + dt * c1 * k1[i] + dt * c3 * k3[i] + dt * c4 * k4[i] + dt * c5 * k5[i] + dt * c6 * k6[i];
      rhs(cx, xn, k7, t + dt);
      real err = 0;
      bool bad = false;
      for (int i = 0; i < len; i++) {
        real xe = dt * dc1 * k1[i] + dt * dc3 * k3[i] + dt * dc4 * k4[i] + dt * dc5 * k5[i] + dt * dc6 * k6[i] + dt * dc7 * k7[i];
        real e = std::fabs(xe) / (eps_abs + eps_rel * (std::fabs(X[i]) + std::fabs(dt) * std::fabs(k1[i])));
        if (e != e) bad = true;
        if (e > err) err = e;
      }
      FL(45 * len);
      if (bad) { for (int i = 0; i < len; i++) X[i] = NAN; return -1; }  // NaN error: odeint would accept garbage; state is NaN either way
      if (err > 1.0) {
        dt *= std::max<real>(0.9 * std::pow(err, (real)(-1.0 / 3.0)), 0.2);
        continue;
      }
      t += dt;
      if (err < 0.5) { real e5 = std::max<real>(std::pow(5.0, -5.0), err); dt *= 0.9 * std::pow(e5, (real)(-1.0 / 5.0)); }
      std::memcpy(X, xn, sizeof(real) * len);
      std::memcpy(k1, k7, sizeof(real) * len);
      break;
    }
    ++count;
  }
  return count;
}

// qr_C (Internals.h:167-171): R factor of arma::qr_econ(X), X rows x cols (rows >= cols), via Householder
// reflections (LAPACK dgeqrf convention).  Signs of R's rows are implementation-defined and are normalised by the
// caller's cholupdate.  R is cols x cols upper triangular, column-major.
void qr_R(int rows, int cols, const real* Xin, real* R) {
  std::vector<real> Xv(Xin, Xin + (size_t)rows * cols);
  real* X = Xv.data();
  for (int k = 0; k < cols; k++) {
    real norm = 0; for (int i = k; i < rows; i++) norm += X[IX(i, k, rows)] * X[IX(i, k, rows)];
    norm = std::sqrt(norm);
    if (norm == 0.0) continue;
    real alpha = X[IX(k, k, rows)] > 0 ? -norm : norm;
    std::vector<real> v(rows, 0.0);
    for (int i = k; i < rows; i++) v[i] = X[IX(i, k, rows)];
    v[k] -= alpha;
    real vtv = 0; for (int i = k; i < rows; i++) vtv += v[i] * v[i];
    if (vtv == 0.0) continue;
    for (int j = k; j < cols; j++) {
      real dot = 0; for (int i = k; i < rows; i++) dot += v[i] * X[IX(i, j, rows)];
      real f = 2.0 * dot / vtv;
      for (int i = k; i < rows; i++) X[IX(i, j, rows)] -= f * v[i];
    }
  }
  for (int j = 0; j < cols; j++) for (int i = 0; i < cols; i++) R[IX(i, j, cols)] = (i <= j) ? X[IX(i, j, rows)] : 0.0;
  FL(2 * rows * cols * cols);
}

// cholupdate_C (Internals.h:133-165), literal — including the v^0.25 scaling (quirk Q3) and |v| for v<0.
// dir: +1 update, -1 downdate.  L is n x n column-major (full storage), x length n.
void cholupdate(int n, real* L, const real* xin, real v, int dir) {
  real x[MMAX > NMAX ? MMAX : NMAX];
  if (v < 0) v = -1 * v;
  v = std::pow(v, .25);
  for (int i = 0; i < n; i++) x[i] = v * xin[i];
  for (int k = 0; k < n; k++) {
    real r = std::sqrt(std::pow(L[IX(k, k, n)], 2.0) + dir * std::pow(x[k], 2.0));
    real c = r / L[IX(k, k, n)];
    real s = x[k] / L[IX(k, k, n)];
    L[IX(k, k, n)] = r;
    if (k < n - 1) {
      for (int i = k + 1; i < n; i++) L[IX(i, k, n)] = (L[IX(i, k, n)] + dir * s * x[i]) / c;
      for (int i = k + 1; i < n; i++) x[i] = c * x[i] - s * L[IX(i, k, n)];
    }
  }
  FL(3 * n * n + 8 * n);
}

// predictSquareRootOfS_C (Internals.h:174-198).  Y o x s (rows = non-missing manifests), mu o, Rchol o x o.
void predict_srS(int o, int s, const real* Y, const real* mu, const real* Rchol, real wc0, real wc1, real* srS) {
  int cols = s + o;
  std::vector<real> Tt((size_t)cols * o);  // trans(tempMat): (s+o) x o
  real sq = std::sqrt(wc1);
  for (int co = 0; co < s; co++) for (int i = 0; i < o; i++) Tt[IX(co, i, cols)] = sq * (Y[IX(i, co, o)] - mu[i]);
  for (int j = 0; j < o; j++) for (int i = 0; i < o; i++) Tt[IX(s + j, i, cols)] = Rchol[IX(i, j, o)];
  real R[MMAX * MMAX];
  qr_R(cols, o, Tt.data(), R);
  for (int j = 0; j < o; j++) for (int i = 0; i < o; i++) srS[IX(i, j, o)] = R[IX(j, i, o)];  // trans(R)
  real d0[MMAX]; for (int i = 0; i < o; i++) d0[i] = Y[IX(i, 0, o)] - mu[i];
  if (wc0 < 0) cholupdate(o, srS, d0, -1 * wc0, -1); else cholupdate(o, srS, d0, wc0, +1);
  for (int j = 0; j < o; j++) for (int i = 0; i < j; i++) srS[IX(i, j, o)] = 0.0;  // trimatl
}

// general inverse / determinant by LU with partial pivoting (arma::inv / arma::det → LAPACK dgetrf/dgetri),
// used by the covariance variant (Internals.h:90-92, 103-115).  Returns det; Ainv may be null.
real lu_inv_det(int n, const real* Ain, real* Ainv) {
  real A[MMAX * MMAX], B[MMAX * MMAX]; int piv[MMAX];
  for (int i = 0; i < n * n; i++) A[i] = Ain[i];
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) B[IX(i, j, n)] = (i == j);
  real det = 1;
  for (int k = 0; k < n; k++) {
    int p = k; real best = std::fabs(A[IX(k, k, n)]);
    for (int i = k + 1; i < n; i++) if (std::fabs(A[IX(i, k, n)]) > best) { best = std::fabs(A[IX(i, k, n)]); p = i; }
    piv[k] = p;
    if (p != k) { det = -det; for (int j = 0; j < n; j++) { std::swap(A[IX(k, j, n)], A[IX(p, j, n)]); std::swap(B[IX(k, j, n)], B[IX(p, j, n)]); } }
    real d = A[IX(k, k, n)]; det *= d;
    for (int i = k + 1; i < n; i++) {
      real f = A[IX(i, k, n)] / d; A[IX(i, k, n)] = f;
      for (int j = k + 1; j < n; j++) A[IX(i, j, n)] -= f * A[IX(k, j, n)];
      for (int j = 0; j < n; j++) B[IX(i, j, n)] -= f * B[IX(k, j, n)];
    }
  }
  if (Ainv) {
    for (int j = 0; j < n; j++) for (int i = n - 1; i >= 0; i--) {
      real a = B[IX(i, j, n)]; for (int k = i + 1; k < n; k++) a -= A[IX(i, k, n)] * Ainv[IX(k, j, n)];
      Ainv[IX(i, j, n)] = a / A[IX(i, i, n)];
    }
  }
  FL(2 * n * n * n);
  return det;
}

// arma::chol(P, "lower") (R/modelTemplate.R:720).  Returns false if not positive definite (Armadillo throws).
bool chol_lower(int n, const real* P, real* A) {
  for (int i = 0; i < n * n; i++) A[i] = 0;
  for (int j = 0; j < n; j++) {
    real d = P[IX(j, j, n)]; for (int k = 0; k < j; k++) d -= A[IX(j, k, n)] * A[IX(j, k, n)];
    if (!(d > 0)) return false;
    d = std::sqrt(d); A[IX(j, j, n)] = d;
    for (int i = j + 1; i < n; i++) {
      real a = P[IX(i, j, n)]; for (int k = 0; k < j; k++) a -= A[IX(i, k, n)] * A[IX(j, k, n)];
      A[IX(i, j, n)] = a / d;
    }
  }
  FL(n * n * n / 3);
  return true;
}

// logChol2Chol_C (Internals.h:205-209): exp() on the diagonal, everything else kept (full matrix)
void logchol2chol(int n, const real* lc, real* out) {
  for (int i = 0; i < n * n; i++) out[i] = lc[i];
  for (int i = 0; i < n; i++) out[IX(i, i, n)] = std::exp(lc[IX(i, i, n)]);
}

// One person: the body of the person loop of fitModel (R/modelTemplate.R:244-411 SR, :660-838 covariance).
// obs: T x m ROW-major slice for this person (obs[t*m + j]); dt: T.  latent (T x n) / pred (T x m) row-major, may be null.
real fit_person(const Settings& st, user_fn drift, user_fn meas, const real* plist, int person,
                  int T, const double* obs, const double* dts, double* latent, double* pred) {
  const int n = st.n, m = st.m, s = 2 * n + 1;
  Ctx cx; cx.s = &st; cx.drift = drift; cx.meas = meas; cx.plist = plist; cx.person = person;
  real kappa = st.kappa;
  if (n + kappa == 0) kappa -= 1;                                  // :192-196
  cx.c = std::pow(st.alpha, 2) * (n + kappa); cx.sqrtc = std::sqrt(cx.c);   // :197
  ut_weights(n, st.alpha, st.beta, kappa, cx.wm, cx.wc);            // :306-309
  std::vector<real> W((size_t)s * s); w_matrix(s, cx.wm, cx.wc, W.data());  // :310
  logchol2chol(n, plist + st.off_L, cx.L);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    real a = 0; for (int k = 0; k < n; k++) a += cx.L[IX(i, k, n)] * cx.L[IX(j, k, n)]; cx.LLt[IX(i, j, n)] = a;
  }
  real mvec[NMAX], m_[NMAX], A[NMAX * NMAX], P[NMAX * NMAX], P_[NMAX * NMAX], Rchol[MMAX * MMAX], Rfull[MMAX * MMAX];
  for (int i = 0; i < n; i++) mvec[i] = plist[st.off_m0 + i];      // :275
  logchol2chol(n, plist + st.off_A0, A);                            // :276
  logchol2chol(m, plist + st.off_R, Rchol);                         // :277-278
  if (!st.sr) {                                                     // :694-698
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < n; k++) a += A[IX(i, k, n)] * A[IX(j, k, n)]; P[IX(i, j, n)] = a; }
    for (int j = 0; j < m; j++) for (int i = 0; i < m; i++) { real a = 0; for (int k = 0; k < m; k++) a += Rchol[IX(i, k, m)] * Rchol[IX(j, k, m)]; Rfull[IX(i, j, m)] = a; }
  }
  real m2ll = 0.0, timeSum = 0.0;
  real X[NMAX * SMAX], Y[MMAX * SMAX], mu[MMAX];
  for (int tp = 0; tp < T; tp++) {
    timeSum += (real)dts[tp];                                             // :285
    real ob[MMAX]; for (int j = 0; j < m; j++) ob[j] = obs[(size_t)tp * m + j];
    int nm[MMAX], o = 0;
    for (int j = 0; j < m; j++) if (std::isfinite(ob[j])) nm[o++] = j;   // :291-293
    for (int i = 0; i < n; i++) m_[i] = mvec[i];                    // :298
    if (!st.sr) {                                                   // :718-720
      std::memcpy(P_, P, sizeof(real) * n * n);
      if (!chol_lower(n, P_, A)) { m2ll = NAN; break; }             // Armadillo throws → caller sees failure
    }
    sigma_points(n, m_, A, cx.sqrtc, X);                            // :303
    real endTime = (real)dts[tp];
    if (std::fabs(0.0 - endTime) > 0) {                             // :320 with fabs semantics (quirk Q1)
      if (st.stepper == 1) integrate_dopri5(cx, X, 0.0, endTime, st.h);
      else {
        integrate_const_rk4(cx, X, 0.0, endTime, st.h);
        if (st.stepper == 2) {                                      // "default": dopri5 on the (NaN) state, :338-343
          bool fin = true; for (int i = 0; i < n * s; i++) if (!std::isfinite(X[i])) fin = false;
          if (!fin) integrate_dopri5(cx, X, 0.0, endTime, st.h);
        }
      }
      for (int i = 0; i < n; i++) m_[i] = X[IX(i, 0, n)];           // :348
      real Af[NMAX * NMAX]; get_A(n, X, cx.wm, cx.sqrtc, Af);
      if (st.sr) {                                                  // :349  A = trimatl(getA(x))
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[IX(i, j, n)] = (i >= j) ? Af[IX(i, j, n)] : 0.0;
      } else {                                                      // :771  P_ = A A'
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < n; k++) a += Af[IX(i, k, n)] * Af[IX(j, k, n)]; P_[IX(i, j, n)] = a; }
      }
    }
    for (int co = 0; co < s; co++) meas(X + co * n, Y + co * m, plist, person, timeSum);   // getY_, :353-355, :62-73
    for (int j = 0; j < m; j++) { real a = 0; for (int co = 0; co < s; co++) a += Y[IX(j, co, m)] * cx.wm[co]; mu[j] = a; }  // :358
    FL(2 * m * s);
    if (pred) for (int j = 0; j < m; j++) pred[(size_t)tp * m + j] = (double)mu[j];   // :361
    if (o > 0 && !st.skip_update) {
      real Yo[MMAX * SMAX], muo[MMAX], Ro[MMAX * MMAX], res[MMAX];
      for (int co = 0; co < s; co++) for (int a = 0; a < o; a++) Yo[IX(a, co, o)] = Y[IX(nm[a], co, m)];
      for (int a = 0; a < o; a++) { muo[a] = mu[nm[a]]; res[a] = ob[nm[a]] - mu[nm[a]]; }   // :374
      // C = X * W * Yo'  (predictC_C, Internals.h:85-88), literal dense W
      real XW[NMAX * SMAX], C[NMAX * MMAX], K[NMAX * MMAX];
      for (int co = 0; co < s; co++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < s; k++) a += X[IX(i, k, n)] * W[IX(k, co, s)]; XW[IX(i, co, n)] = a; }
      for (int b = 0; b < o; b++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < s; k++) a += XW[IX(i, k, n)] * Yo[IX(b, k, o)]; C[IX(i, b, n)] = a; }
      FL(2 * n * o * s + 2 * n * s);
      if (st.sr) {
        for (int b = 0; b < o; b++) for (int a = 0; a < o; a++) Ro[IX(a, b, o)] = Rchol[IX(nm[a], nm[b], m)];
        real srS[MMAX * MMAX], Si[MMAX * MMAX];
        predict_srS(o, s, Yo, muo, Ro, cx.wc[0], cx.wc[1], srS);                      // :365
        tri_inv_lower(o, srS, o, Si);                                                 // computeK_SR_C, Internals.h:200-203
        real CSt[NMAX * MMAX];
        for (int b = 0; b < o; b++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < o; k++) a += C[IX(i, k, n)] * Si[IX(b, k, o)]; CSt[IX(i, b, n)] = a; }
        for (int b = 0; b < o; b++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < o; k++) a += CSt[IX(i, k, n)] * Si[IX(k, b, o)]; K[IX(i, b, n)] = a; }
        for (int i = 0; i < n; i++) { real a = 0; for (int b = 0; b < o; b++) a += K[IX(i, b, n)] * res[b]; mvec[i] = m_[i] + a; }   // :378
        real U[NMAX * MMAX];                                                        // :382-386
        for (int b = 0; b < o; b++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < o; k++) a += K[IX(i, k, n)] * srS[IX(k, b, o)]; U[IX(i, b, n)] = a; }
        for (int b = 0; b < o; b++) {
          cholupdate(n, A, U + b * n, 1.0, -1);
          for (int j = 0; j < n; j++) for (int i = 0; i < j; i++) A[IX(i, j, n)] = 0.0;
        }
        // computeIndividualM2LLChol_C (Internals.h:117-131)
        real ld = 0; for (int a = 0; a < o; a++) ld += std::log(srS[IX(a, a, o)]);
        real z[MMAX], dist = 0;
        for (int a = 0; a < o; a++) { real acc = 0; for (int k = 0; k < o; k++) acc += Si[IX(a, k, o)] * res[k]; z[a] = acc; }
        for (int a = 0; a < o; a++) dist += z[a] * z[a];
        m2ll += o * std::log(2 * M_PI) + 2 * ld + dist;                               // :390-392
        FL(4 * n * o * o + 2 * n * o + 2 * o * o + 4 * o);
      } else {
        real S[MMAX * MMAX], Sinv[MMAX * MMAX];
        // predictS_C (Internals.h:81-83): Yo W Yo' + R, then symmetrised (:790)
        real YW[MMAX * SMAX];
        for (int co = 0; co < s; co++) for (int a = 0; a < o; a++) { real acc = 0; for (int k = 0; k < s; k++) acc += Yo[IX(a, k, o)] * W[IX(k, co, s)]; YW[IX(a, co, o)] = acc; }
        for (int b = 0; b < o; b++) for (int a = 0; a < o; a++) { real acc = 0; for (int k = 0; k < s; k++) acc += YW[IX(a, k, o)] * Yo[IX(b, k, o)]; S[IX(a, b, o)] = acc + Rfull[IX(nm[a], nm[b], m)]; }
        for (int b = 0; b < o; b++) for (int a = 0; a < b; a++) { real v = (S[IX(a, b, o)] + S[IX(b, a, o)]) / 2; S[IX(a, b, o)] = v; S[IX(b, a, o)] = v; }
        real det = lu_inv_det(o, S, Sinv);
        for (int b = 0; b < o; b++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < o; k++) a += C[IX(i, k, n)] * Sinv[IX(k, b, o)]; K[IX(i, b, n)] = a; }   // computeK_C
        for (int i = 0; i < n; i++) { real a = 0; for (int b = 0; b < o; b++) a += K[IX(i, b, n)] * res[b]; mvec[i] = m_[i] + a; }
        real KS[NMAX * MMAX];                                                       // updateP_C (Internals.h:99-101) + symmetrise (:809)
        for (int b = 0; b < o; b++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < o; k++) a += K[IX(i, k, n)] * S[IX(k, b, o)]; KS[IX(i, b, n)] = a; }
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { real a = 0; for (int k = 0; k < o; k++) a += KS[IX(i, k, n)] * K[IX(j, k, n)]; P[IX(i, j, n)] = P_[IX(i, j, n)] - a; }
        for (int j = 0; j < n; j++) for (int i = 0; i < j; i++) { real v = (P[IX(i, j, n)] + P[IX(j, i, n)]) / 2; P[IX(i, j, n)] = v; P[IX(j, i, n)] = v; }
        real z[MMAX], dist = 0;                                                     // computeIndividualM2LL_C (Internals.h:103-115)
        for (int a = 0; a < o; a++) { real acc = 0; for (int k = 0; k < o; k++) acc += Sinv[IX(a, k, o)] * res[k]; z[a] = acc; }
        for (int a = 0; a < o; a++) dist += res[a] * z[a];
        m2ll += o * std::log(2 * M_PI) + std::log(det) + dist;
      }
      if (latent) for (int i = 0; i < n; i++) latent[(size_t)tp * n + i] = (double)mvec[i];   // :379
    } else {
      for (int i = 0; i < n; i++) mvec[i] = m_[i];                                   // :404-408
      if (latent) for (int i = 0; i < n; i++) latent[(size_t)tp * n + i] = (double)mvec[i];
      if (!st.sr) {                                                                   // :832-834
        for (int j = 0; j < n; j++) for (int i = 0; i <= j; i++) { real v = (P_[IX(i, j, n)] + P_[IX(j, i, n)]) / 2; P[IX(i, j, n)] = v; P[IX(j, i, n)] = v; }
      }
    }
  }
  return m2ll;
}

struct Problem {
  Settings st; user_fn drift, meas;
  int N, F, P;
  const double* base_plist;   // plist_len
  const int* slot_off;          // F: flat offset of free slot k inside plist
  const int* theta_index;       // N x F row-major
  const int64_t* row_off;       // N+1
  const double* obs;          // rows x m ROW-major
  const double* dt;           // rows
  const int* person_id;         // N
};

// Stateless parameter semantics (quirk Q4): base values overwritten by THIS person's theta entries
// (setParameterList_C, inst/include/psydiffBasic.h:63-121, with every row flagged changed).
void person_plist(const Problem& pb, const double* theta, int p, real* plist) {
  for (int i = 0; i < pb.st.plist_len; i++) plist[i] = pb.base_plist[i];
  for (int k = 0; k < pb.F; k++) plist[pb.slot_off[k]] = theta[pb.theta_index[(size_t)p * pb.F + k]];
}

double fit_one(const Problem& pb, const double* theta, int p, double* latent, double* pred) {
  std::vector<real> plist(pb.st.plist_len);
  person_plist(pb, theta, p, plist.data());
  int64_t r0 = pb.row_off[p]; int T = (int)(pb.row_off[p + 1] - r0);
  return (double)fit_person(pb.st, pb.drift, pb.meas, plist.data(), pb.person_id[p], T, pb.obs + r0 * pb.st.m, pb.dt + r0,
                    latent ? latent + r0 * pb.st.n : nullptr, pred ? pred + r0 * pb.st.m : nullptr);
}

template <class Fn> void parallel_for(int N, int nthreads, Fn fn) {
  if (nthreads <= 1) { for (int i = 0; i < N; i++) fn(i); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back([=]() { for (int i = t; i < N; i += nthreads) fn(i); });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

struct psy_oracle_problem {
  int n, m; double alpha, beta, kappa; int stepper; double h; int sr; int skip_update;
  int off_m0, off_A0, off_L, off_R, plist_len;
  void* drift; void* meas;
  int N, F, P;
  const double* base_plist; const int* slot_off; const int* theta_index; const int64_t* row_off;
  const double* obs; const double* dt; const int* person_id;
};

static Problem to_problem(const psy_oracle_problem* q) {
  Problem pb;
  pb.st.n = q->n; pb.st.m = q->m; pb.st.alpha = q->alpha; pb.st.beta = q->beta; pb.st.kappa = q->kappa;
  pb.st.stepper = q->stepper; pb.st.h = q->h; pb.st.sr = q->sr; pb.st.skip_update = q->skip_update;
  pb.st.off_m0 = q->off_m0; pb.st.off_A0 = q->off_A0; pb.st.off_L = q->off_L; pb.st.off_R = q->off_R; pb.st.plist_len = q->plist_len;
  pb.drift = (user_fn)q->drift; pb.meas = (user_fn)q->meas;
  pb.N = q->N; pb.F = q->F; pb.P = q->P; pb.base_plist = q->base_plist; pb.slot_off = q->slot_off;
  pb.theta_index = q->theta_index; pb.row_off = q->row_off; pb.obs = q->obs; pb.dt = q->dt; pb.person_id = q->person_id;
  return pb;
}

// fitModel (R/modelTemplate.R:165-420 / :583-847): per-person -2LL (+ optional states, row-major).
// Quirk Q2: every person is evaluated; a failed person is NaN.  Returns 0.
int psy_oracle_fit(const psy_oracle_problem* q, const double* theta, double* m2ll, double* latent, double* pred, int nthreads) {
  Problem pb = to_problem(q);
  parallel_for(pb.N, nthreads, [&](int p) { m2ll[p] = fit_one(pb, theta, p, latent, pred); });
  return 0;
}

// getGradients (R/modelTemplate.R:852-930).  direction: 0 central, 1 left, 2 right.  Perturbs from a pristine copy of
// theta (quirk Q7).  Only persons whose theta_index touches `par` are re-filtered (dirty-flag semantics,
// psydiffBasic.h:22-26 / R/modelTemplate.R:248-254); all others keep their base -2LL.
// Every evaluation is independent of the ones before it: the reference's history effects — the shared parameter list (quirk
// Q4) and the NaN poisoning of a model object after one non-finite evaluation (quirk Q8: results are "reset" by x -= x,
// R/modelTemplate.R:255-261, so its eps[] fallback never recovers) — are deliberately not reproduced; both are pinned by
// tests/test_reference_build.py against the reference's own code.
// grad_ref : the reference's arithmetic (Σ right − Σ left)/Σ eps in double, sequential sums over persons.
// grad_clean: Σ_p (f+_p − f−_p) accumulated in long double — the parity target for the GPU path.
int psy_oracle_gradients(const psy_oracle_problem* q, const double* theta, const double* eps, int n_eps, int direction,
                         double* grad_ref, double* grad_clean, double* base_m2ll_out, int nthreads) {
  Problem pb = to_problem(q);
  const int N = pb.N, P = pb.P, F = pb.F;
  std::vector<double> base(N);
  parallel_for(N, nthreads, [&](int p) { base[p] = fit_one(pb, theta, p, nullptr, nullptr); });
  if (base_m2ll_out) std::memcpy(base_m2ll_out, base.data(), sizeof(double) * N);
  // persons touched by each parameter
  std::vector<std::vector<int>> touch(P);
  for (int p = 0; p < N; p++) for (int k = 0; k < F; k++) {
    int th = pb.theta_index[(size_t)p * F + k];
    if (touch[th].empty() || touch[th].back() != p) touch[th].push_back(p);
  }
  std::vector<double> th2(theta, theta + P);
  for (int par = 0; par < P; par++) {
    double side_sum[2] = {0, 0}, eps_used = 0; long double side_clean[2] = {0, 0};
    bool side_ok[2] = {true, true};
    for (int side = 0; side < 2; side++) {   // 0 = left (−eps), 1 = right (+eps)
      bool active = (direction == 0) || (direction == 1 && side == 0) || (direction == 2 && side == 1);
      if (!active) {   // one-sided: the other side is the base fit (:872-879, :900-902, :920-922)
        double ssum = 0; for (int p = 0; p < N; p++) ssum += base[p];
        side_sum[side] = ssum; side_clean[side] = 0; continue;
      }
      bool ok = false;
      for (int e = 0; e < n_eps && !ok; e++) {
        std::vector<double> thp(th2);
        thp[par] = (side == 0) ? th2[par] - eps[e] : th2[par] + eps[e];
        std::vector<double> f(touch[par].size());
        const std::vector<int>& tl = touch[par];
        parallel_for((int)tl.size(), nthreads, [&](int i) { f[i] = fit_one(pb, thp.data(), tl[i], nullptr, nullptr); });
        // Σ over ALL persons in person order (sum(m2LL), :891-892)
        double ssum = 0; size_t ti = 0; long double cl = 0;
        for (int p = 0; p < N; p++) {
          if (ti < tl.size() && tl[ti] == p) { ssum += f[ti]; cl += (long double)f[ti] - (long double)base[p]; ti++; }
          else ssum += base[p];
        }
        if (std::isfinite(ssum)) { side_sum[side] = ssum; side_clean[side] = cl; eps_used += eps[e]; ok = true; }
      }
      if (!ok) side_ok[side] = false;
    }
    if (!side_ok[0] || !side_ok[1]) { grad_ref[par] = NAN; grad_clean[par] = NAN; continue; }
    grad_ref[par] = (side_sum[1] - side_sum[0]) / eps_used;                          // :926
    grad_clean[par] = (double)((side_clean[1] - side_clean[0]) / (long double)eps_used);
  }
  return 0;
}

int psy_oracle_rk4_steps(double t0, double t1, double h) { return rk4_steps(t0, t1, h); }

unsigned long long psy_oracle_flops_reset(void) {
#ifdef PSY_COUNT_FLOPS
  unsigned long long v = g_flops; g_flops = 0; return v;
#else
  return 0ULL;
#endif
}

#ifndef PSY_EXT  // the primitive wrappers exist in the double build only
// ---- primitives, for unit tests of the oracle itself (src/exposeInlineFunctions.cpp names) ----
void psy_oracle_getMeanWeights(int n, double alpha, double beta, double kappa, double* wm) { double wc[SMAX]; ut_weights(n, alpha, beta, kappa, wm, wc); }
void psy_oracle_getCovWeights(int n, double alpha, double beta, double kappa, double* wc) { double wm[SMAX]; ut_weights(n, alpha, beta, kappa, wm, wc); }
void psy_oracle_getWMatrix(int s, const double* wm, const double* wc, double* W) { w_matrix(s, wm, wc, W); }
void psy_oracle_getSigmaPoints(int n, const double* m, const double* A, double c, double* X) { sigma_points(n, m, A, std::sqrt(c), X); }
void psy_oracle_getAFromSigmaPoints(int n, const double* X, const double* wm, double c, double* A) { get_A(n, X, wm, std::sqrt(c), A); }
void psy_oracle_cholupdate(int n, double* L, const double* x, double v, int dir) { cholupdate(n, L, x, v, dir); }
void psy_oracle_qr(int rows, int cols, const double* X, double* R) { qr_R(rows, cols, X, R); }
void psy_oracle_predictSquareRootOfS(int o, int s, const double* Y, const double* mu, const double* Rchol, double wc0, double wc1, double* srS) { predict_srS(o, s, Y, mu, Rchol, wc0, wc1, srS); }
double psy_oracle_computeIndividualM2LL(int o, const double* raw, const double* mean, const double* cov) {
  double Sinv[MMAX * MMAX], z[MMAX], res[MMAX], dist = 0; double det = lu_inv_det(o, cov, Sinv);
  for (int a = 0; a < o; a++) res[a] = raw[a] - mean[a];
  for (int a = 0; a < o; a++) { double acc = 0; for (int k = 0; k < o; k++) acc += Sinv[IX(a, k, o)] * res[k]; z[a] = acc; }
  for (int a = 0; a < o; a++) dist += res[a] * z[a];
  return o * std::log(2 * M_PI) + std::log(det) + dist;
}

#endif

}  // extern "C"
