// psy_filter.cuh — K1, the fused per-instance SR-UKF filter kernel (hand-written FP64 CUDA for sm_100a).
//
// One CUDA thread owns one filter instance = (person, finite-difference perturbation) and carries it through
// ALL of that person's time points: sigma points, the sigma-point matrix ODE (Särkkä 2007 eq. 34) under RK4,
// the square-root measurement update and the -2LL accumulation never leave registers.  This replaces, for the
// whole person x time loop, the reference's
//   fitModel                         R/modelTemplate.R:165-420  (SR variant)
//   odeintModel::operator()          R/modelTemplate.R:138-163
//   getMMatrix / computeL            R/modelTemplate.R:77-135
//   integrate_const<runge_kutta4>    call sites R/modelTemplate.R:327-329 (Boost.odeint)
//   getSigmaPoints_C, getAFromSigmaPoints_C, getPhi_C, predictMu_C, predictC_C, predictSquareRootOfS_C (qr_C,
//   cholupdate_C), computeK_SR_C, updateM_C, computeIndividualM2LLChol_C, logChol2Chol_C
//                                    inst/include/psydiffInternals.h:4-209
// and, through the instance list, the 2P+1 fitModel sweeps of getGradients (R/modelTemplate.R:852-930).
//
// B200-first formulation: the reference integrates the full n x (2n+1) sigma-point matrix X.
// Every column of dX/dt is  F·wm ± sqrt(c)·[A·Phi(M)]_j, so X keeps the form [m, m + sqrt(c)A, m − sqrt(c)A]
// for all t and RK4 (linear in its increments) commutes with that map.  The kernel therefore integrates the
// REDUCED state (m, A) — n + n(n+1)/2 doubles instead of n(2n+1) — with
//     dm/dt = Σ wm_i f(x_i),   dA/dt = A·Phi(M),
//     M = kap·(G + Gᵀ) + A⁻¹ L Lᵀ A⁻ᵀ,   G = A⁻¹ D,  D_j = f(m + sqrt(c)A_j) − f(m − sqrt(c)A_j),  kap = wc1·sqrt(c).
// This is algebraically identical to getMMatrix's Σ wc_i (x_i − m_x)(f_i − m_f)ᵀ + transpose (the i = 0 term
// vanishes because x_0 = m_x) and differs from the reference only by rounding (parity gate 1e-9 relative).
// It is what lets one instance fit a thread's register file for n <= 6.
//
// Missing manifests (NaN) are handled without dynamic indexing: a missing variable's deviations are zeroed and
// its row/column of Rchol replaced by the unit vector, which makes srS block-diagonal with a unit entry there,
// its Kalman-gain column exactly zero and its covariance downdate an exact no-op — the same numbers the
// reference gets from Y_.rows(nonmissing) / Rchol.submat(nonmissing, nonmissing) (R/modelTemplate.R:365-392).
#pragma once
#include "psy_launch.h"
#include "psy_lang.cuh"

#ifndef PSY_BLOCK
#define PSY_BLOCK 128
#endif
#ifndef PSY_MIN_BLOCKS
#define PSY_MIN_BLOCKS 1
#endif

namespace psy {

__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower, i >= j

enum { L_DIAG = 0, L_FULL = 1 };

template <class Mdl>
struct Filter {
  static constexpr int N = Mdl::N, M = Mdl::M, F = Mdl::NFREE;  // NFREE >= 1 (padded); NSLOT real slots
  static constexpr int TN = N * (N + 1) / 2, TM = M * (M + 1) / 2;
  static constexpr int QLEN = (Mdl::L_STRUCT == L_DIAG) ? N : TN;

  struct UT { double sqrtc, wm0, wm1, wc0, wc1, kap; };

  // ---- RHS of the reduced sigma-point ODE (replaces odeintModel::operator() + getMMatrix) -------------------
  static __device__ __forceinline__ void rhs(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                             double (&dm)[N], double (&dA)[TN], const double (&Q)[QLEN],
                                             const UT& ut, int person, double t) {
    // explicit inverse of the lower-triangular root (arma::inv(arma::trimatl(A)), R/modelTemplate.R:97)
    double rd[N], Ai[TN];
#pragma unroll
    for (int i = 0; i < N; i++) rd[i] = 1.0 / A[tri(i, i)];
#pragma unroll
    for (int j = 0; j < N; j++) {
      Ai[tri(j, j)] = rd[j];
#pragma unroll
      for (int i = j + 1; i < N; i++) {
        double acc = A[tri(i, j)] * rd[j];
#pragma unroll
        for (int k = j + 1; k < i; k++) acc += A[tri(i, k)] * Ai[tri(k, j)];
        Ai[tri(i, j)] = -acc * rd[i];
      }
    }
    // drift at the 2n+1 sigma points (latentEquation, R/prepareModel.R:493-505), one evaluation per column
    Vec<N> xc, f0;
#pragma unroll
    for (int i = 0; i < N; i++) { xc.v[i] = m[i]; f0.v[i] = 0.0; }
    Mdl::latentEquation(xc, f0, p, person, t);
    double fs[N], G[N * N];
#pragma unroll
    for (int i = 0; i < N; i++) fs[i] = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
      Vec<N> xp, xm, fp, fm;
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i >= j) {
          xp.v[i] = fma(ut.sqrtc, A[tri(i, j)], m[i]);
          xm.v[i] = fma(-ut.sqrtc, A[tri(i, j)], m[i]);
        } else {
          xp.v[i] = m[i];
          xm.v[i] = m[i];
        }
        fp.v[i] = 0.0;
        fm.v[i] = 0.0;
      }
      Mdl::latentEquation(xp, fp, p, person, t);
      Mdl::latentEquation(xm, fm, p, person, t);
      double d[N];
#pragma unroll
      for (int i = 0; i < N; i++) {
        d[i] = fp.v[i] - fm.v[i];
        fs[i] += fp.v[i] + fm.v[i];
      }
      // G[:,j] = A⁻¹ d
#pragma unroll
      for (int i = 0; i < N; i++) {
        double acc = Ai[tri(i, 0)] * d[0];
#pragma unroll
        for (int k = 1; k <= i; k++) acc += Ai[tri(i, k)] * d[k];
        G[i + j * N] = acc;
      }
    }
#pragma unroll
    for (int i = 0; i < N; i++) dm[i] = ut.wm0 * f0.v[i] + ut.wm1 * fs[i];

    // lower triangle of M = kap (G + Gᵀ) + A⁻¹ (L Lᵀ) A⁻ᵀ, with Phi's halved diagonal folded in (getPhi_C)
    double Ph[TN];
    if constexpr (Mdl::L_STRUCT == L_DIAG) {
      // Q = diag(l_k²):  (A⁻¹ Q A⁻ᵀ)_ij = Σ_{k<=j} Ai_ik Ai_jk l_k²
      double Bq[TN];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int k = 0; k <= i; k++) Bq[tri(i, k)] = Ai[tri(i, k)] * Q[k];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          double acc = Bq[tri(i, 0)] * Ai[tri(j, 0)];
#pragma unroll
          for (int k = 1; k <= j; k++) acc += Bq[tri(i, k)] * Ai[tri(j, k)];
          double g = G[i + j * N] + G[j + i * N];
          double mij = ut.kap * g + acc;
          Ph[tri(i, j)] = (i == j) ? 0.5 * mij : mij;
        }
    } else {
      // general L: Q = L Lᵀ (symmetric, packed lower).  Z = A⁻¹ Q (full), M_Q = Z A⁻ᵀ (lower part)
      double Z[N * N];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k <= i; k++) acc += Ai[tri(i, k)] * ((k >= j) ? Q[tri(k, j)] : Q[tri(j, k)]);
          Z[i + j * N] = acc;
        }
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k <= j; k++) acc += Z[i + k * N] * Ai[tri(j, k)];
          double g = G[i + j * N] + G[j + i * N];
          double mij = ut.kap * g + acc;
          Ph[tri(i, j)] = (i == j) ? 0.5 * mij : mij;
        }
    }
    // dA = A · Phi(M)   (augmentedMatrix, R/modelTemplate.R:147-151; the sqrt(c) cancels against X_j = m + sqrt(c)A_j)
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double acc = A[tri(i, j)] * Ph[tri(j, j)];
#pragma unroll
        for (int k = j + 1; k <= i; k++) acc += A[tri(i, k)] * Ph[tri(k, j)];
        dA[tri(i, j)] = acc;
      }
  }

  // ---- classical RK4, K fixed steps (Boost.odeint integrate_const<runge_kutta4>, rule T1) -------------------
  static __device__ __forceinline__ void rk4(const double (&p)[F], double (&m)[N], double (&A)[TN],
                                             const double (&Q)[QLEN], const UT& ut, int person, double h, int K) {
#pragma unroll 1
    for (int step = 0; step < K; step++) {
      const double t0 = (double)step * h;
      double mt[N], At[TN], ma[N], Aa[TN];
#pragma unroll
      for (int i = 0; i < N; i++) { mt[i] = m[i]; ma[i] = m[i]; }
#pragma unroll
      for (int i = 0; i < TN; i++) { At[i] = A[i]; Aa[i] = A[i]; }
#pragma unroll 1
      for (int st = 0; st < 4; st++) {
        double km[N], kA[TN];
        const double tc = (st == 0) ? t0 : ((st == 3) ? t0 + h : t0 + 0.5 * h);
        rhs(p, mt, At, km, kA, Q, ut, person, tc);
        const double b = (st == 0 || st == 3) ? h * (1.0 / 6.0) : h * (1.0 / 3.0);
        const double a = (st == 2) ? h : 0.5 * h;
#pragma unroll
        for (int i = 0; i < N; i++) { ma[i] += b * km[i]; mt[i] = m[i] + a * km[i]; }
#pragma unroll
        for (int i = 0; i < TN; i++) { Aa[i] += b * kA[i]; At[i] = A[i] + a * kA[i]; }
      }
#pragma unroll
      for (int i = 0; i < N; i++) m[i] = ma[i];
#pragma unroll
      for (int i = 0; i < TN; i++) A[i] = Aa[i];
    }
  }

  // ---- measurement update, SR variant (R/modelTemplate.R:352-408).  Returns this time point's -2LL term. -----
  static __device__ __forceinline__ double update_sr(const double (&p)[F], double (&m)[N], double (&A)[TN],
                                                     const double (&ob)[M], const UT& ut, int person, double tabs,
                                                     bool skip_update, double* pred_row) {
    // sigma points of (m, A) and their images (getY_, R/modelTemplate.R:62-73); t is ABSOLUTE time here
    Vec<M> Y0, Yp[N], Ym[N];
    {
      Vec<N> xc;
#pragma unroll
      for (int i = 0; i < N; i++) xc.v[i] = m[i];
      Y0.zeros();
      Mdl::measurementEquation(xc, Y0, p, person, tabs);
#pragma unroll
      for (int j = 0; j < N; j++) {
        Vec<N> xp, xm;
#pragma unroll
        for (int i = 0; i < N; i++) {
          if (i >= j) {
            double d = ut.sqrtc * A[tri(i, j)];
            xp.v[i] = m[i] + d;
            xm.v[i] = m[i] - d;
          } else {
            xp.v[i] = m[i];
            xm.v[i] = m[i];
          }
        }
        Yp[j].zeros();
        Ym[j].zeros();
        Mdl::measurementEquation(xp, Yp[j], p, person, tabs);
        Mdl::measurementEquation(xm, Ym[j], p, person, tabs);
      }
    }
    double mu[M];
#pragma unroll
    for (int a = 0; a < M; a++) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) acc += Yp[j].v[a] + Ym[j].v[a];
      mu[a] = ut.wm0 * Y0.v[a] + ut.wm1 * acc;  // predictMu_C
    }
    if (pred_row) {
#pragma unroll
      for (int a = 0; a < M; a++) pred_row[a] = mu[a];
    }
    bool miss[M];
    int o = 0;
#pragma unroll
    for (int a = 0; a < M; a++) {
      miss[a] = !isfinite(ob[a]);
      o += miss[a] ? 0 : 1;
    }
    if (o == 0 || skip_update) return 0.0;

    // srS = chol factor of wc1 Σ_i (Y_i − mu)(·)ᵀ + Rchol Rcholᵀ  (predictSquareRootOfS_C: QR of [√wc1 (Y−mu), Rchol]ᵀ).
    // The QR is done as a streaming Givens triangularisation, one row of the stacked matrix at a time; its R
    // factor has a positive diagonal, which is the sign convention cholupdate_C normalises to anyway (T3).
    double S[TM];
#pragma unroll
    for (int i = 0; i < TM; i++) S[i] = 0.0;
    const double sw = sqrt(ut.wc1);
    auto rotate_in = [&](double (&tr)[M]) {
#pragma unroll
      for (int k = 0; k < M; k++) {
        const double a = S[tri(k, k)], b = tr[k];
        const double r = sqrt(a * a + b * b);
        double c = 1.0, s = 0.0;
        if (r != 0.0) {
          const double ri = 1.0 / r;
          c = a * ri;
          s = b * ri;
        }
        S[tri(k, k)] = r;
#pragma unroll
        for (int j = k + 1; j < M; j++) {
          const double rkj = S[tri(j, k)];
          S[tri(j, k)] = c * rkj + s * tr[j];
          tr[j] = c * tr[j] - s * rkj;
        }
      }
    };
    {
      double tr[M];
#pragma unroll
      for (int a = 0; a < M; a++) tr[a] = miss[a] ? 0.0 : sw * (Y0.v[a] - mu[a]);
      rotate_in(tr);
#pragma unroll
      for (int j = 0; j < N; j++) {
#pragma unroll
        for (int a = 0; a < M; a++) tr[a] = miss[a] ? 0.0 : sw * (Yp[j].v[a] - mu[a]);
        rotate_in(tr);
      }
#pragma unroll
      for (int j = 0; j < N; j++) {
#pragma unroll
        for (int a = 0; a < M; a++) tr[a] = miss[a] ? 0.0 : sw * (Ym[j].v[a] - mu[a]);
        rotate_in(tr);
      }
      // columns of Rchol = logChol2Chol(RLogChol) (R/modelTemplate.R:277-278), masked
      double Rc[M * M];
      Mdl::rchol(Rc, p);
#pragma unroll
      for (int b = 0; b < M; b++) {
        if constexpr (Mdl::R_DIAG) {
#pragma unroll
          for (int a = 0; a < M; a++) tr[a] = 0.0;
          tr[b] = miss[b] ? 1.0 : Rc[b + b * M];
        } else {
#pragma unroll
          for (int a = 0; a < M; a++) tr[a] = (miss[a] || miss[b]) ? ((a == b) ? 1.0 : 0.0) : Rc[a + b * M];
        }
        rotate_in(tr);
      }
    }
    // rank-1 up/down-date with column 0, weight |wc0| applied as v^(1/4) — reference quirk Q3, reproduced
    // literally (cholupdate_C, inst/include/psydiffInternals.h:133-165, 187-195)
    {
      const double dir = (ut.wc0 < 0.0) ? -1.0 : 1.0;
      const double v4 = sqrt(sqrt(fabs(ut.wc0)));
      double x[M];
#pragma unroll
      for (int a = 0; a < M; a++) x[a] = miss[a] ? 0.0 : v4 * (Y0.v[a] - mu[a]);
#pragma unroll
      for (int k = 0; k < M; k++) {
        const double lkk = S[tri(k, k)];
        const double r = sqrt(lkk * lkk + dir * x[k] * x[k]);
        const double il = 1.0 / lkk;
        const double c = r * il, s = x[k] * il;
        const double ic = 1.0 / c;
        S[tri(k, k)] = r;
#pragma unroll
        for (int i = k + 1; i < M; i++) {
          const double lik = (S[tri(i, k)] + dir * s * x[i]) * ic;
          S[tri(i, k)] = lik;
          x[i] = c * x[i] - s * lik;
        }
      }
    }
    double rS[M];
#pragma unroll
    for (int a = 0; a < M; a++) rS[a] = 1.0 / S[tri(a, a)];
    // w = srS⁻¹ (obs − mu): gives both the Mahalanobis term and, through Z, the mean update
    double w[M], dist = 0.0, ld = 0.0;
#pragma unroll
    for (int a = 0; a < M; a++) {
      double acc = miss[a] ? 0.0 : (ob[a] - mu[a]);
#pragma unroll
      for (int k = 0; k < a; k++) acc -= S[tri(a, k)] * w[k];
      w[a] = acc * rS[a];
      dist += w[a] * w[a];
      ld += miss[a] ? 0.0 : log(S[tri(a, a)]);
    }
    // C = X W Yᵀ = wc1 sqrt(c) Σ_j A_j (Yp_j − Ym_j)ᵀ (predictC_C; the column-0 term is (x_0 − X wm) = 0).
    // Z = C srS⁻ᵀ.  K = Z srS⁻¹ (computeK_SR_C) is never formed:  K (obs − mu) = Z w  and  U = K srS = Z.
    double Z[N][M];
#pragma unroll
    for (int i = 0; i < N; i++) {
      double crow[M];
#pragma unroll
      for (int a = 0; a < M; a++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j <= i; j++) acc += A[tri(i, j)] * (Yp[j].v[a] - Ym[j].v[a]);
        crow[a] = miss[a] ? 0.0 : ut.kap * acc;
      }
#pragma unroll
      for (int a = 0; a < M; a++) {
        double acc = crow[a];
#pragma unroll
        for (int k = 0; k < a; k++) acc -= Z[i][k] * S[tri(a, k)];
        Z[i][a] = acc * rS[a];
      }
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < M; a++) acc += Z[i][a] * w[a];
      m[i] += acc;  // updateM_C
    }
    // A ← cholupdate(A, U[:,b], 1, "downdate") for every manifest b (R/modelTemplate.R:382-386)
#pragma unroll
    for (int b = 0; b < M; b++) {
      if (miss[b]) continue;  // U[:,b] is exactly zero: the downdate would be an exact no-op
      double x[N];
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = Z[i][b];
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double lkk = A[tri(k, k)];
        const double r = sqrt(lkk * lkk - x[k] * x[k]);
        const double il = 1.0 / lkk;
        const double c = r * il, s = x[k] * il;
        const double ic = 1.0 / c;
        A[tri(k, k)] = r;
#pragma unroll
        for (int i = k + 1; i < N; i++) {
          const double lik = (A[tri(i, k)] - s * x[i]) * ic;
          A[tri(i, k)] = lik;
          x[i] = c * x[i] - s * lik;
        }
      }
    }
    // computeIndividualM2LLChol_C (inst/include/psydiffInternals.h:117-131)
    return (double)o * 1.8378770664093453 + 2.0 * ld + dist;
  }

  // ---- one instance through all of its person's time points ------------------------------------------------
  static __device__ __forceinline__ void run(const PsyLaunch& L) {
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= L.n_inst) return;
    const PsyInst in = L.inst[ii];
    const int person = in.person;
    // free-slot values of this instance: theta gathered through the person's slot map, one entry shifted by ±eps
    // (setParameterValues_C + setParameterList_C, inst/include/psydiffBasic.h:4-33, 63-121, stateless — quirk Q4)
    double p[F];
    {
      const int* ti = L.theta_index + (size_t)person * Mdl::NSLOT;
      const int pert = (in.code >= 0) ? (in.code >> 1) : -1;
      const double sh = (in.code & 1) ? L.eps : -L.eps;
#pragma unroll
      for (int k = 0; k < F; k++) p[k] = 0.0;
#pragma unroll
      for (int k = 0; k < Mdl::NSLOT; k++) {
        const int t = __ldg(ti + k);
        double v = __ldg(L.theta + t);
        if (t == pert) v += sh;
        p[k] = v;
      }
    }
    UT ut;
    ut.sqrtc = L.sqrtc; ut.wm0 = L.wm0; ut.wm1 = L.wm1; ut.wc0 = L.wc0; ut.wc1 = L.wc1;
    ut.kap = L.wc1 * L.sqrtc;
    const int pid = __ldg(L.person_id + person);

    double m[N], A[TN], Q[QLEN];
    {
      double A0[N * N], Ll[N * N];
      Mdl::initial(m, A0, Ll, p);  // m0, A0LogChol, LLogChol entries (R/modelTemplate.R:275-276, 81)
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) A[tri(i, j)] = (i == j) ? exp(A0[i + j * N]) : A0[i + j * N];  // logChol2Chol_C
      if constexpr (Mdl::L_STRUCT == L_DIAG) {
#pragma unroll
        for (int k = 0; k < N; k++) { const double l = exp(Ll[k + k * N]); Q[k] = l * l; }
      } else {
#pragma unroll
        for (int i = 0; i < N; i++) Ll[i + i * N] = exp(Ll[i + i * N]);
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
          for (int j = 0; j <= i; j++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < N; k++) acc += Ll[i + k * N] * Ll[j + k * N];
            Q[tri(i, j)] = acc;
          }
      }
    }
    const long long r0 = __ldg(L.row_off + person);
    const int T = (int)(__ldg(L.row_off + person + 1) - r0);
    const bool base = (in.code < 0);
    double f = 0.0, tabs = 0.0;
#pragma unroll 1
    for (int tp = 0; tp < T; tp++) {
      const long long row = r0 + tp;
      const double dtv = __ldg(L.dt + row);
      tabs += dtv;
      if (dtv != 0.0) rk4(p, m, A, Q, ut, pid, L.h, __ldg(L.ksteps + row));
      double ob[M];
#pragma unroll
      for (int a = 0; a < M; a++) ob[a] = __ldg(L.obs + row * M + a);
      double* prow = (base && L.pred) ? (L.pred + row * M) : nullptr;
      f += update_sr(p, m, A, ob, ut, pid, tabs, L.skip_update != 0, prow);
      if (base && L.latent) {
#pragma unroll
        for (int i = 0; i < N; i++) L.latent[row * N + i] = m[i];
      }
    }
    L.f_out[L.out_idx ? L.out_idx[ii] : ii] = f;
  }
};

}  // namespace psy

#define PSY_INSTANTIATE(MODEL)                                                                  \
  extern "C" __global__ void __launch_bounds__(PSY_BLOCK, PSY_MIN_BLOCKS) psy_filter_kernel(const PsyLaunch L) { \
    psy::Filter<MODEL>::run(L);                                                                 \
  }
