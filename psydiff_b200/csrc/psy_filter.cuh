// psy_filter.cuh — K1, the fused per-instance SR-UKF filter kernel (hand-written FP64 CUDA for sm_100a).
//
// One CUDA thread owns one filter instance = (person, finite-difference perturbation) and carries it through
// ALL of that person's time points: sigma points, the sigma-point matrix ODE (Särkkä 2007 eq. 34) under RK4 or adaptive
// Dormand-Prince, the square-root (or covariance) measurement update and the -2LL accumulation never leave registers.  This replaces, for the
// whole person x time loop, the reference's
//   fitModel                         R/modelTemplate.R:165-420  (SR variant)
//   odeintModel::operator()          R/modelTemplate.R:138-163
//   getMMatrix / computeL            R/modelTemplate.R:77-135
//   integrate_const<runge_kutta4>    call sites R/modelTemplate.R:327-329 (Boost.odeint); integrate() = dopri5 :331-333
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
// It is what lets one instance fit a thread's register file for n <= 6.  The RHS is evaluated in column-scaled form
// (rhs_scaled: A = Ã·diag(A_jj), unit-lower Ã — pure FMA chains) and skips the sigma-point columns that cannot move a
// drift component (Jacobian sparsity from the code generator, Mdl::dep).
//
// Missing manifests (NaN) are handled without dynamic indexing: a missing variable's deviations are zeroed and
// its row/column of Rchol replaced by the unit vector, which makes srS block-diagonal with a unit entry there,
// its Kalman-gain column exactly zero and its covariance downdate an exact no-op — the same numbers the
// reference gets from Y_.rows(nonmissing) / Rchol.submat(nonmissing, nonmissing) (R/modelTemplate.R:365-392).
#pragma once
#include "psy_launch.h"
#include "psy_lang.cuh"

#ifndef PSY_BLOCK
#define PSY_BLOCK 0   // 0 = by model size: 64 threads, 32 for N >= 7 (Filter::BLOCK)
#endif
#ifndef PSY_MIN_BLOCKS
#define PSY_MIN_BLOCKS 0  // 0 = by model (Filter::MIN_BLOCKS): 8 for the n <= 2 dopri5 kernel (128 registers), else 1
#endif
#ifndef PSY_RHS_SCALED
#define PSY_RHS_SCALED 1
#endif
#ifndef PSY_RHS_FUSED
#define PSY_RHS_FUSED 1   // only meaningful with PSY_RHS_SCALED: diffusion term folded into the triangular product (rhs_fused);
#endif                    // timed on B200 (profiles/r02c_variants.txt): C4 610.8 -> 604.8 ms, n = 8 734.4 -> 709.5 ms
#ifndef PSY_RHS_STREAM
#define PSY_RHS_STREAM 0  // small-live-set ordering of rhs_scaled (rhs_stream)
#endif
#ifndef PSY_CHAIN_COLUMNS
#define PSY_CHAIN_COLUMNS 1
#endif

namespace psy {

__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower, i >= j

enum { L_DIAG = 0, L_FULL = 1 };

// Reciprocal without the IEEE-division slow path: hardware seed (rcp.approx.ftz.f64, >= 20 mantissa bits) + two
// Newton steps -> error <= ~1 ulp, no branches.  0 -> inf -> NaN downstream, like the reference's 1/0.
// (-DPSY_IEEE_DIV restores 1.0 / x.)
__device__ __forceinline__ double rcp(double x) {
#ifdef PSY_IEEE_DIV
  return 1.0 / x;
#else
  double r;
#ifdef __CUDA_ARCH__
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
  r = psy_sim_rcp_seed(x);  // host compile of this header = tests/hostsim (kernel-source emulation in the CPU test suite)
#endif
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#endif
}

// Rotations from 1/sqrt (PSY_RSQRT_ROT: 1 = on, 0 = off, default -1 = on for N <= 6; timed on B200, profiles/r02a_experiments.txt:
// C4 618.5 -> 610.7 ms, n = 8 806.8 -> 827.1 ms because the update's stack grows by 576 B there).
// 1/sqrt(x) from the hardware seed (rsqrt.approx.ftz.f64) + two Newton steps, branch-free like rcp().  The Givens rotations of
// the measurement update and the cholupdate steps then need no IEEE sqrt (≈ 16 instructions with a slow-path branch each) and
// no reciprocal of the root: r = r²·r⁻¹, c = a·r⁻¹, s = b·r⁻¹.  sqrt(negative) = NaN either way; a ZERO radicand gives NaN here
// and 0 (then inf) with sqrt — both non-finite outcomes of a failed downdate.
__device__ __forceinline__ double rsq(double x) {
  double y;
#ifdef __CUDA_ARCH__
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  y = psy_sim_rsqrt_seed(x);
#endif
  const double h = 0.5 * x;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

#ifndef PSY_RSQRT_ROT
#define PSY_RSQRT_ROT (-1)
#endif
#ifndef PSY_DOPRI_LEAN
#define PSY_DOPRI_LEAN 1   // 1 (default): step-size controller of the adaptive stepper without libdevice pow and IEEE division (below);
                           // C3 on B200: 112.0 -> 90.9 ms (profiles/r02n_variants.txt).  0 = pow() and '/' as odeint writes them
#endif

template <class Mdl>
struct Filter {
  static constexpr int N = Mdl::N, M = Mdl::M, F = Mdl::NFREE;  // NFREE >= 1 (padded); NSLOT real slots
  static constexpr int TN = N * (N + 1) / 2, TM = M * (M + 1) / 2;
  static constexpr int QLEN = (Mdl::L_STRUCT == L_DIAG) ? N : TN;
  // Threads per block.  N >= 7: the state no longer fits the register file (1.1 KB spill frame at n = 8), and what bounds the kernel is
  // where the spills live: measured on B200 (profiles/r02a_experiments.txt, n = 8) 8 resident warps per SM thrash the L1 that is left
  // beside the RK4 parking area (L1 hit rate of spill loads 50 %, spill stores written through to L2 at 3.7 TB/s: 807 ms), 4 warps
  // with a 45 % shared-memory carveout keep the frame L1-resident (776 ms), and one-warp blocks schedule best (736 ms).
  static constexpr bool RSQ_ROT = (PSY_RSQRT_ROT == 1) || (PSY_RSQRT_ROT == -1 && Mdl::N <= 6);
  // (n = 10: 64-thread blocks again, 3324 -> 3063 ms, profiles/r02c_variants.txt)
  static constexpr int BLOCK = (PSY_BLOCK > 0) ? PSY_BLOCK : ((N >= 7 && N <= 8) ? 32 : 64);
  // Resident blocks the register allocation is capped for.  The n <= 2 adaptive kernel (C3) is latency bound (divergent step-size
  // control): 128 registers / 16 warps per SM instead of 204 / 8 measured 139.2 -> 112.9 ms on B200 (profiles/r02c_variants.txt).
  static constexpr int MIN_BLOCKS = (PSY_MIN_BLOCKS > 0) ? PSY_MIN_BLOCKS : ((Mdl::STEPPER == 1 && N <= 2) ? 512 / BLOCK : 1);
  static constexpr int SMEM_CARVEOUT_PCT = (N >= 7) ? 45 : -1;   // preferred shared-memory carveout (percent of the maximum), -1 = driver's choice

  struct UT { double sqrtc, wm0, wm1, wc0, wc1, kap; };

  // ---- RHS of the reduced sigma-point ODE (replaces odeintModel::operator() + getMMatrix) -------------------
  static __device__ __forceinline__ void rhs_plain(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                             double (&dm)[N], double (&dA)[TN], const double (&Q)[QLEN],
                                             const UT& ut, int person, double t) {
    // explicit inverse of the lower-triangular root (arma::inv(arma::trimatl(A)), R/modelTemplate.R:97)
    double rd[N], Ai[TN];
#pragma unroll
    for (int i = 0; i < N; i++) rd[i] = rcp(A[tri(i, i)]);
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

  // Column-scaled form of the same RHS (default, -DPSY_RHS_SCALED=0 selects rhs_plain).  With A = Ã·D, D = diag(A_jj), Ã unit
  // lower triangular and W = Ã⁻¹ (unit lower):  A⁻¹ = D⁻¹W,  G = D⁻¹G̃ with G̃ = W·[d_1..d_n],
  //   M = D⁻¹ M̃ D⁻¹,  M̃_ij = kap (G̃_ij A_jj + A_ii G̃_ji) + (W LLᵀ Wᵀ)_ij,   Phi(M) = D⁻¹ Phi(M̃) D⁻¹,   dA = A Phi(M) = (Ã Phi(M̃)) D⁻¹.
  // Every unit diagonal turns the first term of a dot product into the accumulator's seed, so the triangular inverse,
  // G̃, the diffusion term and Ã·Phi are pure FMA chains: 72 fewer FP64 instructions per evaluation at n = 6 for the same
  // numbers up to rounding.
  // SCALE = false leaves out the final D⁻¹ (dA_ij = raw_ij · rd_j) and hands rd back: the RK4 driver folds it into its
  // step coefficients (2n multiplies instead of n(n+1)/2).
  template <bool SCALE>
  static __device__ __forceinline__ void rhs_scaled(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                                    double (&dm)[N], double (&dA)[TN], double (&rd)[N], const double (&Q)[QLEN],
                                                    const UT& ut, int person, double t) {
    double At[TN], W[TN];   // At / W: strictly-lower entries are used, the unit diagonal is implied
#pragma unroll
    for (int i = 0; i < N; i++) rd[i] = rcp(A[tri(i, i)]);
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
      for (int i = j + 1; i < N; i++) At[tri(i, j)] = A[tri(i, j)] * rd[j];
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
      for (int i = j + 1; i < N; i++) {
        double acc = -At[tri(i, j)];
#pragma unroll
        for (int k = j + 1; k < i; k++) acc = fma(-At[tri(i, k)], W[tri(k, j)], acc);
        W[tri(i, j)] = acc;
      }
    // drift at the 2n+1 sigma points, one evaluation per column; G̃[:,j] = W d_j.
    // Column j moves x_j..x_{n-1} only (A is lower triangular).  A drift component that reads none of them
    // (Mdl::dep, the Jacobian sparsity the code generator found) has f(x+) = f(x-) = f(x0) bit for bit: its difference
    // is exactly 0 and its sum exactly 2 f0, so neither is evaluated and the zero never enters an FMA chain.
    Vec<N> xc, f0;
#pragma unroll
    for (int i = 0; i < N; i++) { xc.v[i] = m[i]; f0.v[i] = 0.0; }
    Mdl::latentEquation(xc, f0, p, person, t);
    double G[N * N];
    // weighted mean drift, accumulated by FMA: the columns that cannot move component i contribute 2 wm1 f0_i each
#pragma unroll
    for (int i = 0; i < N; i++) {
      int unmoved = 0;
#pragma unroll
      for (int j = 0; j < N; j++) unmoved += ((Mdl::dep(i) >> j) != 0u) ? 0 : 1;
      dm[i] = (ut.wm0 + 2.0 * (double)unmoved * ut.wm1) * f0.v[i];
    }
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
        if ((Mdl::dep(i) >> j) != 0u) {
          d[i] = fp.v[i] - fm.v[i];
          dm[i] = fma(ut.wm1, fp.v[i] + fm.v[i], dm[i]);   // one add off the critical path, one FMA on the accumulation chain
        } else {
          d[i] = 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        double acc = d[i];
#pragma unroll
        for (int k = 0; k < i; k++)
          if ((Mdl::dep(k) >> j) != 0u) acc = fma(W[tri(i, k)], d[k], acc);
        G[i + j * N] = acc;
      }
    }

    // Phi(M̃): lower triangle with the halved diagonal
    double Ph[TN], ka[N];
#pragma unroll
    for (int i = 0; i < N; i++) ka[i] = ut.kap * A[tri(i, i)];
    if constexpr (Mdl::L_STRUCT == L_DIAG) {
      // W diag(q) Wᵀ:  Bq = W diag(q) (unit diagonal -> Bq_kk = q_k)
      double Bq[TN];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int k = 0; k < i; k++) Bq[tri(i, k)] = W[tri(i, k)] * Q[k];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          double acc = (i == j) ? Q[j] : Bq[tri(i, j)];
#pragma unroll
          for (int k = 0; k < j; k++) acc = fma(Bq[tri(i, k)], W[tri(j, k)], acc);
          if (i == j) acc = fma(G[i + i * N], ka[i], 0.5 * acc);
          else acc = fma(G[i + j * N], ka[j], fma(G[j + i * N], ka[i], acc));
          Ph[tri(i, j)] = acc;
        }
    } else {
      // general L: Z = W Q (Q = L Lᵀ symmetric, packed lower), then lower(Z Wᵀ)
      double Z[N * N];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
          double acc = (i >= j) ? Q[tri(i, j)] : Q[tri(j, i)];
#pragma unroll
          for (int k = 0; k < i; k++) acc = fma(W[tri(i, k)], (k >= j) ? Q[tri(k, j)] : Q[tri(j, k)], acc);
          Z[i + j * N] = acc;
        }
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          double acc = Z[i + j * N];
#pragma unroll
          for (int k = 0; k < j; k++) acc = fma(Z[i + k * N], W[tri(j, k)], acc);
          if (i == j) acc = fma(G[i + i * N], ka[i], 0.5 * acc);
          else acc = fma(G[i + j * N], ka[j], fma(G[j + i * N], ka[i], acc));
          Ph[tri(i, j)] = acc;
        }
    }
    // dA = (Ã Phi(M̃)) D⁻¹
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double acc = Ph[tri(i, j)];
#pragma unroll
        for (int k = j; k < i; k++) acc = fma(At[tri(i, k)], Ph[tri(k, j)], acc);
        dA[tri(i, j)] = SCALE ? acc * rd[j] : acc;
      }
  }

  // Fused form of the same RHS (default; -DPSY_RHS_FUSED=0 selects rhs_scaled).  The diffusion term is symmetric, so half of it
  // can ride along with the drift differences through the ONE triangular product the RHS needs anyway:
  //   D'_j = kap A_jj d_j + ½ (Q Wᵀ)_j,   H = W·[D'_1..D'_n] = G̃·diag(kap A_jj) + ½ W Q Wᵀ,   M̃ = H + Hᵀ,
  //   Phi(M̃)_ij = H_ij + H_ji (i > j),  Phi(M̃)_jj = H_jj.
  // For diagonal Q, (Q Wᵀ)_j has entries q_k W_jk for k <= j only, so the separate W diag(q) Wᵀ product of rhs_scaled
  // (n(n-1)/2 multiplies + n(n-1)(n+1)/6 FMAs: 50 FP64 instructions at n = 6, 112 at n = 8) disappears, and Phi is assembled with one
  // add per entry instead of two FMAs.  QH holds ½Q (run() halves it once per instance).  Same numbers up to rounding.
  template <bool SCALE>
  static __device__ __forceinline__ void rhs_fused(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                                   double (&dm)[N], double (&dA)[TN], double (&rd)[N], const double (&QH)[QLEN],
                                                   const UT& ut, int person, double t) {
    double At[TN], W[TN];   // strictly-lower entries are used, the unit diagonal is implied
#pragma unroll
    for (int i = 0; i < N; i++) rd[i] = rcp(A[tri(i, i)]);
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
      for (int i = j + 1; i < N; i++) At[tri(i, j)] = A[tri(i, j)] * rd[j];
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
      for (int i = j + 1; i < N; i++) {
        double acc = -At[tri(i, j)];
#pragma unroll
        for (int k = j + 1; k < i; k++) acc = fma(-At[tri(i, k)], W[tri(k, j)], acc);
        W[tri(i, j)] = acc;
      }
    Vec<N> xc, f0;
#pragma unroll
    for (int i = 0; i < N; i++) { xc.v[i] = m[i]; f0.v[i] = 0.0; }
    Mdl::latentEquation(xc, f0, p, person, t);
#pragma unroll
    for (int i = 0; i < N; i++) {
      int unmoved = 0;
#pragma unroll
      for (int j = 0; j < N; j++) unmoved += ((Mdl::dep(i) >> j) != 0u) ? 0 : 1;
      dm[i] = (ut.wm0 + 2.0 * (double)unmoved * ut.wm1) * f0.v[i];
    }
    double H[N * N];
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
      const double kaj = ut.kap * A[tri(j, j)];
      // D'_j; entry k is structurally zero when k > j (no diffusion part) and column j cannot move drift component k
      double d[N];
      if constexpr (Mdl::L_STRUCT == L_DIAG) {
#pragma unroll
        for (int k = 0; k < N; k++) {
          const bool moved = (Mdl::dep(k) >> j) != 0u;
          if (moved) dm[k] = fma(ut.wm1, fp.v[k] + fm.v[k], dm[k]);
          if (k < j) {
            const double hq = QH[k] * W[tri(j, k)];
            d[k] = moved ? fma(kaj, fp.v[k] - fm.v[k], hq) : hq;
          } else if (k == j) {
            d[k] = moved ? fma(kaj, fp.v[k] - fm.v[k], QH[k]) : QH[k];
          } else {
            d[k] = moved ? kaj * (fp.v[k] - fm.v[k]) : 0.0;
          }
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
          double acc = d[i];
#pragma unroll
          for (int k = 0; k < i; k++)
            if (k <= j || (Mdl::dep(k) >> j) != 0u) acc = fma(W[tri(i, k)], d[k], acc);
          H[i + j * N] = acc;
        }
      } else {
        // general L: ½ (Q Wᵀ)_j is a full column, QH = ½ L Lᵀ packed lower
#pragma unroll
        for (int k = 0; k < N; k++) {
          const bool moved = (Mdl::dep(k) >> j) != 0u;
          if (moved) dm[k] = fma(ut.wm1, fp.v[k] + fm.v[k], dm[k]);
          double acc = (k >= j) ? QH[tri(k, j)] : QH[tri(j, k)];
#pragma unroll
          for (int l = 0; l < j; l++) acc = fma((k >= l) ? QH[tri(k, l)] : QH[tri(l, k)], W[tri(j, l)], acc);
          d[k] = moved ? fma(kaj, fp.v[k] - fm.v[k], acc) : acc;
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
          double acc = d[i];
#pragma unroll
          for (int k = 0; k < i; k++) acc = fma(W[tri(i, k)], d[k], acc);
          H[i + j * N] = acc;
        }
      }
    }
    // Phi(M̃) and dA = (Ã Phi(M̃)) D⁻¹
    double Ph[TN];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) Ph[tri(i, j)] = (i == j) ? H[i + i * N] : H[i + j * N] + H[j + i * N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double acc = Ph[tri(i, j)];
#pragma unroll
        for (int k = j; k < i; k++) acc = fma(At[tri(i, k)], Ph[tri(k, j)], acc);
        dA[tri(i, j)] = SCALE ? acc * rd[j] : acc;
      }
  }

  // Small-live-set form of rhs_scaled (-DPSY_RHS_STREAM=1; default for N >= 7, where the working set of rhs_scaled does not
  // fit the register file and its spill STORES — write-through to L2 — set the pace, profiles/r02a_*).  Same arithmetic up to
  // rounding, ordered so that the big temporaries never exist:
  //   * Phi(M̃) starts as the diffusion part W diag(q) Wᵀ (needs W only) BEFORE the columns are visited;
  //   * every column's G̃_j = W d_j is folded into Phi at once (Phi_max(i,j),min(i,j) += kap A_jj G̃_ij): G (n² doubles) is never stored;
  //   * dA = (Ã Phi) D⁻¹ is evaluated as A·(D⁻¹ Phi), accumulating row by row of Phi: Ã (n(n-1)/2 doubles) is dead once W exists.
  template <bool SCALE>
  static __device__ __forceinline__ void rhs_stream(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                                    double (&dm)[N], double (&dA)[TN], double (&rd)[N], const double (&Q)[QLEN],
                                                    const UT& ut, int person, double t) {
    static_assert(Mdl::L_STRUCT == L_DIAG, "rhs_stream is written for diagonal L");
    double W[TN], Ph[TN];
    {
      double At[TN];
#pragma unroll
      for (int i = 0; i < N; i++) rd[i] = rcp(A[tri(i, i)]);
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = j + 1; i < N; i++) At[tri(i, j)] = A[tri(i, j)] * rd[j];
#pragma unroll
      for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = j + 1; i < N; i++) {
          double acc = -At[tri(i, j)];
#pragma unroll
          for (int k = j + 1; k < i; k++) acc = fma(-At[tri(i, k)], W[tri(k, j)], acc);
          W[tri(i, j)] = acc;
        }
    }
    // Phi <- W diag(q) Wᵀ (lower, diagonal halved)
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double acc = (i == j) ? Q[j] : W[tri(i, j)] * Q[j];
#pragma unroll
        for (int k = 0; k < j; k++) acc = fma(W[tri(i, k)] * Q[k], W[tri(j, k)], acc);
        Ph[tri(i, j)] = (i == j) ? 0.5 * acc : acc;
      }
    Vec<N> xc, f0;
#pragma unroll
    for (int i = 0; i < N; i++) { xc.v[i] = m[i]; f0.v[i] = 0.0; }
    Mdl::latentEquation(xc, f0, p, person, t);
#pragma unroll
    for (int i = 0; i < N; i++) {
      int unmoved = 0;
#pragma unroll
      for (int j = 0; j < N; j++) unmoved += ((Mdl::dep(i) >> j) != 0u) ? 0 : 1;
      dm[i] = (ut.wm0 + 2.0 * (double)unmoved * ut.wm1) * f0.v[i];
    }
    // PSY_CHAIN_COLUMNS: column j + 1's sigma points are made to depend on the last value column j produces (sq + 0·g: an
    // IEEE-exact no-op that neither nvvm nor ptxas may fold, because g could be NaN — in which case the instance is failing
    // anyway).  Without it ptxas hoists all 2n drift evaluations to the top of the unrolled body for ILP and every d_j stays live.
    double sq = ut.sqrtc;
#pragma unroll
    for (int j = 0; j < N; j++) {
      Vec<N> xp, xm, fp, fm;
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i >= j) {
          xp.v[i] = fma(sq, A[tri(i, j)], m[i]);
          xm.v[i] = fma(-sq, A[tri(i, j)], m[i]);
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
        if ((Mdl::dep(i) >> j) != 0u) {
          d[i] = fp.v[i] - fm.v[i];
          dm[i] = fma(ut.wm1, fp.v[i] + fm.v[i], dm[i]);
        } else {
          d[i] = 0.0;
        }
      }
      const double kaj = ut.kap * A[tri(j, j)];
#pragma unroll
      for (int i = 0; i < N; i++) {
        bool any = (Mdl::dep(i) >> j) != 0u;
        double acc = d[i];
#pragma unroll
        for (int k = 0; k < i; k++)
          if ((Mdl::dep(k) >> j) != 0u) { acc = fma(W[tri(i, k)], d[k], acc); any = true; }
        if (any) {   // G̃_ij enters Phi(M̃) at (max, min); twice on the diagonal, whose half is what Phi keeps
          if (i >= j) Ph[tri(i, j)] = fma(acc, kaj, Ph[tri(i, j)]);
          else Ph[tri(j, i)] = fma(acc, kaj, Ph[tri(j, i)]);
        }
#if PSY_CHAIN_COLUMNS
        if (i == N - 1 && j + 1 < N) sq = fma(acc, 0.0, sq);
#endif
      }
    }
    // dA = A (D⁻¹ Phi) [D⁻¹]: row k of Phi seeds row k of dA and, scaled by 1/A_kk, feeds the rows below it
#pragma unroll
    for (int i = 0; i < TN; i++) dA[i] = Ph[i];
#pragma unroll
    for (int k = 0; k < N - 1; k++)
#pragma unroll
      for (int j = 0; j <= k; j++) {
        const double ps = Ph[tri(k, j)] * rd[k];
#pragma unroll
        for (int i = k + 1; i < N; i++) dA[tri(i, j)] = fma(A[tri(i, k)], ps, dA[tri(i, j)]);
      }
    if (SCALE) {
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) dA[tri(i, j)] *= rd[j];
    }
  }

  // the column-scaled RHS in use: fused (Q holds ½ L Lᵀ) or plain scaled (Q holds L Lᵀ)
  static constexpr double Q_SCALE = (PSY_RHS_SCALED && PSY_RHS_FUSED && PSY_RHS_STREAM != 1) ? 0.5 : 1.0;
  template <bool SCALE>
  static __device__ __forceinline__ void rhs_col(const double (&p)[F], const double (&m)[N], const double (&A)[TN], double (&dm)[N],
                                                 double (&dA)[TN], double (&rd)[N], const double (&Q)[QLEN], const UT& ut, int person, double t) {
#if PSY_RHS_STREAM == 1
    rhs_stream<SCALE>(p, m, A, dm, dA, rd, Q, ut, person, t);
#elif PSY_RHS_FUSED
    rhs_fused<SCALE>(p, m, A, dm, dA, rd, Q, ut, person, t);
#else
    rhs_scaled<SCALE>(p, m, A, dm, dA, rd, Q, ut, person, t);
#endif
  }

  static __device__ __forceinline__ void rhs(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                             double (&dm)[N], double (&dA)[TN], const double (&Q)[QLEN],
                                             const UT& ut, int person, double t) {
#if PSY_RHS_SCALED
    double rd[N];
    rhs_col<true>(p, m, A, dm, dA, rd, Q, ut, person, t);
#else
    rhs_plain(p, m, A, dm, dA, Q, ut, person, t);
#endif
  }

  // ---- classical RK4, K fixed steps (Boost.odeint integrate_const<runge_kutta4>, rule T1) -------------------
  // x_{n+1} = x_n + h/6 k1 + h/3 k2 + h/3 k3 + h/6 k4.
  // For N >= 7 the step's start state x_n and the running sum are parked in shared memory ([element][thread], conflict
  // free; loads batched before the arithmetic, stores after) so that only the stage input lives in registers while the
  // RHS runs.  Measured on B200 (profiles/r01_variants.txt): n = 8: 1167 -> 1096 ms; n = 6: 722 -> 762 ms, i.e. the
  // compiler's own spill placement wins as long as the state nearly fits, hence the threshold.  -DPSY_RK4_SMEM=0/1 overrides.
#ifndef PSY_RK4_SMEM
#define PSY_RK4_SMEM (-1)
#endif
#ifndef PSY_STEP_UNROLL
#define PSY_STEP_UNROLL 1
#endif
#ifndef PSY_STAGE_UNROLL
#define PSY_STAGE_UNROLL 0   // RK4 stages per loop body (1, 2 or 4; 0 = by model size): > 1 lets ptxas schedule across stage boundaries
#endif                       // and drop the dead stage-input update of the 4th stage, at 2x / 4x the hot loop's code size.  Measured on
                             // B200 (profiles/r02m_stage_unroll.txt, r02n_variants.txt): n = 2 (C2) 139.4 -> 127.6 ms with 4 stages per body (311 instructions), n = 4 306.2 -> 290.2 ms;
                             // n = 6 605 -> 609 (2) / 723 ms (4) and the n = 8 lane-pair kernel 1296 -> 1547 (2) / 2180 ms (4): a stage loop of more than
                             // 32 KB of SASS (2000 instructions) no longer fits the L1.5 instruction cache that 8 warps at different
                             // positions of the loop share
  static constexpr int STAGE_UNROLL = (PSY_STAGE_UNROLL > 0) ? PSY_STAGE_UNROLL : (N <= 4 ? 4 : 1);
  static constexpr int STEP_UNROLL = PSY_STEP_UNROLL;   // RK4 steps per loop body of the register-resident path (experiment knob)
  static constexpr bool RK4_SMEM = (PSY_RK4_SMEM == 1) || (PSY_RK4_SMEM == -1 && N >= 7);
  static constexpr int SMEM_BYTES = RK4_SMEM ? 2 * (N + TN) * BLOCK * 8 : 0;
  static __device__ __forceinline__ void rk4(const double (&p)[F], double (&m)[N], double (&A)[TN],
                                             const double (&Q)[QLEN], const UT& ut, int person, double h, int K) {
    if constexpr (RK4_SMEM) {
      extern __shared__ double psy_smem[];
      // volatile: the values must really be re-read from shared memory, not forwarded through registers
      volatile double* sx = psy_smem + threadIdx.x;                       // x_n : sx[e * BLOCK]
      volatile double* sa = psy_smem + (N + TN) * BLOCK + threadIdx.x;  // sum : sa[e * BLOCK]
#pragma unroll 1
      for (int step = 0; step < K; step++) {
        const double t0 = (double)step * h;
#pragma unroll
        for (int i = 0; i < N; i++) { sx[i * BLOCK] = m[i]; sa[i * BLOCK] = m[i]; }
#pragma unroll
        for (int i = 0; i < TN; i++) { sx[(N + i) * BLOCK] = A[i]; sa[(N + i) * BLOCK] = A[i]; }
#pragma unroll 1
        for (int sp = 0; sp < 4 / STAGE_UNROLL; sp++)
#pragma unroll
        for (int su = 0; su < STAGE_UNROLL; su++) {
          const int st = sp * STAGE_UNROLL + su;
          double km[N], kA[TN];
          const double tc = (st == 0) ? t0 : ((st == 3) ? t0 + h : t0 + 0.5 * h);
          const double b = (st == 0 || st == 3) ? h * (1.0 / 6.0) : h * (1.0 / 3.0);
          double rdv[N];
#if PSY_RHS_SCALED
          rhs_col<false>(p, m, A, km, kA, rdv, Q, ut, person, tc);   // D⁻¹ is folded into the step coefficients below
#else
          rhs(p, m, A, km, kA, Q, ut, person, tc);
#pragma unroll
          for (int j = 0; j < N; j++) rdv[j] = 1.0;
#endif
          // phased access: all loads first (they pipeline), then the arithmetic, then all stores
          double ta[N + TN], tx[N + TN];
#pragma unroll
          for (int i = 0; i < N + TN; i++) ta[i] = sa[i * BLOCK];
          if (st < 3) {
#pragma unroll
            for (int i = 0; i < N + TN; i++) tx[i] = sx[i * BLOCK];
            const double a = (st == 2) ? h : 0.5 * h;
#pragma unroll
            for (int i = 0; i < N; i++) { ta[i] = fma(b, km[i], ta[i]); m[i] = fma(a, km[i], tx[i]); }
#pragma unroll
            for (int j = 0; j < N; j++) {
              const double bj = b * rdv[j], aj = a * rdv[j];
#pragma unroll
              for (int i = j; i < N; i++) {
                ta[N + tri(i, j)] = fma(bj, kA[tri(i, j)], ta[N + tri(i, j)]);
                A[tri(i, j)] = fma(aj, kA[tri(i, j)], tx[N + tri(i, j)]);
              }
            }
#pragma unroll
            for (int i = 0; i < N + TN; i++) sa[i * BLOCK] = ta[i];
          } else {
#pragma unroll
            for (int i = 0; i < N; i++) m[i] = fma(b, km[i], ta[i]);
#pragma unroll
            for (int j = 0; j < N; j++) {
              const double bj = b * rdv[j];
#pragma unroll
              for (int i = j; i < N; i++) A[tri(i, j)] = fma(bj, kA[tri(i, j)], ta[N + tri(i, j)]);
            }
          }
        }
      }
    } else {
#pragma unroll STEP_UNROLL
      for (int step = 0; step < K; step++) {
        const double t0 = (double)step * h;
        double mt[N], At[TN], ma[N], Aa[TN];
#pragma unroll
        for (int i = 0; i < N; i++) { mt[i] = m[i]; ma[i] = m[i]; }
#pragma unroll
        for (int i = 0; i < TN; i++) { At[i] = A[i]; Aa[i] = A[i]; }
#pragma unroll 1
        for (int sp = 0; sp < 4 / STAGE_UNROLL; sp++)
#pragma unroll
        for (int su = 0; su < STAGE_UNROLL; su++) {
          const int st = sp * STAGE_UNROLL + su;
          double km[N], kA[TN];
          const double tc = (st == 0) ? t0 : ((st == 3) ? t0 + h : t0 + 0.5 * h);
          const double b = (st == 0 || st == 3) ? h * (1.0 / 6.0) : h * (1.0 / 3.0);
          const double a = (st == 2) ? h : 0.5 * h;
#if PSY_RHS_SCALED
          // column scaling D⁻¹ of dA folded into the step coefficients
          double rdv[N];
          rhs_col<false>(p, mt, At, km, kA, rdv, Q, ut, person, tc);
#pragma unroll
          for (int i = 0; i < N; i++) { ma[i] += b * km[i]; mt[i] = m[i] + a * km[i]; }
#pragma unroll
          for (int j = 0; j < N; j++) {
            const double bj = b * rdv[j], aj = a * rdv[j];
#pragma unroll
            for (int i = j; i < N; i++) {
              Aa[tri(i, j)] = fma(bj, kA[tri(i, j)], Aa[tri(i, j)]);
              At[tri(i, j)] = fma(aj, kA[tri(i, j)], A[tri(i, j)]);
            }
          }
#else
          rhs(p, mt, At, km, kA, Q, ut, person, tc);
#pragma unroll
          for (int i = 0; i < N; i++) { ma[i] += b * km[i]; mt[i] = m[i] + a * km[i]; }
#pragma unroll
          for (int i = 0; i < TN; i++) { Aa[i] += b * kA[i]; At[i] = A[i] + a * kA[i]; }
#endif
        }
#pragma unroll
        for (int i = 0; i < N; i++) m[i] = ma[i];
#pragma unroll
        for (int i = 0; i < TN; i++) A[i] = Aa[i];
      }
    }
  }

  // ---- adaptive Dormand-Prince 5(4) (Boost.odeint integrate() = integrate_adaptive(controlled_runge_kutta<
  // runge_kutta_dopri5>, eps_abs = eps_rel = 1e-6, a_x = a_dxdt = 1), call sites R/modelTemplate.R:331-333; rule T2 of
  // SURVEY §8c).  The controller's error norm is the max over the entries of the FULL sigma-point matrix, so it is
  // evaluated on the implied columns m, m ± sqrt(c) A_j of the reduced state.  A fresh stepper per observation interval
  // (FSAL derivative recomputed at t = 0).  On step-size underflow (odeint throws) or a NaN error the state is NaN.
  static constexpr int NS = N + TN;
  static __device__ __forceinline__ void rhs_z(const double (&p)[F], const double (&z)[NS], double (&k)[NS],
                                               const double (&Q)[QLEN], const UT& ut, int person, double t) {
    rhs(p, reinterpret_cast<const double(&)[N]>(z[0]), reinterpret_cast<const double(&)[TN]>(z[N]),
        reinterpret_cast<double(&)[N]>(k[0]), reinterpret_cast<double(&)[TN]>(k[N]), Q, ut, person, t);
  }
  static __device__ __forceinline__ double err_entry(double xe, double x, double dx, double dt) {
#if PSY_DOPRI_LEAN
    return fabs(xe) * rcp(1e-6 + 1e-6 * (fabs(x) + fabs(dt) * fabs(dx)));
#else
    return fabs(xe) / (1e-6 + 1e-6 * (fabs(x) + fabs(dt) * fabs(dx)));
#endif
  }
  // x^(-1/3) and x^(-1/5) of the step-size controller.  PSY_DOPRI_LEAN: single-precision seed (lg2 / ex2 on the special-function unit,
  // relative error <= 1e-6) and two Newton steps on y^-k = x (quadratic: 1e-6 -> 1e-12 -> 1e-16 before rounding) instead of libdevice's
  // pow (two call sites of ~190 instructions each).  The controller only scales dt with these, accept / reject is decided by err itself.
  template <int KPOW>
  static __device__ __forceinline__ double inv_root(double x) {
#if PSY_DOPRI_LEAN && defined(__CUDA_ARCH__)
    double y = (double)exp2f(log2f((float)x) * (-1.0f / (float)KPOW));
#pragma unroll
    for (int it = 0; it < 2; it++) {
      double yk = y * y;                               // y^2
      if constexpr (KPOW == 3) yk = yk * y;            // y^3
      else yk = yk * yk * y;                           // y^5
      y = y * fma(-x, yk, (double)(KPOW + 1)) * (1.0 / (double)KPOW);
    }
    return y;
#else
    return pow(x, -1.0 / (double)KPOW);
#endif
  }
  static __device__ void dopri5(const double (&p)[F], double (&m)[N], double (&A)[TN], const double (&Q)[QLEN],
                                             const UT& ut, int person, double t1, double dt) {
    constexpr double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9,
                     a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729,
                     a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
                     c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
    constexpr double e1 = c1 - 5179.0 / 57600, e3 = c3 - 7571.0 / 16695, e4 = c4 - 393.0 / 640,
                     e5 = c5 + 92097.0 / 339200, e6 = c6 - 187.0 / 2100, e7 = -1.0 / 40;
    constexpr double DEPS = 2.220446049250313e-16;
    double z[NS], k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS], zt[NS], zn[NS];
#pragma unroll
    for (int i = 0; i < N; i++) z[i] = m[i];
#pragma unroll
    for (int i = 0; i < TN; i++) z[N + i] = A[i];
    double t = 0.0;
    bool dead = false;
    rhs_z(p, z, k1, Q, ut, person, t);
#pragma unroll 1
    while (t1 - t > DEPS && !dead) {
      if ((t + dt) - t1 > DEPS) dt = t1 - t;
      int tries = 0;
#pragma unroll 1
      for (;;) {
        if (++tries > 500) { dead = true; break; }
#pragma unroll
        for (int i = 0; i < NS; i++) zt[i] = z[i] + dt * (a21 * k1[i]);
        rhs_z(p, zt, k2, Q, ut, person, t + dt * (1.0 / 5));
#pragma unroll
        for (int i = 0; i < NS; i++) zt[i] = z[i] + dt * (a31 * k1[i] + a32 * k2[i]);
        rhs_z(p, zt, k3, Q, ut, person, t + dt * (3.0 / 10));
#pragma unroll
        for (int i = 0; i < NS; i++) zt[i] = z[i] + dt * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
        rhs_z(p, zt, k4, Q, ut, person, t + dt * (4.0 / 5));
#pragma unroll
        for (int i = 0; i < NS; i++) zt[i] = z[i] + dt * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
        rhs_z(p, zt, k5, Q, ut, person, t + dt * (8.0 / 9));
#pragma unroll
        for (int i = 0; i < NS; i++) zt[i] = z[i] + dt * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
        rhs_z(p, zt, k6, Q, ut, person, t + dt);
#pragma unroll
        for (int i = 0; i < NS; i++) zn[i] = z[i] + dt * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
        rhs_z(p, zn, k7, Q, ut, person, t + dt);
        // error estimate, entry by entry over the implied sigma-point matrix (old state and old derivative)
        double xe[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) xe[i] = dt * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k7[i]);
        double err = 0.0;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; i++) {
          const double e = err_entry(xe[i], z[i], k1[i], dt);
          bad = bad || (e != e);
          err = fmax(err, e);
#pragma unroll
          for (int j = 0; j <= i; j++) {
            const int q = N + tri(i, j);
            const double ep = err_entry(xe[i] + ut.sqrtc * xe[q], z[i] + ut.sqrtc * z[q], k1[i] + ut.sqrtc * k1[q], dt);
            const double em = err_entry(xe[i] - ut.sqrtc * xe[q], z[i] - ut.sqrtc * z[q], k1[i] - ut.sqrtc * k1[q], dt);
            bad = bad || (ep != ep) || (em != em);
            err = fmax(err, fmax(ep, em));
          }
        }
        if (bad) { dead = true; break; }
        if (err > 1.0) {
          dt *= fmax(0.9 * inv_root<3>(err), 0.2);
          continue;
        }
        t += dt;
        if (err < 0.5) dt *= 0.9 * inv_root<5>(fmax(3.2e-4, err));  // 5^-5 = 3.2e-4
#pragma unroll
        for (int i = 0; i < NS; i++) { z[i] = zn[i]; k1[i] = k7[i]; }
        break;
      }
    }
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int i = 0; i < N; i++) m[i] = dead ? nanv : z[i];
#pragma unroll
    for (int i = 0; i < TN; i++) A[i] = dead ? nanv : z[N + i];
  }

  // ---- predicted manifests: images of the sigma points of (m, A) (getY_, R/modelTemplate.R:62-73, 353-358) --------
  struct Meas {
    Vec<M> Y0, Yp[N], Ym[N];
    double mu[M];
    bool miss[M];
    int o;
  };
  static __device__ __forceinline__ void predict_meas(const double (&p)[F], const double (&m)[N], const double (&A)[TN],
                                                      const double (&ob)[M], const UT& ut, int person, double tabs, Meas& z) {
    Vec<N> xc;
#pragma unroll
    for (int i = 0; i < N; i++) xc.v[i] = m[i];
    z.Y0.zeros();
    Mdl::measurementEquation(xc, z.Y0, p, person, tabs);  // t is ABSOLUTE time here (timeSum)
#pragma unroll
    for (int j = 0; j < N; j++) {
      Vec<N> xp, xm;
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i >= j) {
          xp.v[i] = fma(ut.sqrtc, A[tri(i, j)], m[i]);
          xm.v[i] = fma(-ut.sqrtc, A[tri(i, j)], m[i]);
        } else {
          xp.v[i] = m[i];
          xm.v[i] = m[i];
        }
      }
      z.Yp[j].zeros();
      z.Ym[j].zeros();
      Mdl::measurementEquation(xp, z.Yp[j], p, person, tabs);
      Mdl::measurementEquation(xm, z.Ym[j], p, person, tabs);
    }
#pragma unroll
    for (int a = 0; a < M; a++) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) acc += z.Yp[j].v[a] + z.Ym[j].v[a];
      z.mu[a] = ut.wm0 * z.Y0.v[a] + ut.wm1 * acc;  // predictMu_C
    }
    z.o = 0;
#pragma unroll
    for (int a = 0; a < M; a++) {
      z.miss[a] = !isfinite(ob[a]);
      z.o += z.miss[a] ? 0 : 1;
    }
  }

  // Z = C srS^-T with C = X W Y' = wc1 sqrt(c) Σ_j A_j (Yp_j − Ym_j)' (predictC_C; the column-0 term is (x_0 − X wm) = 0),
  // w = srS^-1 (obs − mu); returns dist = w'w and Σ log diag(srS) over the observed manifests.
  // K = Z srS^-1 (computeK_SR_C / computeK_C) is never formed:  K (obs − mu) = Z w,  K srS = Z,  K S K' = Z Z'.
  static __device__ __forceinline__ void gain(const double (&A)[TN], const Meas& z, const double (&ob)[M], const double (&S)[TM],
                                              const UT& ut, double (&Z)[N][M], double (&w)[M], double& dist, double& ld) {
    double rS[M];
#pragma unroll
    for (int a = 0; a < M; a++) rS[a] = rcp(S[tri(a, a)]);
    dist = 0.0;
    ld = 0.0;
#pragma unroll
    for (int a = 0; a < M; a++) {
      double acc = z.miss[a] ? 0.0 : (ob[a] - z.mu[a]);
#pragma unroll
      for (int k = 0; k < a; k++) acc -= S[tri(a, k)] * w[k];
      w[a] = acc * rS[a];
      dist += w[a] * w[a];
      ld += z.miss[a] ? 0.0 : log(S[tri(a, a)]);
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
      double crow[M];
#pragma unroll
      for (int a = 0; a < M; a++) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j <= i; j++) acc += A[tri(i, j)] * (z.Yp[j].v[a] - z.Ym[j].v[a]);
        crow[a] = z.miss[a] ? 0.0 : ut.kap * acc;
      }
#pragma unroll
      for (int a = 0; a < M; a++) {
        double acc = crow[a];
#pragma unroll
        for (int k = 0; k < a; k++) acc -= Z[i][k] * S[tri(a, k)];
        Z[i][a] = acc * rS[a];
      }
    }
  }

  // ---- measurement update, SR variant (R/modelTemplate.R:363-408).  Returns this time point's -2LL term. ---------
  static __device__ __forceinline__ double update_sr(const double (&p)[F], double (&m)[N], double (&A)[TN],
                                                     const double (&ob)[M], const UT& ut, const Meas& z) {
    // srS = chol factor of wc1 Σ_i (Y_i − mu)(·)ᵀ + Rchol Rcholᵀ  (predictSquareRootOfS_C: QR of [√wc1 (Y−mu), Rchol]ᵀ).
    // The QR is done as a streaming Givens triangularisation, one row of the stacked matrix at a time; its R
    // factor has a positive diagonal, which is the sign convention cholupdate_C normalises to anyway (T3).
    double S[TM];
#pragma unroll
    for (int i = 0; i < TM; i++) S[i] = 0.0;
    const double sw = sqrt(ut.wc1);
    // first: index of the row's first possibly non-zero entry (rotations before it would be identities)
    auto rotate_in = [&](double (&tr)[M], int first) {
#pragma unroll
      for (int k = 0; k < M; k++) {
        if (k < first) continue;
        const double a = S[tri(k, k)], b = tr[k];
        double r, c = 1.0, s = 0.0;
        if constexpr (RSQ_ROT) {
          const double r2 = a * a + b * b;
          const bool nz = r2 != 0.0;
          const double ri = nz ? rsq(r2) : 0.0;      // a = b = 0: identity rotation (r = 0, c = 1, s = 0)
          r = r2 * ri;
          s = b * ri;
          c = nz ? a * ri : 1.0;
        } else {
          r = sqrt(a * a + b * b);
          if (r != 0.0) {
            const double ri = rcp(r);
            c = a * ri;
            s = b * ri;
          }
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
      for (int a = 0; a < M; a++) tr[a] = z.miss[a] ? 0.0 : sw * (z.Y0.v[a] - z.mu[a]);
      rotate_in(tr, 0);
#pragma unroll
      for (int j = 0; j < N; j++) {
#pragma unroll
        for (int a = 0; a < M; a++) tr[a] = z.miss[a] ? 0.0 : sw * (z.Yp[j].v[a] - z.mu[a]);
        rotate_in(tr, 0);
      }
#pragma unroll
      for (int j = 0; j < N; j++) {
#pragma unroll
        for (int a = 0; a < M; a++) tr[a] = z.miss[a] ? 0.0 : sw * (z.Ym[j].v[a] - z.mu[a]);
        rotate_in(tr, 0);
      }
      // columns of Rchol = logChol2Chol(RLogChol) (R/modelTemplate.R:277-278), masked
      double Rc[M * M];
      Mdl::rchol(Rc, p);
#pragma unroll
      for (int b = 0; b < M; b++) {
        if constexpr (Mdl::R_DIAG) {
#pragma unroll
          for (int a = 0; a < M; a++) tr[a] = 0.0;
          tr[b] = z.miss[b] ? 1.0 : Rc[b + b * M];
          rotate_in(tr, b);
        } else {
#pragma unroll
          for (int a = 0; a < M; a++) tr[a] = (z.miss[a] || z.miss[b]) ? ((a == b) ? 1.0 : 0.0) : Rc[a + b * M];
          rotate_in(tr, 0);
        }
      }
    }
    // rank-1 up/down-date with column 0, weight |wc0| applied as v^(1/4) — reference quirk Q3, reproduced
    // literally (cholupdate_C, inst/include/psydiffInternals.h:133-165, 187-195)
    {
      const double dir = (ut.wc0 < 0.0) ? -1.0 : 1.0;
      const double v4 = sqrt(sqrt(fabs(ut.wc0)));
      double x[M];
#pragma unroll
      for (int a = 0; a < M; a++) x[a] = z.miss[a] ? 0.0 : v4 * (z.Y0.v[a] - z.mu[a]);
#pragma unroll
      for (int k = 0; k < M; k++) {
        const double lkk = S[tri(k, k)];
        const double r2 = lkk * lkk + dir * x[k] * x[k];
        double r, ic;
        if constexpr (RSQ_ROT) {
          const double ri = rsq(r2);
          r = r2 * ri;
          ic = lkk * ri;
        } else {
          r = sqrt(r2);
          ic = lkk * rcp(r);
        }
        const double il = rcp(lkk);
        const double c = r * il, s = x[k] * il;
        S[tri(k, k)] = r;
#pragma unroll
        for (int i = k + 1; i < M; i++) {
          const double lik = (S[tri(i, k)] + dir * s * x[i]) * ic;
          S[tri(i, k)] = lik;
          x[i] = c * x[i] - s * lik;
        }
      }
    }
    double Z[N][M], w[M], dist, ld;
    gain(A, z, ob, S, ut, Z, w, dist, ld);
#pragma unroll
    for (int i = 0; i < N; i++) {
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < M; a++) acc += Z[i][a] * w[a];
      m[i] += acc;  // updateM_C
    }
    // A ← cholupdate(A, U[:,b], 1, "downdate") for every manifest b (R/modelTemplate.R:382-386), U = K srS = Z
#pragma unroll
    for (int b = 0; b < M; b++) {
      if (z.miss[b]) continue;  // U[:,b] is exactly zero: the downdate would be an exact no-op
      double x[N];
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = Z[i][b];
#pragma unroll
      for (int k = 0; k < N; k++) {
        const double lkk = A[tri(k, k)];
        const double r2 = lkk * lkk - x[k] * x[k];
        double r, ic;
        if constexpr (RSQ_ROT) {
          const double ri = rsq(r2);
          r = r2 * ri;
          ic = lkk * ri;
        } else {
          r = sqrt(r2);
          ic = lkk * rcp(r);
        }
        const double il = rcp(lkk);
        const double c = r * il, s = x[k] * il;
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
    return (double)z.o * 1.8378770664093453 + 2.0 * ld + dist;
  }

  // ---- measurement update, covariance variant (srUpdate = FALSE, R/modelTemplate.R:785-835) -----------------------
  // S = Y W Y' + R with the textbook weights (predictS_C, Internals.h:81-83), symmetric by construction; its Cholesky
  // factor gives K = C S^-1 (computeK_C), log det S and the Mahalanobis term (computeIndividualM2LL_C, :103-115);
  // P = P_ − K S K' = P_ − Z Z' (updateP_C).  RR = Rchol Rchol' (full product, then the observed sub-block).
  static __device__ __forceinline__ double update_cov(double (&m)[N], const double (&A)[TN], double (&P)[TN],
                                                      const double (&RR)[TM], const double (&ob)[M], const UT& ut, const Meas& z) {
    double S[TM];
#pragma unroll
    for (int a = 0; a < M; a++)
#pragma unroll
      for (int b = 0; b <= a; b++) {
        double acc = ut.wc0 * (z.Y0.v[a] - z.mu[a]) * (z.Y0.v[b] - z.mu[b]);
        double s2 = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++)
          s2 += (z.Yp[j].v[a] - z.mu[a]) * (z.Yp[j].v[b] - z.mu[b]) + (z.Ym[j].v[a] - z.mu[a]) * (z.Ym[j].v[b] - z.mu[b]);
        acc += ut.wc1 * s2 + RR[tri(a, b)];
        S[tri(a, b)] = (z.miss[a] || z.miss[b]) ? ((a == b) ? 1.0 : 0.0) : acc;
      }
    // in-place Cholesky of S
#pragma unroll
    for (int j = 0; j < M; j++) {
      double d = S[tri(j, j)];
#pragma unroll
      for (int k = 0; k < j; k++) d -= S[tri(j, k)] * S[tri(j, k)];
      d = sqrt(d);
      S[tri(j, j)] = d;
      const double id = rcp(d);
#pragma unroll
      for (int i = j + 1; i < M; i++) {
        double a = S[tri(i, j)];
#pragma unroll
        for (int k = 0; k < j; k++) a -= S[tri(i, k)] * S[tri(j, k)];
        S[tri(i, j)] = a * id;
      }
    }
    double Z[N][M], w[M], dist, ld;
    gain(A, z, ob, S, ut, Z, w, dist, ld);
#pragma unroll
    for (int i = 0; i < N; i++) {
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < M; a++) acc += Z[i][a] * w[a];
      m[i] += acc;
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double zz = 0.0;
#pragma unroll
        for (int a = 0; a < M; a++) zz += Z[i][a] * Z[j][a];
        P[tri(i, j)] -= zz;
      }
    }
    return (double)z.o * 1.8378770664093453 + 2.0 * ld + dist;
  }

  // ---- prologue pieces shared with the lane-pair kernel (psy_filter_pair.cuh) -----------------------------------------
  // free-slot values of an instance: theta gathered through the person's slot map, one entry shifted by ±eps
  // (setParameterValues_C + setParameterList_C, inst/include/psydiffBasic.h:4-33, 63-121, stateless — quirk Q4)
  static __device__ __forceinline__ void gather(const PsyLaunch& L, const PsyInst& in, double (&p)[F]) {
    const int* ti = L.theta_index + (size_t)in.person * Mdl::NSLOT;
    const double* th = L.theta + ((in.code < -1) ? (size_t)(-1 - in.code) * (size_t)L.n_theta : (size_t)0);
    const int pert = (in.code >= 0) ? (in.code >> 1) : -1;
    const double sh = (in.code & 1) ? L.eps : -L.eps;
#pragma unroll
    for (int k = 0; k < F; k++) p[k] = 0.0;
#pragma unroll
    for (int k = 0; k < Mdl::NSLOT; k++) {
      const int t = __ldg(ti + k);
      double v = __ldg(th + t);
      if (t == pert) v += sh;
      p[k] = v;
    }
  }
  // m0, A0 = logChol2Chol(A0LogChol), Q = Q_SCALE · L Lᵀ with L = logChol2Chol(LLogChol)  (R/modelTemplate.R:275-276, 81)
  static __device__ __forceinline__ void initial_state(const double (&p)[F], double (&m)[N], double (&A)[TN], double (&Q)[QLEN]) {
    double A0[N * N], Ll[N * N];
    Mdl::initial(m, A0, Ll, p);
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) A[tri(i, j)] = (i == j) ? exp(A0[i + j * N]) : A0[i + j * N];  // logChol2Chol_C
    if constexpr (Mdl::L_STRUCT == L_DIAG) {
#pragma unroll
      for (int k = 0; k < N; k++) { const double l = exp(Ll[k + k * N]); Q[k] = Q_SCALE * (l * l); }
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
          Q[tri(i, j)] = Q_SCALE * acc;
        }
    }
  }

  // ---- one instance through all of its person's time points ------------------------------------------------
  static __device__ __forceinline__ void run(const PsyLaunch& L) {
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= L.n_inst) return;
    const PsyInst in = L.inst[ii];
    const int person = in.person;
    if (person < 0) return;  // warp-alignment padding slot
    double p[F];
    gather(L, in, p);
    UT ut;
    ut.sqrtc = L.sqrtc; ut.wm0 = L.wm0; ut.wm1 = L.wm1; ut.wc0 = L.wc0; ut.wc1 = L.wc1;
    ut.kap = L.wc1 * L.sqrtc;
    const int pid = __ldg(L.person_id + person);

    double m[N], A[TN], Q[QLEN];
    initial_state(p, m, A, Q);
    const long long r0 = __ldg(L.row_off + person);
    const int T = (int)(__ldg(L.row_off + person + 1) - r0);
    const bool base = (in.code == -1);
    // covariance variant state: P = A0 A0', R = Rchol Rchol' (R/modelTemplate.R:694-698)
    double P[Mdl::SR ? 1 : TN], RR[Mdl::SR ? 1 : TM];
    if constexpr (!Mdl::SR) {
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k <= j; k++) acc += A[tri(i, k)] * A[tri(j, k)];
          P[tri(i, j)] = acc;
        }
      double Rc[M * M];
      Mdl::rchol(Rc, p);
#pragma unroll
      for (int a = 0; a < M; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < M; k++) acc += Rc[a + k * M] * Rc[b + k * M];
          RR[tri(a, b)] = acc;
        }
    }
    double f = 0.0, tabs = 0.0;
#pragma unroll 1
    for (int tp = 0; tp < T; tp++) {
      const long long row = r0 + tp;
      const double dtv = __ldg(L.dt + row);
      tabs += dtv;
      if constexpr (!Mdl::SR) {
        // A = chol(P_, "lower") at every time point (R/modelTemplate.R:720); a non-positive pivot gives NaN
#pragma unroll
        for (int j = 0; j < N; j++) {
          double d = P[tri(j, j)];
#pragma unroll
          for (int k = 0; k < j; k++) d -= A[tri(j, k)] * A[tri(j, k)];
          d = sqrt(d);
          A[tri(j, j)] = d;
          const double id = rcp(d);
#pragma unroll
          for (int i = j + 1; i < N; i++) {
            double a = P[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) a -= A[tri(i, k)] * A[tri(j, k)];
            A[tri(i, j)] = a * id;
          }
        }
      }
      if (dtv != 0.0) {
        if constexpr (Mdl::STEPPER == 1) dopri5(p, m, A, Q, ut, pid, dtv, L.h);
        else rk4(p, m, A, Q, ut, pid, L.h, __ldg(L.ksteps + row));
        if constexpr (!Mdl::SR) {  // P_ = computePFromSigmaPoints_C (R/modelTemplate.R:771)
#pragma unroll
          for (int i = 0; i < N; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) {
              double acc = 0.0;
#pragma unroll
              for (int k = 0; k <= j; k++) acc += A[tri(i, k)] * A[tri(j, k)];
              P[tri(i, j)] = acc;
            }
        }
      }
      double ob[M];
#pragma unroll
      for (int a = 0; a < M; a++) ob[a] = __ldg(L.obs + row * M + a);
      Meas ms;
      predict_meas(p, m, A, ob, ut, pid, tabs, ms);
      if (base && L.pred) {
#pragma unroll
        for (int a = 0; a < M; a++) L.pred[row * M + a] = ms.mu[a];  // R/modelTemplate.R:361
      }
      if (ms.o > 0 && !L.skip_update) {
        if constexpr (Mdl::SR) f += update_sr(p, m, A, ob, ut, ms);
        else f += update_cov(m, A, P, RR, ob, ut, ms);
      }
      if (base && L.latent) {
#pragma unroll
        for (int i = 0; i < N; i++) L.latent[row * N + i] = m[i];
      }
    }
    L.f_out[L.out_idx ? L.out_idx[ii] : ii] = f;
  }
};

}  // namespace psy

#include "psy_filter_pair.cuh"

namespace psy {
// the kernel a model gets: the lane-pair form where it is eligible and enabled (psy_filter_pair.cuh), else one thread per instance
template <class Mdl>
struct Kernel {
  using P = FilterPair<Mdl>;
  using S = Filter<Mdl>;
  static constexpr bool PAIR = P::ENABLED;
  static constexpr int BLOCK = PAIR ? P::BLOCK : S::BLOCK;
  static constexpr int MIN_BLOCKS = PAIR ? P::MIN_BLOCKS : S::MIN_BLOCKS;
  static constexpr int SMEM_BYTES = PAIR ? P::SMEM_BYTES : (Mdl::STEPPER == 0 ? S::SMEM_BYTES : 0);
  static constexpr int SMEM_CARVEOUT_PCT = PAIR ? P::SMEM_CARVEOUT_PCT : (Mdl::STEPPER == 0 ? S::SMEM_CARVEOUT_PCT : -1);
  static __device__ __forceinline__ void run(const PsyLaunch& L) {
    if constexpr (PAIR) P::run(L);
    else S::run(L);
  }
};
}  // namespace psy

// psy_module_info = {stepper, srUpdate, nlatent, nmanifest, free slots, dynamic shared memory bytes, preferred shared-memory carveout in
// percent (-1 = the driver's choice), threads per block}: lets the host check a module against its settings and launch it as it was built;
// psy_module_lanes = lanes that integrate one instance in the RK4 stage loop (2 = the lane-pair kernel, psy_filter_pair.cuh)
#define PSY_INSTANTIATE(MODEL)                                                                                      \
  extern "C" __device__ const int psy_module_info[8] = {MODEL::STEPPER, MODEL::SR ? 1 : 0, MODEL::N, MODEL::M, MODEL::NSLOT, \
                                                        psy::Kernel<MODEL>::SMEM_BYTES, psy::Kernel<MODEL>::SMEM_CARVEOUT_PCT,  \
                                                        psy::Kernel<MODEL>::BLOCK};                                          \
  extern "C" __device__ const int psy_module_lanes = psy::Kernel<MODEL>::PAIR ? 2 : 1;                             \
  static constexpr int psy_block_size = psy::Kernel<MODEL>::BLOCK;                                                 \
  static constexpr int psy_lanes_per_instance_in_rk4 = psy::Kernel<MODEL>::PAIR ? 2 : 1;                           \
  extern "C" __global__ void __launch_bounds__(psy::Kernel<MODEL>::BLOCK, psy::Kernel<MODEL>::MIN_BLOCKS) psy_filter_kernel(const PsyLaunch L) {   \
    psy::Kernel<MODEL>::run(L);                                                                                     \
  }
