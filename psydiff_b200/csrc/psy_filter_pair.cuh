// psy_filter_pair.cuh — K1p, the lane-pair form of the fused SR-UKF filter kernel (hand-written FP64 CUDA for sm_100a).
//
// Same path, same numbers up to rounding as psy_filter.cuh (fitModel, R/modelTemplate.R:165-420; odeintModel::operator() :138-163;
// getMMatrix / computeL :77-135; integrate_const<runge_kutta4> :327-329), for models whose filter state no longer fits one thread's
// register file (n >= 7: 1.1 KB spill frame at n = 8, 3.2 KB at n = 10 in the thread-per-instance kernel).
//
// Mapping.  Every thread still OWNS one filter instance (person x perturbation): prologue, measurement update and -2LL run exactly
// as in psy_filter.cuh, one instance per lane.  The RK4 stage loop — 95 % of the arithmetic — is run by lane PAIRS: the two lanes
// (2q, 2q+1) first integrate lane 2q's instance together over its observation interval, then lane 2q+1's.  Inside the pair the
// reduced sigma-point state is split by COLUMNS of the triangular root A: lane r (= lane & 1) holds columns 2c+r, c = 0..n/2-1,
// stored with the row structure of the even column of the pair (rows 2c..n-1; the odd lane's row-2c entry is a structural zero
// carried as data), so both lanes execute the same fully unrolled code on registers with static indices.  The mean m is kept on
// both lanes (identical arithmetic on identical inputs -> bit-identical copies).
//
// RHS of the reduced ODE in the pair (fused column-scaled form, cf. Filter::rhs_fused):
//   own columns of Ã = A D⁻¹            -> one shuffle per entry gives both lanes the whole unit-lower Ã (E = even, O = odd columns)
//   own ROW j of W = Ã⁻¹ (e_jᵀ = wᵀ Ã)  -> the diffusion half ½ (Q Wᵀ)_j of the own column needs nothing else of W
//   drift at the own columns' sigma points, D'_j = kap A_jj (f(x+_j) − f(x−_j)) + ½ (Q Wᵀ)_j
//   H_j = Ã⁻¹ D'_j by forward substitution with Ã (no W), Σ (f+ + f−) all-reduced with one shuffle per component
//   Phi(M̃)_ij = H_ij + H_ji: 2 x 2 block transposes across the pair (one shuffle per block)
//   dA_j = Ã Phi_j (own columns), D⁻¹ folded into the RK4 coefficients.
// Per stage and lane at n = 8 (SASS of the stage loop): 835 FP64 instructions against 1389 of the thread-per-instance kernel, i.e.
// 20 % more arithmetic per instance (centre drift and mean update on both lanes, own-row recurrence of W, padded column structure),
// 92 SHFL, 75 FSEL, 56 LDS + 28 STS, and NO local-memory traffic (thread-per-instance: 92 LDL + 78 STL + 88 LDS + 44 STS per stage).
// The step's start state x_n and the running RK4 sum are parked in shared memory ([element][thread], conflict free).
// Measured on B200 (profiles/r02d..g_pair_variants.txt, profiles/r02e_*): n = 10: 3156 -> 1216 ms (2.6x); n = 8: 1371 -> ≈ 1300 ms at
// 4096 persons (FP64 pipe 52 % -> 65 % busy, which the 20 % extra arithmetic mostly eats); n = 6: 605 -> 841 ms and n = 4: 306 -> 463 ms
// (441 FP64 instructions per lane against 665 / 2: the state fits one thread there, so the pair only adds redundancy) — hence
// the default: pairs for even n >= 8 only.
//
// Control flow is kept WARP-uniform where shuffles are executed (full mask): time points run to the warp's longest series, the
// step loop to the warp's largest step count, and lanes/pairs that are past their own end compute but do not commit.
#pragma once

#ifndef PSY_PAIR
#define PSY_PAIR (-1)       // 1 = on where eligible, 0 = off, -1 = by model size (on for even n >= 8)
#endif
#ifndef PSY_PAIR_PARK
#define PSY_PAIR_PARK 2     // RK4 vectors parked in shared memory: 0 none, 1 the step's start state x_n, 2 x_n and the running sum
#endif
#ifndef PSY_PAIR_BLOCK
#define PSY_PAIR_BLOCK 64
#endif
#ifndef PSY_PAIR_DIRECT
#define PSY_PAIR_DIRECT 1   // 1: both lanes fetch Ã's even / odd columns with source-lane shuffles (2 per entry, no selects): 691.6 -> 681.6 ms
#endif
#ifndef PSY_PAIR_WORDER
#define PSY_PAIR_WORDER 0   // 1: the dot products of the W-row recurrence take the newest entry last (shorter dependency chain on
#endif                      // paper; measured SLOWER on B200: 702.0 vs 681.8 ms, profiles/r02g_pair_variants.txt)
#ifndef PSY_PAIR_MIN_BLOCKS
#define PSY_PAIR_MIN_BLOCKS 1   // resident blocks the register allocation is capped for (launch bounds); 168 registers (6 blocks,
                                // 3 warps per scheduler) measured 841.6 ms against 681.8 ms at 255: the stage loop then spills 97 values
#endif
#ifndef PSY_PAIR_CARVEOUT
#define PSY_PAIR_CARVEOUT (-1)
#endif

namespace psy {

#ifndef PSY_SIM_WARP
__device__ __forceinline__ double lane_xchg(double v) { return __shfl_xor_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double lane_from(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ int lane_from(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ void warp_sync() { __syncwarp(); }
__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31u); }
#endif  // the host simulation (tests/hostsim) supplies these over real threads

template <class Mdl>
struct FilterPair {
  using S = Filter<Mdl>;
  using UT = typename S::UT;
  static constexpr int N = Mdl::N, M = Mdl::M, F = Mdl::NFREE, TN = S::TN;
  static constexpr int NP = N / 2;             // column pairs
  static constexpr int TP = NP * (NP + 1);     // entries of A per lane: Σ_c (N − 2c)
  static constexpr int TW = NP * (NP + 1);     // entries of the own rows of W per lane: Σ_c (2c + 2)
  static constexpr int NS = N + TP;            // per-lane RK4 vector: m (both lanes) + own columns
  static constexpr bool ELIGIBLE = (N % 2 == 0) && N >= 4 && Mdl::STEPPER == 0 && Mdl::SR && Mdl::L_STRUCT == L_DIAG;
  static constexpr bool ENABLED = ELIGIBLE && (PSY_PAIR == 1 || (PSY_PAIR == -1 && N >= 8));
  static constexpr int BLOCK = PSY_PAIR_BLOCK;
  static constexpr int PARK = PSY_PAIR_PARK;
  static constexpr int STAGE_UNROLL = (PSY_STAGE_UNROLL > 0) ? PSY_STAGE_UNROLL : 1;   // see psy_filter.cuh: the pair loop is 18 KB of SASS already
  static constexpr int SMEM_BYTES = PARK * NS * BLOCK * 8;
  // shared-memory carveout (percent of 228 KB) that lets the 65536 / (256 · BLOCK) blocks the register file holds be resident
  static constexpr int MIN_BLOCKS = PSY_PAIR_MIN_BLOCKS;
  static constexpr int RESIDENT_BLOCKS = (MIN_BLOCKS > 256 / BLOCK) ? MIN_BLOCKS : 256 / BLOCK;
  static constexpr int SMEM_CARVEOUT_PCT = (PSY_PAIR_CARVEOUT >= 0) ? PSY_PAIR_CARVEOUT
      : (PARK == 0 ? -1 : (100 * RESIDENT_BLOCKS * (SMEM_BYTES + 1024) + 233471) / 233472);
  __host__ __device__ static constexpr int off(int c) { return c * N - c * (c - 1); }   // first entry of column pair c (row 2c)
  __host__ __device__ static constexpr int woff(int c) { return c * (c + 1); }          // first entry of the own W row of pair c
  // index of Ã(i, j), i > j, in the even (j = 2c) / odd (j = 2c+1) column views
  __host__ __device__ static constexpr int ati(int i, int j) { return off(j >> 1) + i - (j & ~1); }

  // ---- RHS on the pair: (m, own columns Ac) -> (dm on both lanes, own columns of dA·D, rdj = 1/A_jj of the own columns) ----
  static __device__ __forceinline__ void rhs_pair(const double (&p)[F], const double (&m)[N], const double (&Ac)[TP],
                                                  double (&dm)[N], double (&dAc)[TP], double (&rdj)[NP], const double (&QH)[N],
                                                  const UT& ut, int person, double t, const bool r) {
    double ajj[NP], E[TP], O[TP];
#pragma unroll
    for (int c = 0; c < NP; c++) {
      ajj[c] = r ? Ac[off(c) + 1] : Ac[off(c)];
      rdj[c] = rcp(ajj[c]);
    }
    // Ã = A D⁻¹: own columns scaled, partner's received; E / O are the same on both lanes
#pragma unroll
    for (int c = 0; c < NP; c++)
#pragma unroll
      for (int k = 1; k < N - 2 * c; k++) {
        const double own = Ac[off(c) + k] * rdj[c];
#if PSY_PAIR_DIRECT
        E[off(c) + k] = lane_from(own, lane_id() & ~1);
        if (k > 1) O[off(c) + k] = lane_from(own, lane_id() | 1);
#else
        const double oth = lane_xchg(own);
        E[off(c) + k] = r ? oth : own;
        O[off(c) + k] = r ? own : oth;   // k = 1 is the odd column's diagonal (≈ 1, never read)
#endif
      }
#define PSY_AT(i, j) ((((j) & 1) ? O : E)[ati((i), (j))])
    // own row j = 2c + r of W = Ã⁻¹ by back substitution of e_jᵀ = wᵀ Ã over the rows 0..2c+1 (entries above j come out as 0)
    double w[TW];
#pragma unroll
    for (int c = 0; c < NP; c++) {
      const int R = 2 * c + 1;
      w[woff(c) + R] = r ? 1.0 : 0.0;
      w[woff(c) + R - 1] = r ? -PSY_AT(R, R - 1) : 1.0;
#pragma unroll
      for (int k = R - 2; k >= 0; k--) {
#if PSY_PAIR_WORDER
        // the newest entry w_{k+1} enters LAST, so that only one FMA of each dot product waits for it
        double acc = -(w[woff(c) + R] * PSY_AT(R, k));
#pragma unroll
        for (int l = R - 1; l >= k + 1; l--) acc = fma(-w[woff(c) + l], PSY_AT(l, k), acc);
#else
        double acc = -(w[woff(c) + k + 1] * PSY_AT(k + 1, k));
#pragma unroll
        for (int l = k + 2; l <= R; l++) acc = fma(-w[woff(c) + l], PSY_AT(l, k), acc);
#endif
        w[woff(c) + k] = acc;
      }
    }
    // centre point (both lanes)
    Vec<N> xc, f0;
#pragma unroll
    for (int i = 0; i < N; i++) { xc.v[i] = m[i]; f0.v[i] = 0.0; }
    Mdl::latentEquation(xc, f0, p, person, t);
    // own columns: sigma points, drift, D'_j, H_j = Ã⁻¹ D'_j.  A drift component that reads none of x_2c.. (Mdl::dep) is not moved by
    // either column of the pair: f(x+) = f(x−) = f(x0) bit for bit, so its difference and sum are not evaluated.
    double H[NP * N], sacc[N];
#pragma unroll
    for (int i = 0; i < N; i++) sacc[i] = 0.0;
#pragma unroll
    for (int c = 0; c < NP; c++) {
      Vec<N> xp, xm, fp, fm;
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i >= 2 * c) {
          xp.v[i] = fma(ut.sqrtc, Ac[off(c) + i - 2 * c], m[i]);
          xm.v[i] = fma(-ut.sqrtc, Ac[off(c) + i - 2 * c], m[i]);
        } else {
          xp.v[i] = m[i];
          xm.v[i] = m[i];
        }
        fp.v[i] = 0.0;
        fm.v[i] = 0.0;
      }
      Mdl::latentEquation(xp, fp, p, person, t);
      Mdl::latentEquation(xm, fm, p, person, t);
      const double kaj = ut.kap * ajj[c];
      double d[N];
#pragma unroll
      for (int k = 0; k < N; k++) {
        const bool moved = (Mdl::dep(k) >> (2 * c)) != 0u;
        if (moved) sacc[k] += fp.v[k] + fm.v[k];
        if (k <= 2 * c + 1) {
          const double hq = QH[k] * w[woff(c) + k];
          d[k] = moved ? fma(kaj, fp.v[k] - fm.v[k], hq) : hq;
        } else {
          d[k] = moved ? kaj * (fp.v[k] - fm.v[k]) : 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        double acc = d[i];
#pragma unroll
        for (int k = 0; k < i; k++) acc = fma(-PSY_AT(i, k), H[c * N + k], acc);
        H[c * N + i] = acc;
      }
    }
    // weighted mean drift: the pair's partial sums are exchanged; the columns that move nothing contribute 2 wm1 f0 each
#pragma unroll
    for (int i = 0; i < N; i++) {
      int unmoved = 0;
#pragma unroll
      for (int c = 0; c < NP; c++) unmoved += ((Mdl::dep(i) >> (2 * c)) != 0u) ? 0 : 2;
      const double base = (ut.wm0 + 2.0 * (double)unmoved * ut.wm1) * f0.v[i];
      if (Mdl::dep(i) != 0u) {
        const double tot = sacc[i] + lane_xchg(sacc[i]);
        dm[i] = fma(ut.wm1, tot, base);
      } else {
        dm[i] = base;
      }
    }
    // Phi(M̃) of the own columns: Phi_ij = H_ij + H_ji below the diagonal, H_jj on it; H_ji comes from the column of row i's owner
    double Ph[TP];
#pragma unroll
    for (int c = 0; c < NP; c++) {
      const double h0 = H[c * N + 2 * c], h1 = H[c * N + 2 * c + 1];
      const double up = lane_xchg(h0);                 // even lane receives H(2c, 2c+1)
      Ph[off(c)] = r ? 0.0 : h0;
      Ph[off(c) + 1] = r ? h1 : h1 + up;
#pragma unroll
      for (int c2 = c + 1; c2 < NP; c2++) {
        const double a0 = H[c2 * N + 2 * c], a1 = H[c2 * N + 2 * c + 1];   // rows 2c, 2c+1 of the own column of pair c2
        const double send = r ? a0 : a1, keep = r ? a1 : a0;
        const double recv = lane_xchg(send);
        const double t0 = r ? recv : keep;             // H(j, 2c2)
        const double t1 = r ? keep : recv;             // H(j, 2c2 + 1)
        Ph[off(c) + 2 * (c2 - c)] = H[c * N + 2 * c2] + t0;
        Ph[off(c) + 2 * (c2 - c) + 1] = H[c * N + 2 * c2 + 1] + t1;
      }
    }
    // dA_j D = Ã Phi_j on the own columns (the odd lane's structural zero stays zero: Phi's entry above its diagonal is 0)
#pragma unroll
    for (int c = 0; c < NP; c++)
#pragma unroll
      for (int k = 0; k < N - 2 * c; k++) {
        double acc = Ph[off(c) + k];
#pragma unroll
        for (int k2 = 0; k2 < k; k2++) acc = fma(PSY_AT(2 * c + k, 2 * c + k2), Ph[off(c) + k2], acc);
        dAc[off(c) + k] = acc;
      }
#undef PSY_AT
  }

  // ---- classical RK4 on the pair: K fixed steps of the pair's current instance (rule T1); the loop runs to the warp's largest
  // count Kmax so that the shuffles stay warp-uniform, a pair past its own K recomputes and restores its state.
  static __device__ __forceinline__ void rk4_pair(const double (&p)[F], double (&m)[N], double (&Ac)[TP], const double (&QH)[N],
                                                  const UT& ut, int person, double h, int K, int Kmax, const bool r) {
    extern __shared__ double psy_smem[];
    volatile double* sx = psy_smem + threadIdx.x;                                  // x_n : sx[e * BLOCK]  (PARK >= 1)
    volatile double* sa = psy_smem + (PARK >= 2 ? NS * BLOCK : 0) + threadIdx.x;   // sum : sa[e * BLOCK]  (PARK == 2)
#pragma unroll 1
    for (int step = 0; step < Kmax; step++) {
      const double t0 = (double)step * h;
      double xn[PARK >= 1 ? 1 : NS], ac[PARK >= 2 ? 1 : NS];
#pragma unroll
      for (int i = 0; i < NS; i++) {
        const double v = (i < N) ? m[i] : Ac[i - N];
        if constexpr (PARK >= 1) sx[i * BLOCK] = v; else xn[i] = v;
        if constexpr (PARK >= 2) sa[i * BLOCK] = v; else ac[i] = v;
      }
#pragma unroll 1
      for (int sp = 0; sp < 4 / STAGE_UNROLL; sp++)
#pragma unroll
      for (int su = 0; su < STAGE_UNROLL; su++) {
        const int st = sp * STAGE_UNROLL + su;
        double km[N], kA[TP], rdj[NP];
        const double tc = (st == 0) ? t0 : ((st == 3) ? t0 + h : t0 + 0.5 * h);
        const double b = (st == 0 || st == 3) ? h * (1.0 / 6.0) : h * (1.0 / 3.0);
        rhs_pair(p, m, Ac, km, kA, rdj, QH, ut, person, tc, r);
        double ta[NS], tx[NS];
        if constexpr (PARK >= 2) {
#pragma unroll
          for (int i = 0; i < NS; i++) ta[i] = sa[i * BLOCK];
        } else {
#pragma unroll
          for (int i = 0; i < NS; i++) ta[i] = ac[i];
        }
        if (st < 3) {
          if constexpr (PARK >= 1) {
#pragma unroll
            for (int i = 0; i < NS; i++) tx[i] = sx[i * BLOCK];
          } else {
#pragma unroll
            for (int i = 0; i < NS; i++) tx[i] = xn[i];
          }
          const double a = (st == 2) ? h : 0.5 * h;
#pragma unroll
          for (int i = 0; i < N; i++) { ta[i] = fma(b, km[i], ta[i]); m[i] = fma(a, km[i], tx[i]); }
#pragma unroll
          for (int c = 0; c < NP; c++) {
            const double bj = b * rdj[c], aj = a * rdj[c];
#pragma unroll
            for (int k = 0; k < N - 2 * c; k++) {
              ta[N + off(c) + k] = fma(bj, kA[off(c) + k], ta[N + off(c) + k]);
              Ac[off(c) + k] = fma(aj, kA[off(c) + k], tx[N + off(c) + k]);
            }
          }
          if constexpr (PARK >= 2) {
#pragma unroll
            for (int i = 0; i < NS; i++) sa[i * BLOCK] = ta[i];
          } else {
#pragma unroll
            for (int i = 0; i < NS; i++) ac[i] = ta[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < N; i++) m[i] = fma(b, km[i], ta[i]);
#pragma unroll
          for (int c = 0; c < NP; c++) {
            const double bj = b * rdj[c];
#pragma unroll
            for (int k = 0; k < N - 2 * c; k++) Ac[off(c) + k] = fma(bj, kA[off(c) + k], ta[N + off(c) + k]);
          }
        }
      }
      if (step >= K) {   // this pair's interval has ended (another pair of the warp has more steps): undo the step
        if constexpr (PARK >= 1) {
#pragma unroll
          for (int i = 0; i < N; i++) m[i] = sx[i * BLOCK];
#pragma unroll
          for (int i = 0; i < TP; i++) Ac[i] = sx[(N + i) * BLOCK];
        } else {
#pragma unroll
          for (int i = 0; i < N; i++) m[i] = xn[i];
#pragma unroll
          for (int i = 0; i < TP; i++) Ac[i] = xn[N + i];
        }
      }
    }
  }

  // ---- every thread owns an instance; pairs share the integration -----------------------------------------------------------
  static __device__ __forceinline__ void run(const PsyLaunch& L) {
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = lane_id();
    const bool r = (lane & 1) != 0;
    PsyInst in;
    in.person = -1; in.code = -1;
    if (ii < L.n_inst) in = L.inst[ii];
    const bool valid = in.person >= 0;     // false: warp-alignment padding slot or past the end — the lane still helps its partner
    double p[F];
    if (valid) {
      S::gather(L, in, p);
    } else {
#pragma unroll
      for (int k = 0; k < F; k++) p[k] = 0.0;
    }
    UT ut;
    ut.sqrtc = L.sqrtc; ut.wm0 = L.wm0; ut.wm1 = L.wm1; ut.wc0 = L.wc0; ut.wc1 = L.wc1;
    ut.kap = L.wc1 * L.sqrtc;
    const int pid = valid ? __ldg(L.person_id + in.person) : 0;
    double m[N], A[TN], QH[N];
    {
      double Q[S::QLEN];
      S::initial_state(p, m, A, Q);
#pragma unroll
      for (int k = 0; k < N; k++) QH[k] = (0.5 / S::Q_SCALE) * Q[k];
    }
    const long long r0 = valid ? __ldg(L.row_off + in.person) : 0;
    const int T = valid ? (int)(__ldg(L.row_off + in.person + 1) - r0) : 0;
    const int Tmax = warp_max(T);
    const bool base = valid && (in.code == -1);
    double f = 0.0, tabs = 0.0;
#pragma unroll 1
    for (int tp = 0; tp < Tmax; tp++) {
      const bool live = tp < T;
      const long long row = r0 + tp;
      const double dtv = live ? __ldg(L.dt + row) : 0.0;
      tabs += dtv;
      const int Kown = (live && dtv != 0.0) ? __ldg(L.ksteps + row) : 0;
      // the two instances of the pair, one after the other
#pragma unroll 1
      for (int s = 0; s < 2; s++) {
        const int owner = (lane & ~1) | s;
        const int K = lane_from(Kown, owner);
        const int Kmax = warp_max(K);
        if (Kmax == 0) continue;
        double pc[F], qc[N], mc[N], Ac[TP];
#pragma unroll
        for (int k = 0; k < F; k++) pc[k] = lane_from(p[k], owner);
#pragma unroll
        for (int k = 0; k < N; k++) { qc[k] = lane_from(QH[k], owner); mc[k] = lane_from(m[k], owner); }
#pragma unroll
        for (int c = 0; c < NP; c++)
#pragma unroll
          for (int k = 0; k < N - 2 * c; k++) {
            const double e = lane_from(A[tri(2 * c + k, 2 * c)], owner);
            const double o = (k > 0) ? lane_from(A[tri(2 * c + k, 2 * c + 1)], owner) : 0.0;
            Ac[off(c) + k] = r ? o : e;
          }
        const int pidc = lane_from(pid, owner);
        rk4_pair(pc, mc, Ac, qc, ut, pidc, L.h, K, Kmax, r);
        // back to the owner: even columns from the even lane, odd columns from the odd lane
#pragma unroll
        for (int c = 0; c < NP; c++)
#pragma unroll
          for (int k = 0; k < N - 2 * c; k++) {
            const double e = lane_from(Ac[off(c) + k], lane & ~1);
            const double o = lane_from(Ac[off(c) + k], lane | 1);
            if (lane == owner) {
              A[tri(2 * c + k, 2 * c)] = e;
              if (k > 0) A[tri(2 * c + k, 2 * c + 1)] = o;
            }
          }
        if (lane == owner) {
#pragma unroll
          for (int i = 0; i < N; i++) m[i] = mc[i];
        }
      }
      warp_sync();
      if (live) {
        double ob[M];
#pragma unroll
        for (int a = 0; a < M; a++) ob[a] = __ldg(L.obs + row * M + a);
        typename S::Meas ms;
        S::predict_meas(p, m, A, ob, ut, pid, tabs, ms);
        if (base && L.pred) {
#pragma unroll
          for (int a = 0; a < M; a++) L.pred[row * M + a] = ms.mu[a];
        }
        if (ms.o > 0 && !L.skip_update) f += S::update_sr(p, m, A, ob, ut, ms);
        if (base && L.latent) {
#pragma unroll
          for (int i = 0; i < N; i++) L.latent[row * N + i] = m[i];
        }
      }
      warp_sync();
    }
    if (valid) L.f_out[L.out_idx ? L.out_idx[ii] : ii] = f;
  }
};

}  // namespace psy
