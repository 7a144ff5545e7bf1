// psy_k2.cuh — K2, the device-side builders of everything the filter launch and the K3 reduction index by:
//   * the RK4 step count of every row (Boost.odeint's integrate_const rule T1, exact IEEE, no contraction),
//   * the (person x perturbation) instance list that folds getGradients' 2P+1 fitModel sweeps into one launch
//     (R/modelTemplate.R:852-930) — one thread per person, person-major, warp-aligned person blocks,
//   * the per-parameter pair lists K3 reduces over (sorted by theta, then person: a stable radix sort),
//   * the chunk / small / big lists of K3,
//   * the compacted sub-list of a restricted sweep (one-sided direction, eps[] fallback of R/modelTemplate.R:885-898).
// It replaces the string-keyed setParameterValues_C / setParameterList_C walk of inst/include/psydiffBasic.h:63-121 as the
// thing that decides which filter runs see which parameter — here once per slot map, on the device, with no host vector
// proportional to the number of instances.  Part of libpsydiff_b200.so (included by psy_capi.cu, which adds the two library
// calls around these kernels: cub's stable radix sort of the pairs and cub's select for the restricted sweeps).  Plain kernels
// only, so the CPU suite can compile this header for the host and run it thread by thread (tests/hostsim/k2sim.py).
#pragma once
#include <cfloat>

#include "psy_launch.h"
#include "psy_reduce.cuh"

namespace k2 {

constexpr int PERSON_CHUNK = 1024;  // persons one layout-scan block walks
constexpr int CHUNK_LEN = 4096;     // entries of one K3 chunk
constexpr int SMALL_SEG = 64;       // parameters carried by at most this many persons skip the chunk stage

// ---- rule T1 ------------------------------------------------------------------------------------------------------
// steps are taken for k = 0, 1, ... while ((0 + k*h) + h) - t1 <= DBL_EPSILON, every operation rounded to double on its own
// (the intrinsics keep ptxas from contracting the multiply-add); the predicate is monotone in k, so the count is found from
// the guess floor(t1/h) by local search.  Same integer as the host's psy_rk4_step_count.
__device__ __forceinline__ bool t1_cond(int k, double h, double t1) {
  const double time = __dadd_rn(0.0, __dmul_rn((double)k, h));
  const double next = __dadd_rn(time, h);
  return __dadd_rn(next, -t1) <= DBL_EPSILON;
}
__global__ void step_counts(const double* __restrict__ dt, long long rows, double h, int* __restrict__ ksteps) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const double t1 = dt[r];
  int k = 0;
  if (h > 0 && t1 == t1 && t1 > 0) {
    double q = floor(t1 / h);
    if (q > 5e7) q = 5e7;
    k = (int)q;
    while (k > 0 && !t1_cond(k - 1, h, t1)) --k;
    while (k < 50000000 && t1_cond(k, h, t1)) ++k;
  }
  ksteps[r] = k;
}

// ---- distinct theta of each person, in slot order -----------------------------------------------------------------
// dl[p*F + j], j < dcnt[p]: the j-th distinct theta index among person p's F slots (first appearance).  One thread per
// person; F is a few dozen, the quadratic scan is register/L1 work.
__global__ void distinct_theta(const int* __restrict__ tidx, int N, int F, int* __restrict__ dl, int* __restrict__ dcnt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int* row = tidx + (size_t)p * F;
  int* out = dl + (size_t)p * F;
  int d = 0;
  for (int k = 0; k < F; k++) {
    const int t = row[k];
    bool seen = false;
    for (int q = 0; q < d; q++) seen = seen || (out[q] == t);
    if (!seen) out[d++] = t;
  }
  dcnt[p] = d;
}

// ---- layout of the instance list ----------------------------------------------------------------------------------
// A person's block is 1 + 2 d instances (base, then minus/plus per distinct theta).  It is padded to a multiple of 32 and
// started on a multiple of 32 when that idles < 10 % of its lanes: the lanes of a warp then share dt and missingness and the
// RK4 step loop does not diverge.  The running offset o -> o' is of the form o + delta(o mod 32), which composes
// associatively: phase A tabulates delta for every start residue of a 1024-person chunk (lane r = residue r), phase B walks
// the chunks, phase C walks each chunk from its true start.
__device__ __forceinline__ long long advance(long long o, int d, long long* start) {
  const long long cnt = 1 + 2 * (long long)d, padded = (cnt + 31) / 32 * 32;
  const bool align = (padded - cnt) * 10 <= cnt;
  if (align) o = (o + 31) / 32 * 32;
  if (start) *start = o;
  return o + (align ? padded : cnt);
}
__global__ void layout_tabulate(const int* __restrict__ dcnt, int N, long long* __restrict__ delta /*[chunks][32]*/) {
  const int c = blockIdx.x, r = threadIdx.x;
  const int p0 = c * PERSON_CHUNK, p1 = min(N, p0 + PERSON_CHUNK);
  long long o = r;
  for (int p = p0; p < p1; p++) o = advance(o, dcnt[p], nullptr);
  delta[(size_t)c * 32 + r] = o - r;
}
__global__ void layout_chunks(const long long* __restrict__ delta, int n_chunks, long long* __restrict__ chunk_start, long long* __restrict__ totals) {
  if (blockIdx.x || threadIdx.x) return;
  long long o = 0;
  for (int c = 0; c < n_chunks; c++) {
    chunk_start[c] = o;
    o += delta[(size_t)c * 32 + (int)(o & 31)];
  }
  totals[0] = o;  // n_inst including padding
}
__global__ void layout_persons(const int* __restrict__ dcnt, int N, const long long* __restrict__ chunk_start, long long* __restrict__ inst_start) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int p0 = c * PERSON_CHUNK;
  if (p0 >= N) return;
  const int p1 = min(N, p0 + PERSON_CHUNK);
  long long o = chunk_start[c];
  for (int p = p0; p < p1; p++) {
    long long s;
    o = advance(o, dcnt[p], &s);
    inst_start[p] = s;
  }
}

// single-block exclusive scan (int -> long long), n up to a few million: thread-contiguous segments + a block scan of the
// 1024 segment sums.  out[n] receives the total.
__global__ void __launch_bounds__(1024) scan_exclusive(const int* __restrict__ in, long long n, long long* __restrict__ out) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const long long seg = (n + 1023) / 1024, a = min(n, t * seg), b = min(n, a + seg);
  long long s = 0;
  for (long long i = a; i < b; i++) s += in[i];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const long long v = (t >= off) ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = part[t] - s;
  for (long long i = a; i < b; i++) { out[i] = run; run += in[i]; }
  if (t == 1023) out[n] = part[1023];
}

// ---- expansion: one thread per person writes its block of the instance list and its entries of the unsorted pair list ---
// (the list was pre-filled with {-1, -1} = padding slots)
__global__ void expand(const int* __restrict__ dl, const int* __restrict__ dcnt, const long long* __restrict__ inst_start,
                       const long long* __restrict__ pair_off, int N, int F, PsyInst* __restrict__ inst, PsyInst* __restrict__ base,
                       int* __restrict__ key, unsigned long long* __restrict__ val) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const long long o = inst_start[p];
  const long long e0 = pair_off[p];
  const int d = dcnt[p];
  inst[o] = PsyInst{p, -1};
  base[p] = PsyInst{p, -1};
  for (int j = 0; j < d; j++) {
    const int t = dl[(size_t)p * F + j];
    inst[o + 1 + 2 * j] = PsyInst{p, (t << 1) | 0};
    inst[o + 2 + 2 * j] = PsyInst{p, (t << 1) | 1};
    key[e0 + j] = t;
    val[e0 + j] = ((unsigned long long)(unsigned)(o + 1 + 2 * j) << 32) | (unsigned)o;  // (index of the minus instance, index of the base instance)
  }
}
// sorted (theta, person) pairs -> K3's pair lists; the base pseudo-segment (theta == P) follows at n_pairs
__global__ void unpack_pairs(const unsigned long long* __restrict__ val, long long n_pairs, const long long* __restrict__ inst_start, int N,
                             int* __restrict__ pair_minus, int* __restrict__ pair_base) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pairs) {
    const unsigned long long v = val[i];
    pair_minus[i] = (int)(v >> 32);
    pair_base[i] = (int)(v & 0xffffffffu);
  } else if (i < n_pairs + N) {
    const int o = (int)inst_start[i - n_pairs];
    pair_minus[i] = o;
    pair_base[i] = o;
  }
}
// seg_off[t] = first sorted pair with key >= t, t = 0..P  (seg_off[P] = n_pairs)
__global__ void segment_offsets(const int* __restrict__ key_sorted, long long n_pairs, int P, long long* __restrict__ seg_off) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > P) return;
  long long lo = 0, hi = n_pairs;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (key_sorted[mid] < t) lo = mid + 1; else hi = mid;
  }
  seg_off[t] = lo;
}
// per theta (t == P: the base pseudo-segment of N entries): chunk count and small/big flags
__global__ void chunk_counts(const long long* __restrict__ seg_off, int P, int N, int* __restrict__ n_chunks_of, int* __restrict__ is_small, int* __restrict__ is_big) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > P) return;
  const long long len = (t < P) ? seg_off[t + 1] - seg_off[t] : (long long)N;
  const bool small = (t < P) && len <= SMALL_SEG;
  n_chunks_of[t] = small ? 0 : (int)((len + CHUNK_LEN - 1) / CHUNK_LEN);
  is_small[t] = small ? 1 : 0;
  is_big[t] = small ? 0 : 1;
}
__global__ void chunk_fill(const long long* __restrict__ seg_off, const long long* __restrict__ tco64, const long long* __restrict__ small_pos,
                           const long long* __restrict__ big_pos, const int* __restrict__ is_small, int P, int N, long long n_pairs,
                           Chunk* __restrict__ chunks, int* __restrict__ tco, int* __restrict__ small_list, int* __restrict__ big_list) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > P + 1) return;
  tco[t] = (int)tco64[t];
  if (t > P) return;
  if (is_small[t]) { small_list[small_pos[t]] = t; return; }
  big_list[big_pos[t]] = t;
  const long long s0 = (t < P) ? seg_off[t] : n_pairs, s1 = (t < P) ? seg_off[t + 1] : n_pairs + N;
  long long c = tco64[t];
  for (long long s = s0; s < s1; s += CHUNK_LEN, c++) chunks[c] = Chunk{t, (int)s, (int)min((long long)CHUNK_LEN, s1 - s), 0};
}

// ---- restricted sweeps: which slots of the full instance list run ---------------------------------------------------
struct WantInstance {
  const PsyInst* inst;
  const unsigned char* need;  // [2P] by (theta << 1 | side), or null = every perturbed side the direction asks for
  int with_base, left, right;
  __host__ __device__ __forceinline__ bool operator()(int i) const {
    const PsyInst in = inst[i];
    if (in.person < 0) return false;
    if (in.code < 0) return with_base != 0;
    if (need) return need[in.code] != 0;
    return (in.code & 1) ? (right != 0) : (left != 0);
  }
};
__global__ void gather_instances(const PsyInst* __restrict__ inst, const int* __restrict__ idx, const int* __restrict__ n_sel, PsyInst* __restrict__ sub) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < *n_sel) sub[k] = inst[idx[k]];
}
// psy_fit_many: n_sets unperturbed instances per person, set-major
__global__ void many_list(int N, int n_sets, PsyInst* __restrict__ list) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * n_sets) return;
  list[i] = PsyInst{(int)(i % N), -1 - (int)(i / N)};
}
// per parameter set: Σ of the finite -2LL and the count of the others, one block per (set, 4096-person chunk), fixed tree
__global__ void __launch_bounds__(256) sets_chunks(const double* __restrict__ f, int N, int n_chunks, double* __restrict__ part /*[set][chunk][2]*/) {
  __shared__ double sh[2][256];
  const int k = blockIdx.x / n_chunks, c = blockIdx.x % n_chunks;
  const int p0 = c * CHUNK_LEN, p1 = min(N, p0 + CHUNK_LEN);
  double s = 0, bad = 0;
  for (int p = p0 + threadIdx.x; p < p1; p += 256) {
    const double v = f[(size_t)k * N + p];
    if (v == v) s += v; else bad += 1.0;   // NaN = failed person (anyNA); ±inf stays in the sum
  }
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = bad;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) { sh[0][threadIdx.x] += sh[0][threadIdx.x + st]; sh[1][threadIdx.x] += sh[1][threadIdx.x + st]; }
    __syncthreads();
  }
  if (threadIdx.x < 2) part[((size_t)k * n_chunks + c) * 2 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void sets_final(const double* __restrict__ part, int n_sets, int n_chunks, double* __restrict__ out /*[2][n_sets]*/) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sets) return;
  double s = 0, bad = 0;
  for (int c = 0; c < n_chunks; c++) { s += part[((size_t)k * n_chunks + c) * 2]; bad += part[((size_t)k * n_chunks + c) * 2 + 1]; }
  out[k] = s;
  out[n_sets + k] = bad;
}
// out[i] = Σ_s in[s*len + i] over the devices' partial vectors, in device order (deterministic)
__global__ void sum_shards(const double* __restrict__ in, int n_shards, long long len, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = in[i];
  for (int k = 1; k < n_shards; k++) s += in[(size_t)k * len + i];
  out[i] = s;
}

}  // namespace k2
