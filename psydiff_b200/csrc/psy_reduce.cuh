// psy_reduce.cuh — K3, the deterministic reduction of instance results into per-parameter sums (part of libpsydiff_b200.so;
// included by psy_capi.cu).  Replaces the summation inside getGradients — `sum(m2LL)` per side and parameter, R/modelTemplate.R:
// 885-926 — for all parameters at once: D± = Σ_p (f±_p − f0_p) with the differences formed per person before summing (SURVEY
// H5), and the non-finite counts that drive the eps[] fallback.  Fixed thread -> entry assignment and fixed trees: the result is
// bit-reproducible for a given shard layout.  Kept in its own header so that the CPU suite can compile the same source for the
// host and run it block by block (tests/hostsim, tests/test_kernel_hostsim.py).
#pragma once

namespace {

struct Chunk { int theta; int start; int len; int pad; };

__device__ __forceinline__ bool finite_d(double x) { return isfinite(x); }

// one block per chunk (theta, start, len): fixed thread->entry assignment and a fixed shared-memory tree, so the
// result is bit-reproducible run to run.  theta == P is the pseudo-segment of base instances (Σ f0).
__global__ void __launch_bounds__(256) psy_reduce_chunks(const Chunk* __restrict__ chunks, const int* __restrict__ pair_minus,
                                                         const int* __restrict__ pair_base, const double* __restrict__ f,
                                                         int P, double* __restrict__ chunk_part) {
  __shared__ double sh[5][256];
  const Chunk c = chunks[blockIdx.x];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (int e = threadIdx.x; e < c.len; e += 256) {
    const int idx = c.start + e;
    const int ib = pair_base[idx];
    const double f0 = f[ib];
    if (c.theta == P) {
      a0 += f0;
      a1 += finite_d(f0) ? 0.0 : 1.0;
    } else {
      const int im = pair_minus[idx];
      const double fm = f[im], fp = f[im + 1];
      const double f0s = finite_d(f0) ? f0 : 0.0;
      a0 += fp - f0s;                      // D+
      a1 += fm - f0s;                      // D-
      a2 += finite_d(fp) ? 0.0 : 1.0;      // nf+
      a3 += finite_d(fm) ? 0.0 : 1.0;      // nf-
      a4 += finite_d(f0) ? 0.0 : 1.0;      // nbt
    }
  }
  sh[0][threadIdx.x] = a0; sh[1][threadIdx.x] = a1; sh[2][threadIdx.x] = a2; sh[3][threadIdx.x] = a3; sh[4][threadIdx.x] = a4;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int q = 0; q < 5; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 5) chunk_part[(size_t)blockIdx.x * 5 + threadIdx.x] = sh[threadIdx.x][0];
}

// parameters carried by few persons (person- or small-group-specific, e.g. C3's 250 000 "mu | person" labels): one thread
// per parameter walks its <= SMALL_SEG entries in order — no chunk stage, still a fixed summation order
__global__ void psy_reduce_small(const int* __restrict__ small_list, int n_small, const long long* __restrict__ seg_off,
                                 const int* __restrict__ pair_minus, const int* __restrict__ pair_base,
                                 const double* __restrict__ f, int P, double* __restrict__ partial) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_small) return;
  const int th = small_list[q];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (long long idx = seg_off[th]; idx < seg_off[th + 1]; idx++) {
    const double f0 = f[pair_base[idx]];
    const int im = pair_minus[idx];
    const double fm = f[im], fp = f[im + 1];
    const double f0s = finite_d(f0) ? f0 : 0.0;
    a0 += fp - f0s;
    a1 += fm - f0s;
    a2 += finite_d(fp) ? 0.0 : 1.0;
    a3 += finite_d(fm) ? 0.0 : 1.0;
    a4 += finite_d(f0) ? 0.0 : 1.0;
  }
  partial[2 + th] = a0;
  partial[2 + (size_t)P + th] = a1;
  partial[2 + 2 * (size_t)P + th] = a2;
  partial[2 + 3 * (size_t)P + th] = a3;
  partial[2 + 4 * (size_t)P + th] = a4;
}

// one thread per chunked parameter (and one for the base pseudo-segment): sequential sum of its chunks' partials
__global__ void psy_reduce_final(const int* __restrict__ big_list, int n_big, const int* __restrict__ theta_chunk_off,
                                 const double* __restrict__ chunk_part, int P, double* __restrict__ partial) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_big) return;
  const int th = big_list[q];
  double a[5] = {0, 0, 0, 0, 0};
  for (int c = theta_chunk_off[th]; c < theta_chunk_off[th + 1]; c++)
    for (int v = 0; v < 5; v++) a[v] += chunk_part[(size_t)c * 5 + v];
  if (th == P) {
    partial[0] = a[0];
    partial[1] = a[1];
  } else {
    partial[2 + th] = a[0];
    partial[2 + (size_t)P + th] = a[1];
    partial[2 + 2 * (size_t)P + th] = a[2];
    partial[2 + 3 * (size_t)P + th] = a[3];
    partial[2 + 4 * (size_t)P + th] = a[4];
  }
}

}  // namespace
