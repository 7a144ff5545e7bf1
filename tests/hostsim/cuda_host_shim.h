// cuda_host_shim.h — TEST INFRASTRUCTURE ONLY (used by tests/hostsim/hostsim.py, never by the product path).
//
// Lets g++ compile the model translation unit that compileModel hands to nvcc (psy_filter.cuh + psy_lang.cuh + the user's
// equations) as ordinary host C++, so the `-m "not gpu"` suite can run the KERNEL SOURCE ITSELF, one simulated thread at a
// time, against the oracle.  It checks the kernel's arithmetic and indexing on a machine without a GPU; it says nothing
// about performance and is not a fallback: psydiff_b200 itself never loads it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#ifdef PSY_SIM_BLOCK_THREADS
// block-level kernels (K3): one host thread per CUDA thread of ONE block at a time; __shared__ arrays become statics that the
// block's threads share, __syncthreads() a barrier over blockDim.x threads (psy_sim_barrier, set up by the launcher)
#define __shared__ static
void psy_sim_barrier();
#define __syncthreads() psy_sim_barrier()
#else
#define __shared__
#endif
#define __restrict__
#define __launch_bounds__(...)

struct psy_sim_dim3 { unsigned x, y, z; };
static thread_local psy_sim_dim3 threadIdx, blockIdx, blockDim, gridDim;

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, sizeof d); return d; }

// rcp.approx.ftz.f64: the hardware returns a reciprocal whose low 32 bits are zero (>= 20 good mantissa bits)
static inline double psy_sim_rcp_seed(double x) {
  double r = 1.0 / x;
  uint64_t b; memcpy(&b, &r, sizeof b);
  b &= 0xffffffff00000000ULL;
  memcpy(&r, &b, sizeof r);
  return r;
}

// rsqrt.approx.ftz.f64: same convention (low 32 bits zero)
static inline double psy_sim_rsqrt_seed(double x) {
  double r = 1.0 / sqrt(x);
  uint64_t b; memcpy(&b, &r, sizeof b);
  b &= 0xffffffff00000000ULL;
  memcpy(&r, &b, sizeof r);
  return r;
}

#ifdef PSY_COUNT_OPS
#include "op_counter.h"     // from here on `double` is the counting wrapper (sizeof 8, same layout): the kernel source below counts its flops
#endif

namespace psy { double psy_smem[1 << 14]; }  // dynamic shared memory of the ONE simulated thread block (single host thread)
