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

namespace psy { double psy_smem[1 << 14]; }  // dynamic shared memory of the ONE simulated thread block

// Warp collectives of the lane-pair kernel (psy_filter_pair.cuh): a simulated warp is PSY_SIM_WARP host threads that really run
// concurrently; a shuffle is "publish my value, barrier, read the source lane's, barrier".  The kernel only uses lane ^ 1,
// lane & ~1, lane | 1 as sources and whole-warp maxima, so any even width is equivalent to 32.
#ifndef PSY_SIM_WARP
#define PSY_SIM_WARP 4
#endif
#include <atomic>
struct psy_sim_warp_state {
  std::atomic<int> arrived{0};
  std::atomic<int> sense{0};
  unsigned long long box[PSY_SIM_WARP];
};
static psy_sim_warp_state psy_sim_ws;
static thread_local int psy_sim_lane = 0, psy_sim_local_sense = 0;
static inline void psy_sim_warp_barrier() {
  psy_sim_local_sense ^= 1;
  if (psy_sim_ws.arrived.fetch_add(1, std::memory_order_acq_rel) == PSY_SIM_WARP - 1) {
    psy_sim_ws.arrived.store(0, std::memory_order_relaxed);
    psy_sim_ws.sense.store(psy_sim_local_sense, std::memory_order_release);
  } else {
    while (psy_sim_ws.sense.load(std::memory_order_acquire) != psy_sim_local_sense) {}
  }
}
template <class T>
static inline T psy_sim_shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "");
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof v);
  psy_sim_ws.box[psy_sim_lane] = b;
  psy_sim_warp_barrier();
  b = psy_sim_ws.box[src];
  psy_sim_warp_barrier();
  T out;
  memcpy(&out, &b, sizeof out);
  return out;
}
namespace psy {
static inline double lane_xchg(double v) { return psy_sim_shfl(v, psy_sim_lane ^ 1); }
static inline double lane_from(double v, int src) { return psy_sim_shfl(v, src); }
static inline int lane_from(int v, int src) { return psy_sim_shfl(v, src); }
static inline int warp_max(int v) {
  int mx = v;
  for (int l = 0; l < PSY_SIM_WARP; l++) { const int o = psy_sim_shfl(v, l); mx = o > mx ? o : mx; }
  return mx;
}
static inline void warp_sync() { psy_sim_warp_barrier(); }
static inline int lane_id() { return psy_sim_lane; }
}  // namespace psy
