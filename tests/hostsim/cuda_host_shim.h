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
#include <stdlib.h>

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

// Warp collectives of the lane-pair kernel (psy_filter_pair.cuh): a simulated warp is PSY_SIM_WARP fibers (ucontext) of ONE host
// thread, run round robin; a lane that reaches a collective publishes its value and yields until every lane of the warp has
// arrived ("publish, barrier, read the source lane's value, barrier").  The kernel only uses lane ^ 1, lane & ~1, lane | 1 as
// sources and whole-warp maxima, so any even width is equivalent to 32.  threadIdx / blockIdx are per-fiber and swapped on a switch.
#ifndef PSY_SIM_WARP
#define PSY_SIM_WARP 4
#endif
#if !defined(__x86_64__)
#error "tests/hostsim's fiber switch is written for x86-64"
#endif
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#endif
// minimal fiber switch (callee-saved registers + stack pointer; no signal mask, no syscalls): psy_ctx_switch(&save_sp, load_sp)
extern "C" void psy_ctx_switch(void** save_sp, void* load_sp);
asm(".text\n.globl psy_ctx_switch\n.type psy_ctx_switch,@function\npsy_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size psy_ctx_switch, .-psy_ctx_switch\n");
struct psy_sim_warp_state {
  void* sp[PSY_SIM_WARP];
  void* main_sp;
  psy_sim_dim3 tid[PSY_SIM_WARP];
  bool done[PSY_SIM_WARP];
  char* stack[PSY_SIM_WARP];
  size_t stack_bytes;
  int cur;                      // running lane
  int arrived;                  // lanes waiting at the current barrier
  unsigned long long gen;       // barrier generation
  unsigned long long box[2][PSY_SIM_WARP];   // double buffered by generation: a lane is at most one collective ahead of the others
  void (*entry)(int);
};
static psy_sim_warp_state psy_sim_ws;
static int psy_sim_lane = 0;    // = psy_sim_ws.cur while a fiber runs
static inline void psy_sim_switch(void** from, void* to, const void* to_stack, size_t to_bytes) {
#if defined(__SANITIZE_ADDRESS__)
  void* fake = nullptr;
  __sanitizer_start_switch_fiber(&fake, to_stack, to_bytes);
  psy_ctx_switch(from, to);
  __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
  (void)to_stack; (void)to_bytes;
  psy_ctx_switch(from, to);
#endif
}
// hand the thread to the next unfinished lane (round robin); returns when this lane is scheduled again
static inline void psy_sim_yield() {
  psy_sim_warp_state& w = psy_sim_ws;
  const int me = w.cur;
  int nx = me;
  for (int k = 0; k < PSY_SIM_WARP; k++) {
    nx = (nx + 1) % PSY_SIM_WARP;
    if (!w.done[nx]) break;
  }
  if (nx == me) return;
  w.tid[me] = threadIdx;
  w.cur = nx; psy_sim_lane = nx; threadIdx = w.tid[nx];
  psy_sim_switch(&w.sp[me], w.sp[nx], w.stack[nx], w.stack_bytes);
}
static inline void psy_sim_warp_barrier() {
  psy_sim_warp_state& w = psy_sim_ws;
  const unsigned long long g = w.gen;
  if (++w.arrived == PSY_SIM_WARP) { w.arrived = 0; w.gen = g + 1; return; }
  while (w.gen == g) psy_sim_yield();
}
template <class T>
static inline T psy_sim_shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "");
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof v);
  const int buf = (int)(psy_sim_ws.gen & 1ULL);
  psy_sim_ws.box[buf][psy_sim_lane] = b;
  psy_sim_warp_barrier();
  b = psy_sim_ws.box[buf][src];
  T out;
  memcpy(&out, &b, sizeof out);
  return out;
}
extern "C" void psy_sim_fiber_main() {
  psy_sim_warp_state& w = psy_sim_ws;
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_finish_switch_fiber(nullptr, nullptr, nullptr);
#endif
  w.entry(w.cur);
  w.done[w.cur] = true;
  // a finished lane leaves the warp: every lane executes the same collectives, so nobody waits for it any more
  const int me = w.cur;
  int nx = -1;
  for (int k = 1; k <= PSY_SIM_WARP; k++) if (!w.done[(me + k) % PSY_SIM_WARP]) { nx = (me + k) % PSY_SIM_WARP; break; }
  if (nx < 0) {
    psy_sim_switch(&w.sp[me], w.main_sp, nullptr, 0);
  } else {
    w.cur = nx; psy_sim_lane = nx; threadIdx = w.tid[nx];
    psy_sim_switch(&w.sp[me], w.sp[nx], w.stack[nx], w.stack_bytes);
  }
  abort();   // a finished fiber is never resumed
}
// run one simulated warp: entry(lane) is called on PSY_SIM_WARP fibers with threadIdx.x = first_thread + lane
static inline void psy_sim_run_warp(void (*entry)(int), unsigned first_thread) {
  psy_sim_warp_state& w = psy_sim_ws;
  if (!w.stack[0]) {
    w.stack_bytes = (size_t)4 << 20;
    for (int l = 0; l < PSY_SIM_WARP; l++) w.stack[l] = (char*)aligned_alloc(64, w.stack_bytes);
  }
  w.entry = entry; w.arrived = 0;
  for (int l = 0; l < PSY_SIM_WARP; l++) {
    // initial frame: six callee-saved registers, the entry address psy_ctx_switch "returns" to, one slot that keeps the stack
    // 16-byte aligned the way a call would
    void** top = (void**)(w.stack[l] + w.stack_bytes);
    top[-1] = nullptr;
    top[-2] = (void*)&psy_sim_fiber_main;
    for (int k = 3; k <= 8; k++) top[-k] = nullptr;
    w.sp[l] = (void*)(top - 8);
    w.done[l] = false;
    w.tid[l] = threadIdx; w.tid[l].x = first_thread + (unsigned)l;
  }
  w.cur = 0; psy_sim_lane = 0; threadIdx = w.tid[0];
  psy_sim_switch(&w.main_sp, w.sp[0], w.stack[0], w.stack_bytes);
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
