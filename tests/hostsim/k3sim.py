"""k3sim — TEST INFRASTRUCTURE ONLY: the reduction kernels of psydiff_b200/csrc/psy_reduce.cuh (K3) compiled by g++ and run on the
host: psy_reduce_chunks with one host thread per CUDA thread of a block and a real barrier for __syncthreads(), blocks one after
the other; psy_reduce_small / psy_reduce_final thread by thread.  Lets the CPU suite check the reduction's arithmetic, its
non-finite bookkeeping and its determinism against numpy."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "psydiff_b200", "csrc")
_BUILD = os.path.join(_HERE, "_build")

_SRC = r"""
#define PSY_SIM_BLOCK_THREADS 1
#include "cuda_host_shim.h"
#include <barrier>
#include <memory>
#include <thread>
#include <vector>
static std::unique_ptr<std::barrier<>> g_barrier;
void psy_sim_barrier() { g_barrier->arrive_and_wait(); }
#include "psy_reduce.cuh"

extern "C" void k3_reduce_chunks(const int* chunks4, int n_chunks, const int* pair_minus, const int* pair_base, const double* f, int P,
                                 double* chunk_part) {
  const Chunk* chunks = reinterpret_cast<const Chunk*>(chunks4);
  for (int b = 0; b < n_chunks; b++) {
    g_barrier.reset(new std::barrier<>(256));
    std::vector<std::thread> th;
    for (int t = 0; t < 256; t++)
      th.emplace_back([=] {
        blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)t; blockDim.x = 256; gridDim.x = (unsigned)n_chunks;
        psy_reduce_chunks(chunks, pair_minus, pair_base, f, P, chunk_part);
      });
    for (auto& x : th) x.join();
  }
}
extern "C" void k3_reduce_small(const int* small_list, int n_small, const long long* seg_off, const int* pair_minus, const int* pair_base,
                                const double* f, int P, double* partial) {
  blockDim.x = 128;
  for (int q = 0; q < (n_small + 127) / 128 * 128; q++) {
    blockIdx.x = (unsigned)(q / 128); threadIdx.x = (unsigned)(q % 128);
    psy_reduce_small(small_list, n_small, seg_off, pair_minus, pair_base, f, P, partial);
  }
}
extern "C" void k3_reduce_final(const int* big_list, int n_big, const int* theta_chunk_off, const double* chunk_part, int P, double* partial) {
  blockDim.x = 128;
  for (int q = 0; q < (n_big + 127) / 128 * 128; q++) {
    blockIdx.x = (unsigned)(q / 128); threadIdx.x = (unsigned)(q % 128);
    psy_reduce_final(big_list, n_big, theta_chunk_off, chunk_part, P, partial);
  }
}
"""


def _lib():
    text = _SRC + open(os.path.join(_CSRC, "psy_reduce.cuh")).read() + open(os.path.join(_HERE, "cuda_host_shim.h")).read()
    so = os.path.join(_BUILD, "k3sim_" + hashlib.sha1(text.encode()).hexdigest()[:16] + ".so")
    if not os.path.exists(so):
        os.makedirs(_BUILD, exist_ok=True)
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I", _HERE, "-I", _CSRC, "-o", so + ".tmp", "-x", "c++", "-"]
        r = subprocess.run(cmd, input=_SRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("k3sim: g++ failed:\n" + r.stderr[:4000])
        os.replace(so + ".tmp", so)
    return C.CDLL(so)


def lists(theta_index, P, pad=True, CH=4096, SMALL_SEG=64):
    """The instance layout and reduction lists of psy_set_slots (psy_capi.cu), restated: per person the base instance followed
    by (minus, plus) pairs of its distinct theta in slot order, warp-aligned person blocks; pairs sorted by theta then person;
    the base pseudo-segment (theta == P) last; chunks of <= CH entries for segments longer than SMALL_SEG."""
    N = theta_index.shape[0]
    start, distinct, n_inst = np.zeros(N, dtype=np.int64), [], 0
    for p in range(N):
        dl = list(dict.fromkeys(int(t) for t in theta_index[p]))
        distinct.append(dl)
        cnt = 1 + 2 * len(dl)
        padded = (cnt + 31) // 32 * 32
        align = pad and (padded - cnt) * 10 <= cnt
        if align:
            n_inst = (n_inst + 31) // 32 * 32
        start[p] = n_inst
        n_inst += padded if align else cnt
    seg = [[] for _ in range(P)]
    for p in range(N):
        for j, t in enumerate(distinct[p]):
            seg[t].append((int(start[p]) + 1 + 2 * j, int(start[p])))
    pair_minus = [a for s in seg for a, _ in s] + [int(o) for o in start]
    pair_base = [b for s in seg for _, b in s] + [int(o) for o in start]
    seg_off = np.concatenate([[0], np.cumsum([len(s) for s in seg])]).astype(np.int64)
    n_pairs = int(seg_off[-1])
    chunks, tco, small, big = [], [], [], []
    for t in range(P + 1):
        s0, s1 = (int(seg_off[t]), int(seg_off[t + 1])) if t < P else (n_pairs, n_pairs + N)
        tco.append(len(chunks))
        if t < P and s1 - s0 <= SMALL_SEG:
            small.append(t)
            continue
        big.append(t)
        for s in range(s0, s1, CH):
            chunks.append((t, s, min(CH, s1 - s), 0))
    tco.append(len(chunks))
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return dict(n_inst=n_inst, start=start, distinct=distinct, pair_minus=i32(pair_minus), pair_base=i32(pair_base), seg_off=seg_off,
                chunks=i32(chunks).reshape(-1, 4), tco=i32(tco), small=i32(small), big=i32(big))


def reduce(f, ls, P):
    """run_reduction of psy_capi.cu on the host: chunk partials, small segments, final -> partial[5P + 2]."""
    L = _lib()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    f = np.ascontiguousarray(f, dtype=np.float64)
    chunk_part = np.zeros(max(len(ls["chunks"]), 1) * 5)
    partial = np.zeros(5 * P + 2)
    if len(ls["chunks"]):
        L.k3_reduce_chunks(ip(ls["chunks"]), len(ls["chunks"]), ip(ls["pair_minus"]), ip(ls["pair_base"]), dp(f), P, dp(chunk_part))
    if len(ls["small"]):
        L.k3_reduce_small(ip(ls["small"]), len(ls["small"]), ls["seg_off"].ctypes.data_as(C.POINTER(C.c_longlong)), ip(ls["pair_minus"]),
                          ip(ls["pair_base"]), dp(f), P, dp(partial))
    L.k3_reduce_final(ip(ls["big"]), len(ls["big"]), ip(ls["tco"]), dp(chunk_part), P, dp(partial))
    return partial
