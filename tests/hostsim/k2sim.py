"""k2sim — TEST INFRASTRUCTURE ONLY: the K2 builder kernels of psydiff_b200/csrc/psy_k2.cuh compiled by g++ and run on the host
in the order psy_capi.cu launches them (build_slots_on): thread by thread for the per-person / per-parameter kernels, one host
thread per CUDA thread with a real barrier for the single-block scan.  cub's stable radix sort of the (theta, person) pairs —
the one library call between them — is a std::stable_sort here.  Lets the CPU suite check the device-built instance list, pair
lists and chunk lists against the plain restatement in k3sim.lists, and the device's RK4 step counts against the host rule."""
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
#include <algorithm>
#include <barrier>
#include <memory>
#include <numeric>
#include <thread>
#include <vector>
using std::min;
using std::max;
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static std::unique_ptr<std::barrier<>> g_barrier;
void psy_sim_barrier() { g_barrier->arrive_and_wait(); }
#include "psy_k2.cuh"

// run `fn` for every thread of a <<<grid, block>>> launch, one after the other (kernels without __syncthreads)
template <class Fn>
static void launch_serial(int grid, int block, Fn fn) {
  blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
  for (int b = 0; b < grid; b++)
    for (int t = 0; t < block; t++) { blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)t; fn(); }
}
// one block of `block` concurrent host threads with a barrier (kernels with __syncthreads)
template <class Fn>
static void launch_block(int block, Fn fn) {
  g_barrier.reset(new std::barrier<>(block));
  std::vector<std::thread> th;
  for (int t = 0; t < block; t++)
    th.emplace_back([=] { blockIdx.x = 0; threadIdx.x = (unsigned)t; blockDim.x = (unsigned)block; gridDim.x = 1; fn(); });
  for (auto& x : th) x.join();
}
static int blocks_for(long long n, int block) { return (int)std::max<long long>(1, (n + block - 1) / block); }

extern "C" void k2_step_counts(const double* dt, long long rows, double h, int* ksteps) {
  launch_serial(blocks_for(rows, 256), 256, [&] { k2::step_counts(dt, rows, h, ksteps); });
}

// out_counts = {n_inst, n_pairs, n_chunks, n_small, n_big}; the arrays are sized by the caller from a first call with null outputs
extern "C" void k2_build(const int* tidx, int N, int F, int P, long long* out_counts, long long* inst_start_out, int* inst2 /*n_inst x 2*/,
                         int* base2 /*N x 2*/, int* pair_minus, int* pair_base, long long* seg_off, int* chunks4, int* tco, int* small_list,
                         int* big_list) {
  const int n_pchunks = (N + k2::PERSON_CHUNK - 1) / k2::PERSON_CHUNK;
  std::vector<int> dl((size_t)N * std::max(F, 1)), dcnt(N);
  std::vector<long long> pair_off(N + 1), inst_start(N), delta((size_t)n_pchunks * 32), chunk_start(n_pchunks), totals(4);
  launch_serial(blocks_for(N, 128), 128, [&] { k2::distinct_theta(tidx, N, F, dl.data(), dcnt.data()); });
  launch_serial(n_pchunks, 32, [&] { k2::layout_tabulate(dcnt.data(), N, delta.data()); });
  launch_serial(1, 32, [&] { k2::layout_chunks(delta.data(), n_pchunks, chunk_start.data(), totals.data()); });
  launch_serial(blocks_for(n_pchunks, 64), 64, [&] { k2::layout_persons(dcnt.data(), N, chunk_start.data(), inst_start.data()); });
  launch_block(1024, [&] { k2::scan_exclusive(dcnt.data(), N, pair_off.data()); });
  const long long n_inst = totals[0], n_pairs = pair_off[N];
  out_counts[0] = n_inst; out_counts[1] = n_pairs;
  if (!inst2) return;
  for (int p = 0; p < N; p++) inst_start_out[p] = inst_start[p];
  std::vector<PsyInst> inst((size_t)n_inst, PsyInst{-1, -1}), base(N);
  std::vector<int> key(n_pairs), key_sorted(n_pairs);
  std::vector<unsigned long long> val(n_pairs), val_sorted(n_pairs);
  launch_serial(blocks_for(N, 128), 128, [&] { k2::expand(dl.data(), dcnt.data(), inst_start.data(), pair_off.data(), N, F, inst.data(), base.data(), key.data(), val.data()); });
  std::vector<long long> order(n_pairs);
  std::iota(order.begin(), order.end(), 0LL);
  std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return key[a] < key[b]; });
  for (long long i = 0; i < n_pairs; i++) { key_sorted[i] = key[order[i]]; val_sorted[i] = val[order[i]]; }
  launch_serial(blocks_for(n_pairs + N, 256), 256, [&] { k2::unpack_pairs(val_sorted.data(), n_pairs, inst_start.data(), N, pair_minus, pair_base); });
  launch_serial(blocks_for(P + 1, 256), 256, [&] { k2::segment_offsets(key_sorted.data(), n_pairs, P, seg_off); });
  std::vector<int> nch(P + 1), is_small(P + 1), is_big(P + 1);
  std::vector<long long> tco64(P + 2), small_pos(P + 2), big_pos(P + 2);
  launch_serial(blocks_for(P + 1, 256), 256, [&] { k2::chunk_counts(seg_off, P, N, nch.data(), is_small.data(), is_big.data()); });
  launch_block(1024, [&] { k2::scan_exclusive(nch.data(), P + 1, tco64.data()); });
  launch_block(1024, [&] { k2::scan_exclusive(is_small.data(), P + 1, small_pos.data()); });
  launch_block(1024, [&] { k2::scan_exclusive(is_big.data(), P + 1, big_pos.data()); });
  out_counts[2] = tco64[P + 1]; out_counts[3] = small_pos[P + 1]; out_counts[4] = big_pos[P + 1];
  if (!chunks4) return;
  launch_serial(blocks_for(P + 2, 256), 256, [&] {
    k2::chunk_fill(seg_off, tco64.data(), small_pos.data(), big_pos.data(), is_small.data(), P, N, n_pairs, reinterpret_cast<Chunk*>(chunks4), tco,
                   small_list, big_list);
  });
  for (long long i = 0; i < n_inst; i++) { inst2[2 * i] = inst[i].person; inst2[2 * i + 1] = inst[i].code; }
  for (int p = 0; p < N; p++) { base2[2 * p] = base[p].person; base2[2 * p + 1] = base[p].code; }
}

// the selection predicate of a restricted sweep, applied to every slot of the instance list
extern "C" int k2_select(const int* inst2, int n_inst, const unsigned char* need, int with_base, int left, int right, int* out_idx) {
  k2::WantInstance w;
  w.inst = reinterpret_cast<const PsyInst*>(inst2); w.need = need; w.with_base = with_base; w.left = left; w.right = right;
  int n = 0;
  for (int i = 0; i < n_inst; i++) if (w(i)) out_idx[n++] = i;
  return n;
}

extern "C" void k2_sets(const double* f, int N, int n_sets, double* out) {
  const int n_chunks = (N + k2::CHUNK_LEN - 1) / k2::CHUNK_LEN;
  std::vector<double> part((size_t)n_sets * n_chunks * 2);
  for (int b = 0; b < n_sets * n_chunks; b++) {
    g_barrier.reset(new std::barrier<>(256));
    std::vector<std::thread> th;
    for (int t = 0; t < 256; t++)
      th.emplace_back([=, &part] { blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)t; blockDim.x = 256; gridDim.x = (unsigned)(n_sets * n_chunks);
                                   k2::sets_chunks(f, N, n_chunks, part.data()); });
    for (auto& x : th) x.join();
  }
  launch_serial(blocks_for(n_sets, 128), 128, [&] { k2::sets_final(part.data(), n_sets, n_chunks, out); });
}
"""


def _lib():
    text = _SRC + "".join(open(os.path.join(d, f)).read() for d, f in
                          ((_CSRC, "psy_k2.cuh"), (_CSRC, "psy_reduce.cuh"), (_CSRC, "psy_launch.h"), (_HERE, "cuda_host_shim.h")))
    so = os.path.join(_BUILD, "k2sim_" + hashlib.sha1(text.encode()).hexdigest()[:16] + ".so")
    if not os.path.exists(so):
        os.makedirs(_BUILD, exist_ok=True)
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I", _HERE, "-I", _CSRC, "-o", so + ".tmp", "-x", "c++", "-"]
        r = subprocess.run(cmd, input=_SRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("k2sim: g++ failed:\n" + r.stderr[:4000])
        os.replace(so + ".tmp", so)
    L = C.CDLL(so)
    L.k2_select.restype = C.c_int
    return L


_ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
_lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))
_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def step_counts(dt, h):
    dt = np.ascontiguousarray(dt, dtype=np.float64)
    out = np.zeros(len(dt), dtype=np.int32)
    _lib().k2_step_counts(_dp(dt), C.c_longlong(len(dt)), C.c_double(h), _ip(out))
    return out


def build(theta_index, P):
    """The K2 kernels on a slot map; returns the same dictionary as k3sim.lists plus the instance list itself."""
    L = _lib()
    tidx = np.ascontiguousarray(theta_index, dtype=np.int32)
    N, F = tidx.shape
    counts = np.zeros(5, dtype=np.int64)
    null = None
    L.k2_build(_ip(tidx), N, F, P, _lp(counts), null, null, null, null, null, null, null, null, null, null)
    n_inst, n_pairs = int(counts[0]), int(counts[1])
    start = np.zeros(N, dtype=np.int64)
    inst = np.zeros((max(n_inst, 1), 2), dtype=np.int32)
    base = np.zeros((N, 2), dtype=np.int32)
    pm, pb = np.zeros(n_pairs + N, dtype=np.int32), np.zeros(n_pairs + N, dtype=np.int32)
    seg_off = np.zeros(P + 1, dtype=np.int64)
    L.k2_build(_ip(tidx), N, F, P, _lp(counts), _lp(start), _ip(inst), _ip(base), _ip(pm), _ip(pb), _lp(seg_off), null, null, null, null)
    chunks = np.zeros((max(int(counts[2]), 1), 4), dtype=np.int32)
    tco = np.zeros(P + 2, dtype=np.int32)
    small, big = np.zeros(max(int(counts[3]), 1), dtype=np.int32), np.zeros(max(int(counts[4]), 1), dtype=np.int32)
    L.k2_build(_ip(tidx), N, F, P, _lp(counts), _lp(start), _ip(inst), _ip(base), _ip(pm), _ip(pb), _lp(seg_off), _ip(chunks), _ip(tco),
               _ip(small), _ip(big))
    return dict(n_inst=n_inst, start=start, inst=inst[:n_inst], base=base, pair_minus=pm, pair_base=pb, seg_off=seg_off,
                chunks=chunks[:int(counts[2])], tco=tco, small=small[:int(counts[3])], big=big[:int(counts[4])])


def select(inst, need, with_base, left, right):
    inst = np.ascontiguousarray(inst, dtype=np.int32)
    out = np.zeros(len(inst), dtype=np.int32)
    nd = None if need is None else np.ascontiguousarray(need, dtype=np.uint8)
    n = _lib().k2_select(_ip(inst), len(inst), None if nd is None else nd.ctypes.data_as(C.POINTER(C.c_ubyte)), int(with_base), int(left),
                         int(right), _ip(out))
    return out[:n]


def sets(f, N, n_sets):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out = np.zeros(2 * n_sets)
    _lib().k2_sets(_dp(f), N, n_sets, _dp(out))
    return out[:n_sets], out[n_sets:]
