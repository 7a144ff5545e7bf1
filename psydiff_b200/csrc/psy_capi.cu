// psy_capi.cu — libpsydiff_b200.so: the device side of the C ABI declared in include/psydiff_b200.h.
//
// Host orchestration around the generated filter kernel (psy_filter.cuh): module load, dataset residency in HBM, and — per
// device — the K2 builders (psy_k2.cuh: instance list, pair lists, chunk lists, RK4 step counts, all made on the device), the
// batched launch that folds getGradients' 2P+1 fitModel sweeps (R/modelTemplate.R:852-930) into one kernel, and K3, the
// deterministic per-parameter reduction (psy_reduce.cuh).
//
// ONE handle drives ALL the devices it was given (psy_model_set_devices): persons are independent units
// (R/modelTemplate.R:244), so the dataset is cut into contiguous person blocks balanced by filter work, every device filters
// its block on its own stream, and the per-parameter partial sums (5P+2 doubles per device) are brought to the first device
// by peer copies over NVLink and added there in device order.  This is what getGradientsParallel's PSOCK workers
// (R/prepareModel.R:572-616) become: one host thread, no process per GPU, no collective on the data path.
// There is NO CPU fallback in this file: without a device every compute entry point returns PSY_ERR_CUDA.
#include "psy_internal.h"
#include "psy_launch.h"
#include "psy_reduce.cuh"
#include "psy_k2.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define fail psy_fail

namespace {

long long g_launches = 0;

#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) return fail(PSY_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define RC_TRY(expr)        \
  do {                      \
    int rc__ = (expr);      \
    if (rc__) return rc__;  \
  } while (0)

// ---- driver entry points, resolved through the runtime so the library carries no link-time libcuda dependency
struct Drv {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};
Drv g_drv;
int g_ndev = 0;

int load_driver() {
  if (g_drv.ok) return PSY_OK;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PSY_ERR_CUDA, "no CUDA device available (%s); libpsydiff_b200 has no CPU fallback", cudaGetErrorString(e));
  CU_TRY(cudaFree(0));
  struct { const char* name; void** slot; } tab[] = {
      {"cuModuleLoadData", (void**)&g_drv.ModuleLoadData}, {"cuModuleUnload", (void**)&g_drv.ModuleUnload},
      {"cuModuleGetFunction", (void**)&g_drv.ModuleGetFunction}, {"cuModuleGetGlobal", (void**)&g_drv.ModuleGetGlobal}, {"cuLaunchKernel", (void**)&g_drv.LaunchKernel},
      {"cuFuncGetAttribute", (void**)&g_drv.FuncGetAttribute}, {"cuFuncSetAttribute", (void**)&g_drv.FuncSetAttribute}, {"cuGetErrorString", (void**)&g_drv.GetErrorString}};
  for (auto& t : tab) {
    cudaDriverEntryPointQueryResult qr;
    cudaError_t ee = cudaGetDriverEntryPoint(t.name, t.slot, cudaEnableDefault, &qr);
    if (ee != cudaSuccess || qr != cudaDriverEntryPointSuccess || !*t.slot)
      return fail(PSY_ERR_CUDA, "cannot resolve driver entry point %s", t.name);
  }
  g_ndev = ndev;
  g_drv.ok = true;
  return PSY_OK;
}

const char* cu_str(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

// the caller's current device is put back when an entry point returns (torch and R code rely on it)
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// device buffer; alloc/upload act on the CURRENT device, so callers select the shard's device first
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n && p) return PSY_OK;
    release();
    if (count == 0) count = 1;
    CU_TRY(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
    return PSY_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// odeint integrate_const step count, rule T1: steps are taken for k = 0, 1, ... while
//   ((0 + k*h) + h) - t1 <= DBL_EPSILON      (exact IEEE double, no FMA; this file's host code is compiled without contraction)
// and the remainder interval is dropped.  The predicate is monotone in k, so the count is found from the guess
// floor(t1/h) by local search instead of walking all k — same integer, O(1) per row.  (The device twin that fills the
// per-row table is k2::step_counts.)
inline bool t1_cond(int k, double h, double t1) {
  volatile double time = 0.0 + (double)k * h;
  volatile double next = time + h;
  return (next - t1) <= DBL_EPSILON;
}
int rk4_step_count(double t1, double h) {
  if (!(h > 0) || !(t1 == t1) || !(t1 > 0)) return 0;
  double q = std::floor(t1 / h);
  if (q > 5e7) q = 5e7;
  int k = (int)q;
  while (k > 0 && !t1_cond(k - 1, h, t1)) --k;
  while (k < 50000000 && t1_cond(k, h, t1)) ++k;
  return k;
}

// row-major [rows][k] -> column-major rows x k (R matrix layout) on the device, coalesced on the write side
__global__ void psy_to_colmajor(const double* __restrict__ src, double* __restrict__ dst, long long rows, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * k) return;
  const long long r = i % rows;
  const int c = (int)(i / rows);
  dst[i] = src[r * k + c];
}

// dependent-free DFMA stream: 8 independent accumulators per thread, the FP64 roofline denominator
__global__ void __launch_bounds__(256) psy_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

inline unsigned blocks_for(long long n, int block) { return (unsigned)std::max<long long>(1, (n + block - 1) / block); }

}  // namespace

// One device's share of the model: a contiguous block of persons, resident in that device's HBM, with its own stream.
struct Shard {
  int dev = 0;
  int p0 = 0, N = 0;             // first person (position in the dataset) and person count
  int64_t r0 = 0, rows = 0;      // first row and row count
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_part = nullptr;
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  DevBuf<double> d_obs, d_dt, d_theta, d_f, d_latent, d_pred, d_chunk_part, d_partial, d_sets_part, d_sets_out;
  DevBuf<int> d_ksteps, d_pid, d_tidx, d_pair_minus, d_pair_base, d_out_idx, d_tco, d_small_list, d_big_list, d_nsel;
  DevBuf<long long> d_row_off, d_seg_off;
  DevBuf<unsigned char> d_need;
  DevBuf<PsyInst> d_inst, d_inst_sub, d_inst_base;
  DevBuf<Chunk> d_chunks;
  DevBuf<char> d_tmp;            // scratch of the K2 builders / cub
  int64_t n_inst = 0, n_inst_real = 0, n_pairs = 0, n_chunks = 0, n_small = 0, n_big = 0;
  double ksteps_h = -1.0;
  bool slots_ready = false;
  int lists_N = -1;           // persons the K2 lists of this shard were built for
  double last_ms = 0.0;
  int64_t last_instances = 0;
  int* h_nsel = nullptr;         // pinned: selected-instance count of a restricted sweep
  void release_buffers() {
    d_obs.release(); d_dt.release(); d_theta.release(); d_f.release(); d_latent.release(); d_pred.release(); d_chunk_part.release();
    d_partial.release(); d_sets_part.release(); d_sets_out.release(); d_ksteps.release(); d_pid.release(); d_tidx.release();
    d_pair_minus.release(); d_pair_base.release(); d_out_idx.release(); d_tco.release(); d_small_list.release(); d_big_list.release();
    d_nsel.release(); d_row_off.release(); d_seg_off.release(); d_need.release(); d_inst.release(); d_inst_sub.release();
    d_inst_base.release(); d_chunks.release(); d_tmp.release();
  }
};

struct psy_model {
  int n = 0, m = 0, F = 0, P = 0, N = 0;
  int64_t rows = 0;
  // settings
  double alpha = .2, beta = 2.0, kappa = 0.0;
  int stepper = PSY_STEPPER_RK4, sr = 1;
  std::vector<double> time_step;
  // module (one image, loaded once per device)
  std::string image;
  int block = 128;                  // = the module's __launch_bounds__ block size
  int smem_bytes = 0;               // dynamic shared memory the kernel asks for (psy_module_info[5])
  int lanes = 1;                    // lanes per instance in the RK4 stage loop (psy_module_lanes; 2 = lane-pair kernel)
  int mod_stepper = 0, mod_sr = 1;  // what the loaded module was generated for (psy_module_info)
  // devices
  std::vector<int> devices;
  std::vector<Shard> sh;
  DevBuf<double> d_gather, d_total;  // on devices[0]: the shards' partial vectors side by side, and their sum
  // host copy of the dataset in the layout the devices hold (pinned; what a change of the device set re-distributes)
  double* stage = nullptr;           // [rows][m] row-major
  size_t stage_n = 0;
  std::vector<double> h_dt;
  std::vector<int64_t> h_row_off;
  std::vector<int> h_pid, h_tidx;
  bool have_data = false, have_slots = false;
  double last_ms = 0.0;
  int64_t last_instances = 0;
  psy_sweep sweep;  // gradient sweep state (psy_grad_resolve)
};

namespace {

int use(const Shard& s) {
  CU_TRY(cudaSetDevice(s.dev));
  return PSY_OK;
}

void destroy_shard(Shard& s) {
  if (cudaSetDevice(s.dev) != cudaSuccess) return;
  s.release_buffers();
  if (s.h_nsel) cudaFreeHost(s.h_nsel);
  if (s.ev0) cudaEventDestroy(s.ev0);
  if (s.ev1) cudaEventDestroy(s.ev1);
  if (s.ev_part) cudaEventDestroy(s.ev_part);
  if (s.st) cudaStreamDestroy(s.st);
  if (s.mod && g_drv.ModuleUnload) g_drv.ModuleUnload(s.mod);
  s = Shard();
}

// load the handle's module image onto the shard's device and read back what it was generated for
int load_module_on(psy_model* M, Shard& s, const char* what) {
  RC_TRY(use(s));
  CU_TRY(cudaFree(0));
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  CUresult r = g_drv.ModuleLoadData(&mod, M->image.data());
  if (r != CUDA_SUCCESS) return fail(PSY_ERR_CUDA, "cuModuleLoadData(%s) on device %d: %s", what, s.dev, cu_str(r));
  r = g_drv.ModuleGetFunction(&fn, mod, "psy_filter_kernel");
  if (r != CUDA_SUCCESS) { g_drv.ModuleUnload(mod); return fail(PSY_ERR_CUDA, "psy_filter_kernel not found in %s: %s", what, cu_str(r)); }
  int info[8] = {0, 1, M->n, M->m, M->F, 0, -1, 0};
  CUdeviceptr gp = 0;
  size_t gs = 0;
  if (g_drv.ModuleGetGlobal(&gp, &gs, mod, "psy_module_info") == CUDA_SUCCESS && gs == sizeof info)
    CU_TRY(cudaMemcpy(info, (const void*)gp, sizeof info, cudaMemcpyDeviceToHost));
  if (info[2] != M->n || info[3] != M->m || info[4] != M->F) {
    g_drv.ModuleUnload(mod);
    return fail(PSY_ERR_ARG, "module %s was generated for n=%d m=%d slots=%d, handle has n=%d m=%d slots=%d", what, info[2],
                info[3], info[4], M->n, M->m, M->F);
  }
  if (s.mod) g_drv.ModuleUnload(s.mod);
  s.mod = mod; s.fn = fn;
  M->mod_stepper = info[0]; M->mod_sr = info[1]; M->smem_bytes = info[5];
  M->lanes = 1;
  if (g_drv.ModuleGetGlobal(&gp, &gs, mod, "psy_module_lanes") == CUDA_SUCCESS && gs == sizeof(int))
    CU_TRY(cudaMemcpy(&M->lanes, (const void*)gp, sizeof(int), cudaMemcpyDeviceToHost));
  if (M->smem_bytes > 48 * 1024) {
    r = g_drv.FuncSetAttribute(s.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, M->smem_bytes);
    if (r != CUDA_SUCCESS) return fail(PSY_ERR_CUDA, "cannot reserve %d bytes of shared memory: %s", M->smem_bytes, cu_str(r));
  }
  // Shared-memory carveout of the unified L1 (percent of the maximum): the module states what it was tuned for
  // (psy_module_info[6]; a kernel whose spill frame must stay L1-resident wants a small carveout, which also bounds the resident
  // blocks per SM); the environment variable PSY_SMEM_CARVEOUT overrides it (experiments).
  int carve = info[6];
  if (const char* cv = getenv("PSY_SMEM_CARVEOUT")) carve = atoi(cv);
  if (carve >= 0 && carve <= 100) g_drv.FuncSetAttribute(s.fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, carve);
  int mt = 0;
  M->block = 128;
  if (g_drv.FuncGetAttribute(&mt, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, s.fn) == CUDA_SUCCESS && mt >= 32 && mt <= 256) M->block = mt;
  return PSY_OK;
}

int init_shard(psy_model* M, Shard& s, int dev) {
  s.dev = dev;
  RC_TRY(use(s));
  CU_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreate(&s.ev0));
  CU_TRY(cudaEventCreate(&s.ev1));
  CU_TRY(cudaEventCreateWithFlags(&s.ev_part, cudaEventDisableTiming));
  CU_TRY(cudaMallocHost((void**)&s.h_nsel, sizeof(int)));
  return load_module_on(M, s, "the model module");
}

// Person blocks per device, cut where the running sum of filter work crosses k/n of the total.  Work of a person =
// Σ_t (1 measurement update + dt_t / h integration steps); h = timeStep[0] when the settings are known, else the package
// default min(dt > 0)/30 (R/prepareModel.R:446-449).  Equal series give equal counts.
std::vector<int> shard_cuts(const psy_model* M, int n_sh) {
  const int N = M->N;
  std::vector<int> cut(n_sh + 1, 0);
  cut[n_sh] = N;
  if (n_sh == 1) return cut;
  double h = M->time_step.empty() ? 0.0 : M->time_step[0];
  if (!(h > 0)) {
    double mn = INFINITY;
    for (double d : M->h_dt) if (d > 0 && d < mn) mn = d;
    h = std::isfinite(mn) ? mn / 30.0 : 1.0;
  }
  std::vector<double> cum(N + 1, 0.0);
  for (int p = 0; p < N; p++) {
    double w = 0.0;
    for (int64_t r = M->h_row_off[p]; r < M->h_row_off[p + 1]; r++) {
      const double d = M->h_dt[r];
      w += 1.0 + ((d > 0 && std::isfinite(d)) ? std::min(d / h, 1e7) : 0.0);
    }
    cum[p + 1] = cum[p] + w;
  }
  for (int k = 1; k < n_sh; k++) {
    const double target = cum[N] * k / n_sh;
    int c = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    if (c > 0 && target - cum[c - 1] < cum[std::min(c, N)] - target) c--;
    c = std::max(c, cut[k - 1] + 1);               // every device gets at least one person
    c = std::min(c, N - (n_sh - k));
    cut[k] = c;
  }
  return cut;
}

// (re)create the shards for the handle's device list and person count; modules are loaded, buffers are empty
int make_shards(psy_model* M) {
  const int want = (int)std::max<size_t>(1, std::min<size_t>(M->devices.size(), M->N > 0 ? (size_t)M->N : 1));
  if ((int)M->sh.size() == want) {
    bool same = true;
    for (int k = 0; k < want; k++) same = same && M->sh[k].dev == M->devices[k];
    if (same) return PSY_OK;
  }
  for (auto& s : M->sh) destroy_shard(s);
  M->sh.assign(want, Shard());
  for (int k = 0; k < want; k++) RC_TRY(init_shard(M, M->sh[k], M->devices[k]));
  // direct NVLink path for the partial-sum copies (ignored where peer access is unavailable: the copy is then staged)
  for (int k = 1; k < want; k++) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, M->devices[0], M->devices[k]) == cudaSuccess && can) {
      cudaSetDevice(M->devices[0]);
      if (cudaDeviceEnablePeerAccess(M->devices[k], 0) != cudaSuccess) cudaGetLastError();
      cudaSetDevice(M->devices[k]);
      if (cudaDeviceEnablePeerAccess(M->devices[0], 0) != cudaSuccess) cudaGetLastError();
    }
  }
  M->d_gather.release();
  M->d_total.release();
  return PSY_OK;
}

int sync_all(psy_model* M) {
  for (auto& s : M->sh) {
    RC_TRY(use(s));
    CU_TRY(cudaStreamSynchronize(s.st));
  }
  return PSY_OK;
}

// copy the staged dataset to the devices: every shard gets its rows, asynchronously on its own stream
int distribute(psy_model* M) {
  RC_TRY(make_shards(M));
  const std::vector<int> cut = shard_cuts(M, (int)M->sh.size());
  const int m = M->m;
  std::vector<std::vector<long long>> ro(M->sh.size());
  for (size_t k = 0; k < M->sh.size(); k++) {
    Shard& s = M->sh[k];
    RC_TRY(use(s));
    s.p0 = cut[k]; s.N = cut[k + 1] - cut[k];
    s.r0 = M->h_row_off[s.p0]; s.rows = M->h_row_off[s.p0 + s.N] - s.r0;
    s.ksteps_h = -1.0;
    // The lists K2 builds depend on the slot map and on WHICH persons a device holds, not on their rows: with one device (all persons,
    // always) a re-upload of the same persons keeps them; with several devices the work-balanced cuts may move, so they are rebuilt.
    const bool keep_lists = M->sh.size() == 1 && s.slots_ready && M->have_slots && s.lists_N == s.N;
    if (!keep_lists) { s.slots_ready = false; s.n_inst = s.n_inst_real = 0; }
    RC_TRY(s.d_obs.alloc((size_t)s.rows * m));
    RC_TRY(s.d_dt.alloc((size_t)s.rows));
    RC_TRY(s.d_row_off.alloc((size_t)s.N + 1));
    RC_TRY(s.d_pid.alloc((size_t)s.N));
    ro[k].resize((size_t)s.N + 1);
    for (int p = 0; p <= s.N; p++) ro[k][p] = (long long)(M->h_row_off[s.p0 + p] - s.r0);
    CU_TRY(cudaMemcpyAsync(s.d_obs.p, M->stage + (size_t)s.r0 * m, sizeof(double) * (size_t)s.rows * m, cudaMemcpyHostToDevice, s.st));
    CU_TRY(cudaMemcpyAsync(s.d_dt.p, M->h_dt.data() + s.r0, sizeof(double) * (size_t)s.rows, cudaMemcpyHostToDevice, s.st));
    CU_TRY(cudaMemcpyAsync(s.d_row_off.p, ro[k].data(), sizeof(long long) * ro[k].size(), cudaMemcpyHostToDevice, s.st));
    CU_TRY(cudaMemcpyAsync(s.d_pid.p, M->h_pid.data() + s.p0, sizeof(int) * (size_t)s.N, cudaMemcpyHostToDevice, s.st));
  }
  return sync_all(M);
}

// K2 on one device: everything indexed by (person, theta) is built from the shard's slice of the slot map, on its stream.
int build_slots_on(psy_model* M, Shard& s) {
  RC_TRY(use(s));
  const int N = s.N, F = M->F, P = M->P;
  cudaStream_t st = s.st;
  RC_TRY(s.d_tidx.alloc((size_t)N * std::max(F, 1)));
  if (F > 0) CU_TRY(cudaMemcpyAsync(s.d_tidx.p, M->h_tidx.data() + (size_t)s.p0 * F, sizeof(int) * (size_t)N * F, cudaMemcpyHostToDevice, st));
  const int n_pchunks = (N + k2::PERSON_CHUNK - 1) / k2::PERSON_CHUNK;
  // scratch carved from one allocation: dl[N*F] dcnt[N] | pair_off[N+1] inst_start[N] delta[chunks*32] chunk_start[chunks] totals[4]
  DevBuf<int> dl, dcnt;
  DevBuf<long long> pair_off, inst_start, delta, chunk_start, totals;
  RC_TRY(dl.alloc((size_t)N * std::max(F, 1)));
  RC_TRY(dcnt.alloc((size_t)N));
  RC_TRY(pair_off.alloc((size_t)N + 1));
  RC_TRY(inst_start.alloc((size_t)N));
  RC_TRY(delta.alloc((size_t)n_pchunks * 32));
  RC_TRY(chunk_start.alloc((size_t)n_pchunks));
  RC_TRY(totals.alloc(4));
  auto cleanup = [&]() { dl.release(); dcnt.release(); pair_off.release(); inst_start.release(); delta.release(); chunk_start.release(); totals.release(); };
  k2::distinct_theta<<<blocks_for(N, 128), 128, 0, st>>>(s.d_tidx.p, N, F, dl.p, dcnt.p);
  k2::layout_tabulate<<<n_pchunks, 32, 0, st>>>(dcnt.p, N, delta.p);
  k2::layout_chunks<<<1, 32, 0, st>>>(delta.p, n_pchunks, chunk_start.p, totals.p);
  k2::layout_persons<<<blocks_for(n_pchunks, 64), 64, 0, st>>>(dcnt.p, N, chunk_start.p, inst_start.p);
  k2::scan_exclusive<<<1, 1024, 0, st>>>(dcnt.p, N, pair_off.p);
  g_launches += 5;
  long long h_tot[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(&h_tot[0], totals.p, sizeof(long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h_tot[1], pair_off.p + N, sizeof(long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cleanup(); return fail(PSY_ERR_CUDA, "instance-list layout on device %d: %s", s.dev, cudaGetErrorString(e)); }
  s.n_inst = h_tot[0];
  s.n_pairs = h_tot[1];
  s.n_inst_real = (int64_t)N + 2 * s.n_pairs;
  if (s.n_inst > 2147483000LL) {
    cleanup();
    return fail(PSY_ERR_ARG, "instance count %lld on device %d exceeds the per-device limit; shard persons over more devices", (long long)s.n_inst, s.dev);
  }
  int rc = PSY_OK;
  DevBuf<int> key, key_sorted;
  DevBuf<unsigned long long> val, val_sorted;
  DevBuf<int> nch, is_small, is_big;
  DevBuf<long long> tco64, small_pos, big_pos;
  auto cleanup2 = [&]() { key.release(); key_sorted.release(); val.release(); val_sorted.release(); nch.release(); is_small.release();
                          is_big.release(); tco64.release(); small_pos.release(); big_pos.release(); cleanup(); };
#define K2_TRY(expr) do { if ((rc = (expr))) { cleanup2(); return rc; } } while (0)
#define K2_CU(expr) do { cudaError_t e2 = (expr); if (e2 != cudaSuccess) { cleanup2(); return fail(PSY_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e2)); } } while (0)
  K2_TRY(s.d_inst.alloc((size_t)s.n_inst));
  K2_TRY(s.d_inst_base.alloc((size_t)N));
  K2_TRY(s.d_pair_minus.alloc((size_t)s.n_pairs + N));
  K2_TRY(s.d_pair_base.alloc((size_t)s.n_pairs + N));
  K2_TRY(key.alloc((size_t)s.n_pairs)); K2_TRY(key_sorted.alloc((size_t)s.n_pairs));
  K2_TRY(val.alloc((size_t)s.n_pairs)); K2_TRY(val_sorted.alloc((size_t)s.n_pairs));
  K2_CU(cudaMemsetAsync(s.d_inst.p, 0xFF, sizeof(PsyInst) * (size_t)s.n_inst, st));  // {-1, -1}: warp-alignment padding slots
  k2::expand<<<blocks_for(N, 128), 128, 0, st>>>(dl.p, dcnt.p, inst_start.p, pair_off.p, N, F, s.d_inst.p, s.d_inst_base.p, key.p, val.p);
  g_launches++;
  // pairs sorted by theta; the radix sort is stable, so persons stay in dataset order inside every theta's segment
  if (s.n_pairs > 0) {
    int end_bit = 1;
    while (end_bit < 31 && (1LL << end_bit) < (long long)std::max(P, 1)) end_bit++;
    size_t tmp_bytes = 0;
    K2_CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key.p, key_sorted.p, val.p, val_sorted.p, (int)s.n_pairs, 0, end_bit, st));
    K2_TRY(s.d_tmp.alloc(tmp_bytes));
    K2_CU(cub::DeviceRadixSort::SortPairs(s.d_tmp.p, tmp_bytes, key.p, key_sorted.p, val.p, val_sorted.p, (int)s.n_pairs, 0, end_bit, st));
    g_launches++;
  }
  k2::unpack_pairs<<<blocks_for(s.n_pairs + N, 256), 256, 0, st>>>(val_sorted.p, s.n_pairs, inst_start.p, N, s.d_pair_minus.p, s.d_pair_base.p);
  K2_TRY(s.d_seg_off.alloc((size_t)P + 1));
  k2::segment_offsets<<<blocks_for(P + 1, 256), 256, 0, st>>>(key_sorted.p, s.n_pairs, P, s.d_seg_off.p);
  K2_TRY(nch.alloc((size_t)P + 1)); K2_TRY(is_small.alloc((size_t)P + 1)); K2_TRY(is_big.alloc((size_t)P + 1));
  K2_TRY(tco64.alloc((size_t)P + 2)); K2_TRY(small_pos.alloc((size_t)P + 2)); K2_TRY(big_pos.alloc((size_t)P + 2));
  k2::chunk_counts<<<blocks_for(P + 1, 256), 256, 0, st>>>(s.d_seg_off.p, P, N, nch.p, is_small.p, is_big.p);
  k2::scan_exclusive<<<1, 1024, 0, st>>>(nch.p, P + 1, tco64.p);
  k2::scan_exclusive<<<1, 1024, 0, st>>>(is_small.p, P + 1, small_pos.p);
  k2::scan_exclusive<<<1, 1024, 0, st>>>(is_big.p, P + 1, big_pos.p);
  g_launches += 6;
  long long h_cnt[3] = {0, 0, 0};
  K2_CU(cudaMemcpyAsync(&h_cnt[0], tco64.p + P + 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
  K2_CU(cudaMemcpyAsync(&h_cnt[1], small_pos.p + P + 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
  K2_CU(cudaMemcpyAsync(&h_cnt[2], big_pos.p + P + 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
  K2_CU(cudaStreamSynchronize(st));
  s.n_chunks = h_cnt[0]; s.n_small = h_cnt[1]; s.n_big = h_cnt[2];
  K2_TRY(s.d_chunks.alloc((size_t)s.n_chunks));
  K2_TRY(s.d_tco.alloc((size_t)P + 2));
  K2_TRY(s.d_small_list.alloc((size_t)s.n_small));
  K2_TRY(s.d_big_list.alloc((size_t)s.n_big));
  k2::chunk_fill<<<blocks_for(P + 2, 256), 256, 0, st>>>(s.d_seg_off.p, tco64.p, small_pos.p, big_pos.p, is_small.p, P, N, s.n_pairs, s.d_chunks.p,
                                                       s.d_tco.p, s.d_small_list.p, s.d_big_list.p);
  g_launches++;
  K2_TRY(s.d_chunk_part.alloc((size_t)s.n_chunks * 5));
  K2_TRY(s.d_partial.alloc(5 * (size_t)P + 2));
  K2_TRY(s.d_f.alloc((size_t)s.n_inst));
  K2_TRY(s.d_theta.alloc((size_t)std::max(P, 1)));
  K2_CU(cudaStreamSynchronize(st));
  K2_CU(cudaGetLastError());
#undef K2_TRY
#undef K2_CU
  cleanup2();
  s.slots_ready = true;
  s.lists_N = N;
  return PSY_OK;
}

int build_slots(psy_model* M) {
  for (auto& s : M->sh) RC_TRY(build_slots_on(M, s));
  RC_TRY(use(M->sh[0]));
  if (M->sh.size() > 1) {
    RC_TRY(M->d_gather.alloc((size_t)M->sh.size() * (5 * (size_t)M->P + 2)));
    RC_TRY(M->d_total.alloc(5 * (size_t)M->P + 2));
  }
  return PSY_OK;
}

int check_ready(const psy_model* M) {
  if (M->sh.empty() || !M->sh[0].fn) return fail(PSY_ERR_STATE, "model module not loaded");
  if (!M->have_data) return fail(PSY_ERR_STATE, "psy_upload has not been called");
  if (!M->have_slots) return fail(PSY_ERR_STATE, "psy_set_slots has not been called");
  if (M->time_step.empty()) return fail(PSY_ERR_STATE, "psy_model_settings has not been called");
  const int want_stepper = (M->stepper == PSY_STEPPER_DOPRI5) ? 1 : 0;  // "default" == rk4 (a NaN state stays NaN under dopri5)
  if (want_stepper != M->mod_stepper || (M->sr ? 1 : 0) != M->mod_sr)
    return fail(PSY_ERR_STATE, "loaded module was generated for stepper=%d srUpdate=%d but the settings ask for stepper=%d srUpdate=%d: "
                "recompile the model (psy_model_compile + psy_model_load_module)", M->mod_stepper, M->mod_sr, want_stepper, M->sr ? 1 : 0);
  return PSY_OK;
}

// enqueue one filter launch on the shard's stream (no host synchronisation); the shard's device is current
int enqueue_filter(psy_model* M, Shard& s, const PsyInst* d_list, const int* d_out_idx, int64_t count, double eps, int skip_update,
                   double* d_latent, double* d_pred) {
  s.last_instances = count;
  if (count <= 0) {
    CU_TRY(cudaEventRecord(s.ev0, s.st));
    CU_TRY(cudaEventRecord(s.ev1, s.st));
    return PSY_OK;
  }
  if (count > 2147483000LL) return fail(PSY_ERR_ARG, "too many instances for one launch (%lld)", (long long)count);
  const double h = M->time_step[0];
  if (s.ksteps_h != h || !s.d_ksteps.p) {
    RC_TRY(s.d_ksteps.alloc((size_t)s.rows));
    k2::step_counts<<<blocks_for(s.rows, 256), 256, 0, s.st>>>(s.d_dt.p, s.rows, h, s.d_ksteps.p);
    g_launches++;
    s.ksteps_h = h;
  }
  double kappa = M->kappa;
  if (M->n + kappa == 0) kappa -= 1;  // R/modelTemplate.R:192-196
  const double lambda = M->alpha * M->alpha * (M->n + kappa) - M->n;
  PsyLaunch L;
  memset(&L, 0, sizeof L);
  L.obs = s.d_obs.p; L.dt = s.d_dt.p; L.ksteps = s.d_ksteps.p; L.row_off = s.d_row_off.p;
  L.person_id = s.d_pid.p; L.theta_index = s.d_tidx.p; L.theta = s.d_theta.p; L.inst = d_list;
  L.f_out = s.d_f.p; L.out_idx = d_out_idx; L.latent = d_latent; L.pred = d_pred;
  L.eps = eps; L.h = h;
  L.c = M->alpha * M->alpha * (M->n + kappa);
  L.sqrtc = sqrt(L.c);
  L.wm0 = lambda / (M->n + lambda);
  L.wm1 = 1 / (2 * (M->n + lambda));
  L.wc0 = lambda / (M->n + lambda) + (1 - M->alpha * M->alpha + M->beta);
  L.wc1 = L.wm1;
  L.n_inst = (int)count;
  L.skip_update = skip_update;
  L.n_theta = M->P;
  const unsigned block = (unsigned)M->block;
  const unsigned grid = (unsigned)((count + block - 1) / block);
  void* args[] = {&L};
  CU_TRY(cudaEventRecord(s.ev0, s.st));
  CUresult r = g_drv.LaunchKernel(s.fn, grid, 1, 1, block, 1, 1, (unsigned)M->smem_bytes, (CUstream)s.st, args, nullptr);
  if (r != CUDA_SUCCESS) return fail(PSY_ERR_CUDA, "cuLaunchKernel(psy_filter_kernel) on device %d: %s", s.dev, cu_str(r));
  CU_TRY(cudaEventRecord(s.ev1, s.st));
  g_launches++;
  return PSY_OK;
}

// wait for every shard's stream; kernel time of the step = the slowest device's (they run concurrently)
int finish_filter(psy_model* M) {
  double worst = 0.0;
  int64_t inst = 0;
  for (auto& s : M->sh) {
    RC_TRY(use(s));
    CU_TRY(cudaStreamSynchronize(s.st));
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    s.last_ms = ms;
    worst = std::max(worst, (double)ms);
    inst += s.last_instances;
    CU_TRY(cudaGetLastError());
  }
  M->last_ms = worst;
  M->last_instances = inst;
  return PSY_OK;
}

int enqueue_reduction(psy_model* M, Shard& s) {
  if (s.n_chunks > 0) {
    psy_reduce_chunks<<<(unsigned)s.n_chunks, 256, 0, s.st>>>(s.d_chunks.p, s.d_pair_minus.p, s.d_pair_base.p, s.d_f.p, M->P, s.d_chunk_part.p);
    g_launches++;
  }
  if (s.n_small > 0) {
    psy_reduce_small<<<blocks_for(s.n_small, 128), 128, 0, s.st>>>(s.d_small_list.p, (int)s.n_small, s.d_seg_off.p, s.d_pair_minus.p,
                                                                   s.d_pair_base.p, s.d_f.p, M->P, s.d_partial.p);
    g_launches++;
  }
  psy_reduce_final<<<blocks_for(s.n_big, 128), 128, 0, s.st>>>(s.d_big_list.p, (int)s.n_big, s.d_tco.p, s.d_chunk_part.p, M->P, s.d_partial.p);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return PSY_OK;
}

}  // namespace

// =================================================================================================================
extern "C" {

int psy_rk4_step_count(double t1, double h) { return rk4_step_count(t1, h); }

int psy_device_count(void) {
  if (load_driver()) return 0;
  return g_ndev;
}

int psy_model_load_module(psy_model* M, const char* cubin_path) {
  if (!M || !cubin_path) return fail(PSY_ERR_ARG, "psy_model_load_module: bad arguments");
  RC_TRY(load_driver());
  DeviceGuard guard;
  std::string image = slurp(cubin_path);
  if (image.empty()) return fail(PSY_ERR_ARG, "cannot read compiled module %s", cubin_path);
  std::string old = M->image;
  M->image.swap(image);
  if (M->sh.empty()) {
    M->sh.assign(1, Shard());
    int rc = init_shard(M, M->sh[0], M->devices[0]);
    if (rc) { destroy_shard(M->sh[0]); M->sh.clear(); M->image.swap(old); return rc; }
    return PSY_OK;
  }
  for (auto& s : M->sh) {
    int rc = load_module_on(M, s, cubin_path);
    if (rc) { M->image.swap(old); return rc; }
  }
  return PSY_OK;
}

psy_model* psy_model_create(int n, int m, int nfree, const char* cubin_path) {
  if (n < 1 || m < 1 || nfree < 0 || !cubin_path) { fail(PSY_ERR_ARG, "psy_model_create: bad arguments"); return nullptr; }
  if (load_driver()) return nullptr;
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) { fail(PSY_ERR_CUDA, "cudaGetDevice failed"); return nullptr; }
  psy_model* M = new psy_model;
  M->n = n; M->m = m; M->F = nfree;
  M->devices.assign(1, cur);
  if (psy_model_load_module(M, cubin_path)) { delete M; return nullptr; }
  return M;
}

void psy_model_destroy(psy_model* M) {
  if (!M) return;
  DeviceGuard guard;
  if (!M->devices.empty() && cudaSetDevice(M->devices[0]) == cudaSuccess) { M->d_gather.release(); M->d_total.release(); }
  for (auto& s : M->sh) destroy_shard(s);
  if (M->stage) cudaFreeHost(M->stage);
  delete M;
}

int psy_model_set_devices(psy_model* M, int n_devices, const int* device_ids) {
  if (!M || n_devices < 0) return fail(PSY_ERR_ARG, "psy_model_set_devices: bad arguments");
  RC_TRY(load_driver());
  DeviceGuard guard;
  std::vector<int> dv;
  if (n_devices == 0) {
    for (int d = 0; d < g_ndev; d++) dv.push_back(d);
  } else if (device_ids) {
    dv.assign(device_ids, device_ids + n_devices);
  } else {
    // the first n_devices devices, starting with the one the handle was created on
    dv.push_back(M->devices[0]);
    for (int d = 0; d < g_ndev && (int)dv.size() < n_devices; d++) if (d != M->devices[0]) dv.push_back(d);
    if ((int)dv.size() < n_devices) return fail(PSY_ERR_ARG, "psy_model_set_devices: %d devices asked for, %d visible", n_devices, g_ndev);
  }
  for (size_t i = 0; i < dv.size(); i++) {
    if (dv[i] < 0 || dv[i] >= g_ndev) return fail(PSY_ERR_ARG, "psy_model_set_devices: device %d does not exist (%d visible)", dv[i], g_ndev);
    for (size_t j = 0; j < i; j++) if (dv[j] == dv[i]) return fail(PSY_ERR_ARG, "psy_model_set_devices: device %d listed twice", dv[i]);
  }
  if (dv == M->devices) return PSY_OK;
  if (!M->devices.empty() && cudaSetDevice(M->devices[0]) == cudaSuccess) { M->d_gather.release(); M->d_total.release(); }
  M->devices = dv;
  if (!M->have_data) {   // nothing resident yet: only the first device's module is needed until psy_upload
    for (auto& s : M->sh) destroy_shard(s);
    M->sh.assign(1, Shard());
    return init_shard(M, M->sh[0], M->devices[0]);
  }
  RC_TRY(distribute(M));
  if (M->have_slots) RC_TRY(build_slots(M));
  return PSY_OK;
}

int psy_model_devices(const psy_model* M, int* device_ids, int cap) {
  if (!M) return 0;
  const int n = (int)(M->have_data ? M->sh.size() : M->devices.size());
  for (int k = 0; k < n && k < cap && device_ids; k++) device_ids[k] = M->have_data ? M->sh[k].dev : M->devices[k];
  return n;
}

int psy_shard_info(const psy_model* M, int shard, int* device, int* first_person, int* n_persons, int64_t* n_instances, double* last_ms) {
  if (!M || shard < 0 || shard >= (int)M->sh.size()) return fail(PSY_ERR_ARG, "psy_shard_info: no such shard");
  const Shard& s = M->sh[shard];
  if (device) *device = s.dev;
  if (first_person) *first_person = s.p0;
  if (n_persons) *n_persons = s.N;
  if (n_instances) *n_instances = s.n_inst_real;
  if (last_ms) *last_ms = s.last_ms;
  return PSY_OK;
}

int psy_model_settings(psy_model* M, double alpha, double beta, double kappa, int stepper, const double* ts, int n_ts, int sr) {
  if (!M || !ts || n_ts < 1) return fail(PSY_ERR_ARG, "psy_model_settings: bad arguments");
  if (stepper < 0 || stepper > 2) return fail(PSY_ERR_ARG, "unknown integrateFunction code %d", stepper);
  M->alpha = alpha; M->beta = beta; M->kappa = kappa; M->stepper = stepper; M->sr = sr;
  M->time_step.assign(ts, ts + n_ts);
  return PSY_OK;
}

int psy_upload(psy_model* M, int64_t rows, int N, const double* obs_cm, const double* dt, const int64_t* person_off, const int* person_id) {
  if (!M || rows < 1 || N < 1 || !obs_cm || !dt || !person_off || !person_id) return fail(PSY_ERR_ARG, "psy_upload: bad arguments");
  if (person_off[0] != 0 || person_off[N] != rows) return fail(PSY_ERR_ARG, "psy_upload: person_off must run from 0 to rows");
  for (int p = 0; p < N; p++)
    if (person_off[p + 1] <= person_off[p]) return fail(PSY_ERR_ARG, "psy_upload: person %d has no rows", p);
  RC_TRY(load_driver());
  DeviceGuard guard;
  const int m = M->m;
  // R's column-major rows x m  ->  row-major [rows][m], so one time point is m contiguous doubles in HBM.
  // The transpose goes into a pinned staging buffer so the H2D copies run at full PCIe/C2C rate on every device at once;
  // the buffer is kept: it is what psy_model_set_devices re-distributes.
  const size_t cnt = (size_t)rows * m;
  if (M->stage_n < cnt) {
    if (M->stage) cudaFreeHost(M->stage);
    M->stage = nullptr; M->stage_n = 0;
    CU_TRY(cudaMallocHost((void**)&M->stage, cnt * sizeof(double)));
    M->stage_n = cnt;
  }
  double* rm = M->stage;
  {
    const int64_t RB = 4096;
    const int64_t nblk = (rows + RB - 1) / RB;
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, nblk / 8}));
    auto work = [&](int t) {
      for (int64_t b = t; b < nblk; b += nth) {
        const int64_t r0 = b * RB, r1 = std::min(rows, r0 + RB);
        for (int j = 0; j < m; j++) {
          const double* col = obs_cm + (size_t)j * rows;
          for (int64_t r = r0; r < r1; r++) rm[(size_t)r * m + j] = col[r];
        }
      }
    };
    if (nth == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nth; t++) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
  }
  M->h_dt.assign(dt, dt + rows);
  M->h_row_off.assign(person_off, person_off + N + 1);
  M->h_pid.assign(person_id, person_id + N);
  if (N != M->N) {  // a dataset with a different person count invalidates the slot map
    M->have_slots = false;
    M->h_tidx.clear();
  }
  M->rows = rows; M->N = N;
  M->have_data = true;
  RC_TRY(distribute(M));
  if (M->have_slots) {   // same persons, new rows: the person blocks may have moved between devices (never with one device)
    bool kept = M->sh.size() == 1 && M->sh[0].slots_ready;
    if (!kept) RC_TRY(build_slots(M));
  }
  return PSY_OK;
}

int psy_set_slots(psy_model* M, int P, const int* tidx) {
  if (!M || P < 0 || (M->F > 0 && !tidx)) return fail(PSY_ERR_ARG, "psy_set_slots: bad arguments");
  if (M->N < 1 || !M->have_data) return fail(PSY_ERR_STATE, "psy_set_slots: call psy_upload first");
  const int N = M->N, F = M->F;
  for (size_t i = 0; i < (size_t)N * F; i++)
    if (tidx[i] < 0 || tidx[i] >= P) return fail(PSY_ERR_ARG, "psy_set_slots: theta_index[%zu] = %d out of range [0,%d)", i, tidx[i], P);
  DeviceGuard guard;
  M->P = P;
  M->h_tidx.assign(tidx, tidx + (size_t)N * F);
  M->have_slots = false;
  RC_TRY(build_slots(M));
  M->have_slots = true;
  return PSY_OK;
}

int psy_fit(psy_model* M, const double* theta, int skip_update, double* m2ll, double* latent_cm, double* pred_cm) {
  if (!M || !m2ll || (M->P > 0 && !theta)) return fail(PSY_ERR_ARG, "psy_fit: bad arguments");
  if (!M->have_slots) return fail(PSY_ERR_STATE, "psy_fit: call psy_upload and psy_set_slots first");
  RC_TRY(check_ready(M));
  DeviceGuard guard;
  // base instances only; results land in d_f[0..N) of every shard.  First every device gets its work, then the results are
  // collected (a copy into the caller's pageable memory blocks the host, so it must not sit between two devices' launches).
  for (auto& s : M->sh) {
    RC_TRY(use(s));
    if (M->P > 0) CU_TRY(cudaMemcpyAsync(s.d_theta.p, theta, sizeof(double) * M->P, cudaMemcpyHostToDevice, s.st));
    double *dl = nullptr, *dp = nullptr;
    if (latent_cm) { RC_TRY(s.d_latent.alloc((size_t)s.rows * M->n * 2)); dl = s.d_latent.p; }
    if (pred_cm) { RC_TRY(s.d_pred.alloc((size_t)s.rows * M->m * 2)); dp = s.d_pred.p; }
    RC_TRY(enqueue_filter(M, s, s.d_inst_base.p, nullptr, s.N, 0.0, skip_update, dl, dp));
    if (latent_cm) {
      const long long tot = (long long)s.rows * M->n;
      psy_to_colmajor<<<blocks_for(tot, 256), 256, 0, s.st>>>(dl, dl + tot, s.rows, M->n);
      g_launches++;
    }
    if (pred_cm) {
      const long long tot = (long long)s.rows * M->m;
      psy_to_colmajor<<<blocks_for(tot, 256), 256, 0, s.st>>>(dp, dp + tot, s.rows, M->m);
      g_launches++;
    }
  }
  for (auto& s : M->sh) {
    RC_TRY(use(s));
    CU_TRY(cudaMemcpyAsync(m2ll + s.p0, s.d_f.p, sizeof(double) * s.N, cudaMemcpyDeviceToHost, s.st));
    // the R matrices are rows x k column-major over ALL persons: every shard owns a band of rows in each column
    if (latent_cm)
      CU_TRY(cudaMemcpy2DAsync(latent_cm + s.r0, sizeof(double) * (size_t)M->rows, s.d_latent.p + (size_t)s.rows * M->n,
                               sizeof(double) * (size_t)s.rows, sizeof(double) * (size_t)s.rows, (size_t)M->n, cudaMemcpyDeviceToHost, s.st));
    if (pred_cm)
      CU_TRY(cudaMemcpy2DAsync(pred_cm + s.r0, sizeof(double) * (size_t)M->rows, s.d_pred.p + (size_t)s.rows * M->m,
                               sizeof(double) * (size_t)s.rows, sizeof(double) * (size_t)s.rows, (size_t)M->m, cudaMemcpyDeviceToHost, s.st));
  }
  return finish_filter(M);
}

// Several parameter vectors in ONE launch per device: N x n_sets unperturbed instances (a line search's candidate steps,
// random restarts, ...).  totals[k] = Σ_p -2LL at thetas[k*P ..] over the persons whose value is not NaN, nonfinite[k] = the
// number of persons whose -2LL is NaN (the reference's callers test anyNA, R/optimWrapper.R:153).  Person sums are formed on
// the device by a fixed tree (4096-person chunks, then chunks in order), devices are added in order: deterministic.
int psy_fit_many(psy_model* M, const double* thetas, int n_sets, int skip_update, double* totals, double* nonfinite) {
  if (!M || !thetas || n_sets < 1 || !totals || !nonfinite) return fail(PSY_ERR_ARG, "psy_fit_many: bad arguments");
  if (!M->have_slots) return fail(PSY_ERR_STATE, "psy_fit_many: call psy_upload and psy_set_slots first");
  RC_TRY(check_ready(M));
  DeviceGuard guard;
  const int P = M->P;
  std::vector<std::vector<double>> out(M->sh.size(), std::vector<double>(2 * (size_t)n_sets));
  for (size_t k = 0; k < M->sh.size(); k++) {
    Shard& s = M->sh[k];
    RC_TRY(use(s));
    const int64_t cnt = (int64_t)s.N * n_sets;
    if (cnt > 2147483000LL) return fail(PSY_ERR_ARG, "psy_fit_many: too many instances");
    RC_TRY(s.d_theta.alloc((size_t)std::max(P, 1) * n_sets));
    if (P > 0) CU_TRY(cudaMemcpyAsync(s.d_theta.p, thetas, sizeof(double) * P * n_sets, cudaMemcpyHostToDevice, s.st));
    RC_TRY(s.d_f.alloc((size_t)std::max<int64_t>(cnt, s.n_inst)));
    RC_TRY(s.d_inst_sub.alloc((size_t)cnt));
    // set-major blocks of persons: lanes of a warp share their person's neighbours' dt pattern as in psy_fit
    k2::many_list<<<blocks_for(cnt, 256), 256, 0, s.st>>>(s.N, n_sets, s.d_inst_sub.p);
    g_launches++;
    RC_TRY(enqueue_filter(M, s, s.d_inst_sub.p, nullptr, cnt, 0.0, skip_update, nullptr, nullptr));
    const int n_chunks = (s.N + k2::CHUNK_LEN - 1) / k2::CHUNK_LEN;
    RC_TRY(s.d_sets_part.alloc((size_t)n_sets * n_chunks * 2));
    RC_TRY(s.d_sets_out.alloc((size_t)n_sets * 2));
    k2::sets_chunks<<<(unsigned)(n_sets * n_chunks), 256, 0, s.st>>>(s.d_f.p, s.N, n_chunks, s.d_sets_part.p);
    k2::sets_final<<<blocks_for(n_sets, 128), 128, 0, s.st>>>(s.d_sets_part.p, n_sets, n_chunks, s.d_sets_out.p);
    g_launches += 2;
  }
  for (size_t k = 0; k < M->sh.size(); k++) {
    Shard& s = M->sh[k];
    RC_TRY(use(s));
    CU_TRY(cudaMemcpyAsync(out[k].data(), s.d_sets_out.p, sizeof(double) * 2 * n_sets, cudaMemcpyDeviceToHost, s.st));
  }
  RC_TRY(finish_filter(M));
  for (int j = 0; j < n_sets; j++) {
    double t = 0.0, b = 0.0;
    for (size_t k = 0; k < M->sh.size(); k++) { t += out[k][j]; b += out[k][n_sets + j]; }
    totals[j] = t;
    nonfinite[j] = b;
  }
  return PSY_OK;
}

int64_t psy_partial_len(const psy_model* M) { return M ? 5 * (int64_t)M->P + 2 : 0; }
int64_t psy_n_instances(const psy_model* M) {
  int64_t n = 0;
  if (M) for (auto& s : M->sh) n += s.n_inst_real;
  return n;
}

int psy_grad_level(psy_model* M, const double* theta, double eps, int direction, const unsigned char* need, void** partial_device) {
  if (!M || (M->P > 0 && !theta)) return fail(PSY_ERR_ARG, "psy_grad_level: bad arguments");
  if (direction < 0 || direction > 2) return fail(PSY_ERR_ARG, "Unknown direction argument. Possible are central, left and right");
  if (!M->have_slots) return fail(PSY_ERR_STATE, "psy_grad_level: call psy_upload and psy_set_slots first");
  RC_TRY(check_ready(M));
  DeviceGuard guard;
  const int P = M->P;
  const size_t plen = 5 * (size_t)P + 2;
  const bool all = (direction == PSY_DIR_CENTRAL) && !need;
  for (auto& s : M->sh) {
    RC_TRY(use(s));
    if (P > 0) CU_TRY(cudaMemcpyAsync(s.d_theta.p, theta, sizeof(double) * P, cudaMemcpyHostToDevice, s.st));
    if (all) {
      RC_TRY(enqueue_filter(M, s, s.d_inst.p, nullptr, s.n_inst, eps, 0, nullptr, nullptr));
    } else {
      // restricted sweep (one-sided direction, or the eps[] fallback for the sides that failed): the slots of the full list
      // that run are selected on the device; results go back to their slots, so K3 reads the same layout
      k2::WantInstance want;
      want.inst = s.d_inst.p;
      want.need = nullptr;
      want.with_base = need ? 0 : 1;   // the first level of a one-sided sweep also needs f0
      want.left = (direction == PSY_DIR_CENTRAL || direction == PSY_DIR_LEFT);
      want.right = (direction == PSY_DIR_CENTRAL || direction == PSY_DIR_RIGHT);
      if (need) {
        RC_TRY(s.d_need.alloc(2 * (size_t)std::max(P, 1)));
        CU_TRY(cudaMemcpyAsync(s.d_need.p, need, 2 * (size_t)P, cudaMemcpyHostToDevice, s.st));
        want.need = s.d_need.p;
      }
      RC_TRY(s.d_out_idx.alloc((size_t)s.n_inst));
      RC_TRY(s.d_inst_sub.alloc((size_t)s.n_inst));
      RC_TRY(s.d_nsel.alloc(1));
      thrust::counting_iterator<int> idx(0);
      size_t tmp_bytes = 0;
      CU_TRY(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, s.d_out_idx.p, s.d_nsel.p, (int)s.n_inst, want, s.st));
      RC_TRY(s.d_tmp.alloc(tmp_bytes));
      CU_TRY(cub::DeviceSelect::If(s.d_tmp.p, tmp_bytes, idx, s.d_out_idx.p, s.d_nsel.p, (int)s.n_inst, want, s.st));
      g_launches++;
      CU_TRY(cudaMemcpyAsync(s.h_nsel, s.d_nsel.p, sizeof(int), cudaMemcpyDeviceToHost, s.st));
      CU_TRY(cudaStreamSynchronize(s.st));
      const int nsel = *s.h_nsel;
      k2::gather_instances<<<blocks_for(nsel, 256), 256, 0, s.st>>>(s.d_inst.p, s.d_out_idx.p, s.d_nsel.p, s.d_inst_sub.p);
      g_launches++;
      if (!need) CU_TRY(cudaMemsetAsync(s.d_f.p, 0, sizeof(double) * s.n_inst, s.st));  // unused sides read as 0 (never consumed)
      RC_TRY(enqueue_filter(M, s, s.d_inst_sub.p, s.d_out_idx.p, nsel, eps, 0, nullptr, nullptr));
    }
    RC_TRY(enqueue_reduction(M, s));
    if (M->sh.size() > 1) {
      // this device's partial sums travel to the first device (peer copy, NVLink where the devices are peers)
      const size_t k = (size_t)(&s - &M->sh[0]);
      CU_TRY(cudaMemcpyPeerAsync(M->d_gather.p + k * plen, M->devices[0], s.d_partial.p, s.dev, sizeof(double) * plen, s.st));
      CU_TRY(cudaEventRecord(s.ev_part, s.st));
    }
  }
  if (M->sh.size() > 1) {
    Shard& s0 = M->sh[0];
    RC_TRY(use(s0));
    for (size_t k = 1; k < M->sh.size(); k++) CU_TRY(cudaStreamWaitEvent(s0.st, M->sh[k].ev_part, 0));
    k2::sum_shards<<<blocks_for((long long)plen, 256), 256, 0, s0.st>>>(M->d_gather.p, (int)M->sh.size(), (long long)plen, M->d_total.p);
    g_launches++;
  }
  RC_TRY(finish_filter(M));
  if (partial_device) *partial_device = (M->sh.size() > 1) ? M->d_total.p : M->sh[0].d_partial.p;
  return PSY_OK;
}

int psy_grad_resolve(psy_model* M, const double* part, double eps, int level, int n_eps, int direction, unsigned char* need, int* done) {
  if (!M) return fail(PSY_ERR_ARG, "psy_grad_resolve: bad arguments");
  M->sweep.P = M->P;
  return psy_sweep_resolve(&M->sweep, part, eps, level, n_eps, direction, need, done);
}

int psy_grad_finish(psy_model* M, int direction, double* grad, double* total) {
  if (!M) return fail(PSY_ERR_ARG, "psy_grad_finish: bad arguments");
  return psy_sweep_finish(&M->sweep, direction, grad, total);
}

static int grad_sweep(psy_model* M, const double* theta, const double* eps, int n_eps, int direction, double* grad, double* total) {
  if (!M || !eps || n_eps < 1 || !grad) return fail(PSY_ERR_ARG, "psy_gradients: bad arguments");
  if (direction < 0 || direction > 2) return fail(PSY_ERR_ARG, "Unknown direction argument. Possible are central, left and right");
  const int P = M->P;
  std::vector<unsigned char> need((size_t)P * 2 + 1, 0);
  std::vector<double> part((size_t)psy_partial_len(M));
  int done = 0, rc;
  for (int level = 0; level < n_eps && !done; level++) {
    void* dptr = nullptr;
    if ((rc = psy_grad_level(M, theta, eps[level], direction, (level == 0) ? nullptr : need.data(), &dptr))) return rc;
    {
      DeviceGuard guard;
      CU_TRY(cudaSetDevice(M->devices[0]));
      CU_TRY(cudaMemcpy(part.data(), dptr, sizeof(double) * part.size(), cudaMemcpyDeviceToHost));
    }
    if ((rc = psy_grad_resolve(M, part.data(), eps[level], level, n_eps, direction, need.data(), &done))) return rc;
  }
  return psy_grad_finish(M, direction, grad, total);
}

int psy_gradients(psy_model* M, const double* theta, const double* eps, int n_eps, int direction, double* grad, double* total) {
  return grad_sweep(M, theta, eps, n_eps, direction, grad, total);
}

int psy_gradient(psy_model* M, const double* theta, int par, const double* eps, int n_eps, int direction, double* grad) {
  if (!M || par < 0 || par >= M->P) return fail(PSY_ERR_ARG, "psy_gradient: parameter index out of range");
  // the full sweep is one launch; a single parameter is a view of it (getGradient is what the reference's PSOCK
  // workers call once per parameter, R/prepareModel.R:611-613 — here every "worker" reads the same batched result)
  std::vector<double> g(M->P);
  int rc = grad_sweep(M, theta, eps, n_eps, direction, g.data(), nullptr);
  if (rc) return rc;
  *grad = g[par];
  return PSY_OK;
}

double psy_last_kernel_ms(const psy_model* M) { return M ? M->last_ms : 0.0; }
int64_t psy_last_kernel_instances(const psy_model* M) { return M ? M->last_instances : 0; }
int64_t psy_launch_count(void) { return g_launches; }
void psy_counters_reset(void) { g_launches = 0; }

int psy_measure_fp64_peak(double* tflops) {
  if (!tflops) return fail(PSY_ERR_ARG, "psy_measure_fp64_peak: null argument");
  RC_TRY(load_driver());
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* out = nullptr;
  CU_TRY(cudaMalloc((void**)&out, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(a, 0);
    psy_dfma_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(b, 0);
    g_launches++;
    CU_TRY(cudaEventSynchronize(b));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double fl = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  *tflops = best;
  return PSY_OK;
}

int psy_kernel_info(const psy_model* M, int* regs, int* local_bytes, int* max_threads) {
  if (!M || M->sh.empty() || !M->sh[0].fn) return fail(PSY_ERR_STATE, "psy_kernel_info: no module loaded");
  DeviceGuard guard;
  CU_TRY(cudaSetDevice(M->sh[0].dev));
  CUfunction fn = M->sh[0].fn;
  int v = 0;
  if (regs) { g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, fn); *regs = v; }
  if (local_bytes) { g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, fn); *local_bytes = v; }
  if (max_threads) { g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, fn); *max_threads = v; }
  return PSY_OK;
}

int psy_kernel_lanes(const psy_model* M) { return M ? M->lanes : 0; }

}  // extern "C"
