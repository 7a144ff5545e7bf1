// psy_capi.cu — libpsydiff_b200.so: the C ABI declared in include/psydiff_b200.h.
//
// Host orchestration around the generated filter kernel (psy_filter.cuh): model compile/load, dataset residency
// in HBM, the integer slot map that replaces psydiff's string-keyed parameterTable (inst/include/psydiffBasic.h),
// the (person x perturbation) instance list that folds getGradients' 2P+1 fitModel sweeps
// (R/modelTemplate.R:852-930) into one launch, and K3 — the deterministic per-parameter reduction.
// There is NO CPU fallback in this file: without a device every compute entry point returns PSY_ERR_CUDA.
#include "../../include/psydiff_b200.h"
#include "psy_launch.h"
#include "psy_reduce.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_err;
long long g_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) return fail(PSY_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
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
  g_drv.ok = true;
  return PSY_OK;
}

const char* cu_str(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

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
  int upload(const T* h, size_t count) {
    int rc = alloc(count);
    if (rc) return rc;
    if (count) CU_TRY(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
    return PSY_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace

// error reporting for the other translation units of the library (psy_codegen.cpp)
int psy_set_error(int code, const char* msg) { return fail(code, "%s", msg); }

// host-only bookkeeping of one finite-difference sweep (the eps[] fallback of R/modelTemplate.R:885-898, 905-919)
struct psy_sweep {
  int P = 0;
  double base_total = 0.0, base_nonfinite = 0.0;
  std::vector<double> Dval;             // [P*2]  Σ_p (f_side − f0)
  std::vector<double> eps_used;         // [P]
  std::vector<unsigned char> resolved;  // [P*2]
  std::vector<unsigned char> wanted;    // [P*2]
};

struct psy_model {
  int n = 0, m = 0, F = 0, P = 0, N = 0;
  int64_t rows = 0;
  // settings
  double alpha = .2, beta = 2.0, kappa = 0.0;
  int stepper = PSY_STEPPER_RK4, sr = 1;
  std::vector<double> time_step;
  // module
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  int block = 128;  // = the module's __launch_bounds__ block size
  int smem_bytes = 0;  // dynamic shared memory the kernel asks for (psy_module_info[5])
  int mod_stepper = 0, mod_sr = 1;  // what the loaded module was generated for (psy_module_info)
  // host copies
  std::vector<double> h_dt;
  std::vector<int64_t> h_row_off;
  std::vector<int> h_tidx;           // N x F
  std::vector<int64_t> h_inst_off;   // N+1 (end of each person's possibly padded block)
  std::vector<int64_t> h_inst_start; // N first instance (= base instance) of each person
  std::vector<int> h_dl;             // distinct theta of each person in slot order, flattened (see psy_set_slots)
  std::vector<int64_t> h_dl_off;     // N+1
  std::vector<int64_t> h_seg_off;    // P+1 (pairs per theta)
  double ksteps_h = -1.0;
  // device
  DevBuf<double> d_obs, d_dt, d_theta, d_f, d_latent, d_pred, d_chunk_part, d_partial;
  DevBuf<int> d_ksteps, d_pid, d_tidx, d_pair_minus, d_pair_base, d_out_idx, d_theta_chunk_off;
  DevBuf<long long> d_row_off, d_seg_off;
  DevBuf<int> d_small_list, d_big_list;
  int64_t n_small = 0, n_big = 0;
  DevBuf<PsyInst> d_inst, d_inst_sub, d_inst_base;
  DevBuf<Chunk> d_chunks;
  int64_t n_inst = 0, n_inst_real = 0, n_chunks = 0;  // n_inst counts padding slots, n_inst_real does not
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0.0;
  int64_t last_instances = 0;
  psy_sweep sweep;  // gradient sweep state (psy_grad_resolve)
  double* stage = nullptr;  // pinned staging buffer for psy_upload
  size_t stage_n = 0;
};

// =================================================================================================================
// K3: deterministic reduction of instance results into per-parameter sums
// =================================================================================================================
namespace {

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

// ---- optional in-process compile through NVRTC (PSY_COMPILER=nvrtc); libnvrtc is dlopen'ed, never linked ------------
struct Nvrtc {
  void* lib = nullptr;
  int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(void*, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(void*, size_t*) = nullptr;
  int (*GetCUBIN)(void*, char*) = nullptr;
  int (*GetProgramLogSize)(void*, size_t*) = nullptr;
  int (*GetProgramLog)(void*, char*) = nullptr;
  int (*DestroyProgram)(void**) = nullptr;
};
Nvrtc g_nvrtc;

bool load_nvrtc() {
  if (g_nvrtc.lib) return true;
  for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
    g_nvrtc.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (g_nvrtc.lib) break;
  }
  if (!g_nvrtc.lib) return false;
  struct { const char* n; void** f; } tab[] = {
      {"nvrtcCreateProgram", (void**)&g_nvrtc.CreateProgram}, {"nvrtcCompileProgram", (void**)&g_nvrtc.CompileProgram},
      {"nvrtcGetCUBINSize", (void**)&g_nvrtc.GetCUBINSize}, {"nvrtcGetCUBIN", (void**)&g_nvrtc.GetCUBIN},
      {"nvrtcGetProgramLogSize", (void**)&g_nvrtc.GetProgramLogSize}, {"nvrtcGetProgramLog", (void**)&g_nvrtc.GetProgramLog},
      {"nvrtcDestroyProgram", (void**)&g_nvrtc.DestroyProgram}};
  for (auto& t : tab) {
    *t.f = dlsym(g_nvrtc.lib, t.n);
    if (!*t.f) { dlclose(g_nvrtc.lib); g_nvrtc.lib = nullptr; return false; }
  }
  return true;
}

// returns 0 and fills `cubin`, or an error code with the compiler log in `log`
int nvrtc_compile(const char* src, const std::string& inc, const std::string& flags, std::string& cubin, std::string& log) {
  if (!load_nvrtc()) { log = "libnvrtc not found"; return 1; }
  void* prog = nullptr;
  if (g_nvrtc.CreateProgram(&prog, src, "psy_model.cu", 0, nullptr, nullptr) != 0) { log = "nvrtcCreateProgram failed"; return 1; }
  std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--include-path=" + inc};
  std::stringstream ss(flags);
  for (std::string tok; ss >> tok;) opts.push_back(tok);
  std::vector<const char*> o;
  for (auto& x : opts) o.push_back(x.c_str());
  const int rc = g_nvrtc.CompileProgram(prog, (int)o.size(), o.data());
  size_t ls = 0;
  g_nvrtc.GetProgramLogSize(prog, &ls);
  if (ls > 1) { log.resize(ls); g_nvrtc.GetProgramLog(prog, &log[0]); }
  if (rc == 0) {
    size_t cs = 0;
    g_nvrtc.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    g_nvrtc.GetCUBIN(prog, &cubin[0]);
  }
  g_nvrtc.DestroyProgram(&prog);
  return rc;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ULL) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && st.st_size > 0; }

// odeint integrate_const step count, rule T1: steps are taken for k = 0, 1, ... while
//   ((0 + k*h) + h) - t1 <= DBL_EPSILON      (exact IEEE double, no FMA; this file's host code is compiled without contraction)
// and the remainder interval is dropped.  The predicate is monotone in k, so the count is found from the guess
// floor(t1/h) by local search instead of walking all k — same integer, O(1) per row.
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

int ensure_ksteps(psy_model* M) {
  const double h = M->time_step.empty() ? 0.0 : M->time_step[0];
  if (M->ksteps_h == h && M->d_ksteps.p) return PSY_OK;
  std::vector<int> ks(M->rows);
  {
    // rows are independent: split them over the host cores (12.5 M rows of C4 per GPU take ~0.2 s on one core)
    const int64_t rows = M->rows;
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, rows / 65536}));
    auto work = [&](int t) {
      const int64_t r0 = rows * t / nth, r1 = rows * (t + 1) / nth;
      for (int64_t r = r0; r < r1; r++) ks[r] = rk4_step_count(M->h_dt[r], h);
    };
    if (nth == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nth; t++) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
  }
  int rc = M->d_ksteps.upload(ks.data(), ks.size());
  if (rc) return rc;
  M->ksteps_h = h;
  return PSY_OK;
}

int launch_filter(psy_model* M, const PsyInst* d_list, const int* d_out_idx, int64_t count, double eps, int skip_update,
                  double* d_latent, double* d_pred) {
  if (!M->fn) return fail(PSY_ERR_STATE, "model module not loaded");
  if (!M->d_obs.p) return fail(PSY_ERR_STATE, "psy_upload has not been called");
  if (!M->d_tidx.p) return fail(PSY_ERR_STATE, "psy_set_slots has not been called");
  if (M->time_step.empty()) return fail(PSY_ERR_STATE, "psy_model_settings has not been called");
  const int want_stepper = (M->stepper == PSY_STEPPER_DOPRI5) ? 1 : 0;  // "default" == rk4 (a NaN state stays NaN under dopri5)
  if (want_stepper != M->mod_stepper || (M->sr ? 1 : 0) != M->mod_sr)
    return fail(PSY_ERR_STATE, "loaded module was generated for stepper=%d srUpdate=%d but the settings ask for stepper=%d srUpdate=%d: "
                "recompile the model (psy_model_compile + psy_model_load_module)", M->mod_stepper, M->mod_sr, want_stepper, M->sr ? 1 : 0);
  if (count <= 0) return PSY_OK;
  if (count > 2147483000LL) return fail(PSY_ERR_ARG, "too many instances for one launch (%lld)", (long long)count);
  int rc = ensure_ksteps(M);
  if (rc) return rc;
  double kappa = M->kappa;
  if (M->n + kappa == 0) kappa -= 1;  // R/modelTemplate.R:192-196
  const double lambda = M->alpha * M->alpha * (M->n + kappa) - M->n;
  PsyLaunch L;
  memset(&L, 0, sizeof L);
  L.obs = M->d_obs.p; L.dt = M->d_dt.p; L.ksteps = M->d_ksteps.p; L.row_off = M->d_row_off.p;
  L.person_id = M->d_pid.p; L.theta_index = M->d_tidx.p; L.theta = M->d_theta.p; L.inst = d_list;
  L.f_out = M->d_f.p; L.out_idx = d_out_idx; L.latent = d_latent; L.pred = d_pred;
  L.eps = eps; L.h = M->time_step[0];
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
  CU_TRY(cudaEventRecord(M->ev0, 0));
  CUresult r = g_drv.LaunchKernel(M->fn, grid, 1, 1, block, 1, 1, (unsigned)M->smem_bytes, (CUstream)0, args, nullptr);
  if (r != CUDA_SUCCESS) return fail(PSY_ERR_CUDA, "cuLaunchKernel(psy_filter_kernel): %s", cu_str(r));
  CU_TRY(cudaEventRecord(M->ev1, 0));
  g_launches++;
  CU_TRY(cudaEventSynchronize(M->ev1));
  float ms = 0;
  CU_TRY(cudaEventElapsedTime(&ms, M->ev0, M->ev1));
  M->last_ms = ms;
  M->last_instances = count;
  CU_TRY(cudaGetLastError());
  return PSY_OK;
}

int run_reduction(psy_model* M) {
  if (M->n_chunks > 0) {
    psy_reduce_chunks<<<(unsigned)M->n_chunks, 256>>>(M->d_chunks.p, M->d_pair_minus.p, M->d_pair_base.p, M->d_f.p, M->P,
                                                      M->d_chunk_part.p);
    g_launches++;
  }
  if (M->n_small > 0) {
    psy_reduce_small<<<(unsigned)((M->n_small + 127) / 128), 128>>>(M->d_small_list.p, (int)M->n_small, M->d_seg_off.p, M->d_pair_minus.p,
                                                                    M->d_pair_base.p, M->d_f.p, M->P, M->d_partial.p);
    g_launches++;
  }
  psy_reduce_final<<<(unsigned)((M->n_big + 127) / 128), 128>>>(M->d_big_list.p, (int)M->n_big, M->d_theta_chunk_off.p, M->d_chunk_part.p,
                                                                M->P, M->d_partial.p);
  g_launches++;
  CU_TRY(cudaGetLastError());
  return PSY_OK;
}

}  // namespace

// =================================================================================================================
extern "C" {

const char* psy_last_error(void) { return g_err.c_str(); }
int psy_version(void) { return 100; }
int psy_rk4_step_count(double t1, double h) { return rk4_step_count(t1, h); }

int psy_model_compile(const char* cuda_source, const char* include_dir, const char* cache_dir, const char* extra_flags,
                      char* cubin_path, size_t cap) {
  if (!cuda_source || !include_dir || !cache_dir || !cubin_path) return fail(PSY_ERR_ARG, "psy_model_compile: null argument");
  const std::string inc(include_dir), cache(cache_dir), flags(extra_flags ? extra_flags : "");
  uint64_t h = fnv1a(cuda_source);
  for (const char* hdr : {"/psy_filter.cuh", "/psy_lang.cuh", "/psy_launch.h"}) {
    std::string body = slurp(inc + hdr);
    if (body.empty()) return fail(PSY_ERR_COMPILE, "kernel header %s%s not found", include_dir, hdr);
    h = fnv1a(body, h);
  }
  h = fnv1a(flags, h);
  const char* comp_env = getenv("PSY_COMPILER");
  const bool use_nvrtc = comp_env && std::string(comp_env) == "nvrtc";
  if (use_nvrtc) h = fnv1a("nvrtc", h);
  char name[64];
  snprintf(name, sizeof name, "psy_%016llx", (unsigned long long)h);
  mkdir(cache.c_str(), 0755);
  const std::string cu = cache + "/" + name + ".cu", cubin = cache + "/" + name + ".cubin", log = cache + "/" + name + ".log";
  if (cubin.size() + 1 > cap) return fail(PSY_ERR_ARG, "cubin_path buffer too small");
  if (!file_exists(cubin)) {
    {
      std::ofstream f(cu);
      f << cuda_source;
      if (!f) return fail(PSY_ERR_COMPILE, "cannot write %s", cu.c_str());
    }
    if (use_nvrtc) {
      std::string image, clog;
      if (nvrtc_compile(cuda_source, inc, flags, image, clog) != 0) {
        if (clog.size() > 3000) clog = clog.substr(0, 3000) + "\n...";
        return fail(PSY_ERR_COMPILE,
                    "model does not compile for the device (statements outside the supported equation subset are "
                    "unsupported on device):\n%s", clog.c_str());
      }
      const std::string tmpn = cubin + ".tmp" + std::to_string((long long)getpid());
      {
        std::ofstream f(tmpn, std::ios::binary);
        f.write(image.data(), (std::streamsize)image.size());
        if (!f) return fail(PSY_ERR_COMPILE, "cannot write %s", tmpn.c_str());
      }
      if (rename(tmpn.c_str(), cubin.c_str()) != 0) return fail(PSY_ERR_COMPILE, "cannot move %s into the cache", tmpn.c_str());
      strcpy(cubin_path, cubin.c_str());
      return PSY_OK;
    }
    const char* nvcc_env = getenv("PSY_NVCC");
    std::string nvcc = nvcc_env ? nvcc_env : (file_exists("/usr/local/cuda/bin/nvcc") ? "/usr/local/cuda/bin/nvcc" : "nvcc");
    const std::string tmp = cubin + ".tmp" + std::to_string((long long)getpid());
    std::string cmd = nvcc + " -cubin -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -I" + inc + " " +
                      flags + " -o " + tmp + " " + cu + " > " + log + " 2>&1";
    int rc = system(cmd.c_str());
    if (rc != 0 || !file_exists(tmp)) {
      std::string msg = slurp(log);
      if (msg.size() > 3000) msg = msg.substr(0, 3000) + "\n...";
      remove(tmp.c_str());
      return fail(PSY_ERR_COMPILE,
                  "model does not compile for the device (statements outside the supported equation subset are "
                  "unsupported on device):\n%s", msg.c_str());
    }
    if (rename(tmp.c_str(), cubin.c_str()) != 0) return fail(PSY_ERR_COMPILE, "cannot move %s into the cache", tmp.c_str());
  }
  strcpy(cubin_path, cubin.c_str());
  return PSY_OK;
}

int psy_model_load_module(psy_model* M, const char* cubin_path) {
  if (!M || !cubin_path) return fail(PSY_ERR_ARG, "psy_model_load_module: bad arguments");
  int rc = load_driver();
  if (rc) return rc;
  std::string image = slurp(cubin_path);
  if (image.empty()) return fail(PSY_ERR_ARG, "cannot read compiled module %s", cubin_path);
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  CUresult r = g_drv.ModuleLoadData(&mod, image.data());
  if (r != CUDA_SUCCESS) return fail(PSY_ERR_CUDA, "cuModuleLoadData(%s): %s", cubin_path, cu_str(r));
  r = g_drv.ModuleGetFunction(&fn, mod, "psy_filter_kernel");
  if (r != CUDA_SUCCESS) { g_drv.ModuleUnload(mod); return fail(PSY_ERR_CUDA, "psy_filter_kernel not found in %s: %s", cubin_path, cu_str(r)); }
  int info[6] = {0, 1, M->n, M->m, M->F, 0};
  CUdeviceptr gp = 0;
  size_t gs = 0;
  if (g_drv.ModuleGetGlobal(&gp, &gs, mod, "psy_module_info") == CUDA_SUCCESS && gs == sizeof info)
    CU_TRY(cudaMemcpy(info, (const void*)gp, sizeof info, cudaMemcpyDeviceToHost));
  if (info[2] != M->n || info[3] != M->m || info[4] != M->F) {
    g_drv.ModuleUnload(mod);
    return fail(PSY_ERR_ARG, "module %s was generated for n=%d m=%d slots=%d, handle has n=%d m=%d slots=%d", cubin_path, info[2],
                info[3], info[4], M->n, M->m, M->F);
  }
  if (M->mod) g_drv.ModuleUnload(M->mod);
  M->mod = mod; M->fn = fn; M->mod_stepper = info[0]; M->mod_sr = info[1]; M->smem_bytes = info[5];
  if (M->smem_bytes > 48 * 1024) {
    r = g_drv.FuncSetAttribute(M->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, M->smem_bytes);
    if (r != CUDA_SUCCESS) return fail(PSY_ERR_CUDA, "cannot reserve %d bytes of shared memory: %s", M->smem_bytes, cu_str(r));
  }
  int mt = 0;
  M->block = 128;
  if (g_drv.FuncGetAttribute(&mt, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, M->fn) == CUDA_SUCCESS && mt >= 32 && mt <= 256) M->block = mt;
  return PSY_OK;
}

psy_model* psy_model_create(int n, int m, int nfree, const char* cubin_path) {
  if (n < 1 || m < 1 || nfree < 0 || !cubin_path) { fail(PSY_ERR_ARG, "psy_model_create: bad arguments"); return nullptr; }
  if (load_driver()) return nullptr;
  psy_model* M = new psy_model;
  M->n = n; M->m = m; M->F = nfree;
  if (psy_model_load_module(M, cubin_path)) { delete M; return nullptr; }
  cudaEventCreate(&M->ev0);
  cudaEventCreate(&M->ev1);
  return M;
}

void psy_model_destroy(psy_model* M) {
  if (!M) return;
  M->d_obs.release(); M->d_dt.release(); M->d_theta.release(); M->d_f.release(); M->d_latent.release(); M->d_pred.release();
  M->d_chunk_part.release(); M->d_partial.release(); M->d_ksteps.release(); M->d_pid.release(); M->d_tidx.release();
  M->d_pair_minus.release(); M->d_pair_base.release(); M->d_small_list.release(); M->d_big_list.release(); M->d_seg_off.release(); M->d_out_idx.release(); M->d_theta_chunk_off.release();
  M->d_row_off.release(); M->d_inst.release(); M->d_inst_sub.release(); M->d_inst_base.release(); M->d_chunks.release();
  if (M->stage) cudaFreeHost(M->stage);
  if (M->ev0) cudaEventDestroy(M->ev0);
  if (M->ev1) cudaEventDestroy(M->ev1);
  if (M->mod && g_drv.ModuleUnload) g_drv.ModuleUnload(M->mod);
  delete M;
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
  const int m = M->m;
  // R's column-major rows x m  ->  row-major [rows][m], so one time point is m contiguous doubles in HBM.
  // The transpose goes through a pinned staging buffer so the H2D copy runs at full PCIe/C2C rate.
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
  int rc;
  if ((rc = M->d_obs.upload(rm, cnt))) return rc;
  if ((rc = M->d_dt.upload(dt, rows))) return rc;
  std::vector<long long> ro(person_off, person_off + N + 1);
  if ((rc = M->d_row_off.upload(ro.data(), ro.size()))) return rc;
  if ((rc = M->d_pid.upload(person_id, N))) return rc;
  M->h_dt.assign(dt, dt + rows);
  M->h_row_off.assign(person_off, person_off + N + 1);
  if (N != M->N) {  // a dataset with a different person count invalidates the slot map
    M->d_tidx.release();
    M->n_inst = 0;
  }
  M->rows = rows; M->N = N; M->ksteps_h = -1.0;
  return PSY_OK;
}

int psy_set_slots(psy_model* M, int P, const int* tidx) {
  if (!M || P < 0 || (M->F > 0 && !tidx)) return fail(PSY_ERR_ARG, "psy_set_slots: bad arguments");
  if (M->N < 1) return fail(PSY_ERR_STATE, "psy_set_slots: call psy_upload first");
  const int N = M->N, F = M->F;
  for (size_t i = 0; i < (size_t)N * F; i++)
    if (tidx[i] < 0 || tidx[i] >= P) return fail(PSY_ERR_ARG, "psy_set_slots: theta_index[%zu] = %d out of range [0,%d)", i, tidx[i], P);
  M->P = P;
  M->h_tidx.assign(tidx, tidx + (size_t)N * F);
  // distinct theta per person, in slot order
  M->h_dl.clear(); M->h_dl_off.assign(N + 1, 0); M->h_inst_off.assign(N + 1, 0); M->h_inst_start.assign(N, 0);
  int64_t n_real = 0;
  std::vector<int64_t> cnt(P + 1, 0);
  for (int p = 0; p < N; p++) {
    const int* row = tidx + (size_t)p * F;
    size_t b = M->h_dl.size();
    for (int k = 0; k < F; k++) {
      bool seen = false;
      for (size_t q = b; q < M->h_dl.size(); q++) if (M->h_dl[q] == row[k]) { seen = true; break; }
      if (!seen) { M->h_dl.push_back(row[k]); cnt[row[k]]++; }
    }
    M->h_dl_off[p + 1] = (int64_t)M->h_dl.size();
    // Warp alignment: lanes of one warp then share their person's dt sequence and missingness pattern, so the RK4 step
    // loop does not diverge.  A person's block is padded to a multiple of 32 only when that idles < 10 % of its lanes.
    const int64_t cnt_p = 1 + 2 * (int64_t)(M->h_dl.size() - b);
    const int64_t padded = (cnt_p + 31) / 32 * 32;
    int64_t start = M->h_inst_off[p];
    const bool align = (padded - cnt_p) * 10 <= cnt_p;
    if (align) start = (start + 31) / 32 * 32;
    M->h_inst_start[p] = start;
    M->h_inst_off[p + 1] = start + (align ? padded : cnt_p);
    n_real += cnt_p;
  }
  M->n_inst = M->h_inst_off[N];
  M->n_inst_real = n_real;
  if (M->n_inst > 2147483000LL) return fail(PSY_ERR_ARG, "instance count %lld exceeds the per-rank limit; shard persons over more ranks", (long long)M->n_inst);
  std::vector<PsyInst> inst((size_t)M->n_inst, PsyInst{-1, -1}), base((size_t)N);
  M->h_seg_off.assign(P + 1, 0);
  for (int t = 0; t < P; t++) M->h_seg_off[t + 1] = M->h_seg_off[t] + cnt[t];
  const int64_t n_pairs = M->h_seg_off[P];
  std::vector<int> pair_minus((size_t)n_pairs + N), pair_base((size_t)n_pairs + N);
  std::vector<int64_t> fill(M->h_seg_off.begin(), M->h_seg_off.end() - 1);
  for (int p = 0; p < N; p++) {
    int64_t o = M->h_inst_start[p];
    inst[o] = PsyInst{p, -1};
    base[p] = PsyInst{p, -1};
    int64_t j = 0;
    for (int64_t q = M->h_dl_off[p]; q < M->h_dl_off[p + 1]; q++, j++) {
      const int t = M->h_dl[q];
      inst[o + 1 + 2 * j] = PsyInst{p, (t << 1) | 0};
      inst[o + 2 + 2 * j] = PsyInst{p, (t << 1) | 1};
      pair_minus[fill[t]] = (int)(o + 1 + 2 * j);
      pair_base[fill[t]] = (int)o;
      fill[t]++;
    }
    pair_minus[n_pairs + p] = (int)o;  // base pseudo-segment (theta == P)
    pair_base[n_pairs + p] = (int)o;
  }
  // chunk list
  const int CH = 4096, SMALL_SEG = 64;
  std::vector<Chunk> chunks;
  std::vector<int> tco(P + 2, 0), small_list, big_list;
  for (int t = 0; t <= P; t++) {
    const int64_t s0 = (t < P) ? M->h_seg_off[t] : n_pairs, s1 = (t < P) ? M->h_seg_off[t + 1] : n_pairs + N;
    tco[t] = (int)chunks.size();
    if (t < P && s1 - s0 <= SMALL_SEG) { small_list.push_back(t); continue; }
    big_list.push_back(t);
    for (int64_t s = s0; s < s1; s += CH) chunks.push_back(Chunk{t, (int)s, (int)std::min<int64_t>(CH, s1 - s), 0});
  }
  tco[P + 1] = (int)chunks.size();
  M->n_chunks = (int64_t)chunks.size();
  M->n_small = (int64_t)small_list.size();
  M->n_big = (int64_t)big_list.size();
  std::vector<long long> seg_ll(M->h_seg_off.begin(), M->h_seg_off.end());
  int rc;
  if ((rc = M->d_tidx.upload(M->h_tidx.data(), M->h_tidx.size()))) return rc;
  if ((rc = M->d_inst.upload(inst.data(), inst.size()))) return rc;
  if ((rc = M->d_inst_base.upload(base.data(), base.size()))) return rc;
  if ((rc = M->d_pair_minus.upload(pair_minus.data(), pair_minus.size()))) return rc;
  if ((rc = M->d_pair_base.upload(pair_base.data(), pair_base.size()))) return rc;
  if ((rc = M->d_chunks.upload(chunks.data(), chunks.size()))) return rc;
  if ((rc = M->d_theta_chunk_off.upload(tco.data(), tco.size()))) return rc;
  if ((rc = M->d_small_list.upload(small_list.data(), small_list.size()))) return rc;
  if ((rc = M->d_big_list.upload(big_list.data(), big_list.size()))) return rc;
  if ((rc = M->d_seg_off.upload(seg_ll.data(), seg_ll.size()))) return rc;
  if ((rc = M->d_chunk_part.alloc(chunks.size() * 5))) return rc;
  if ((rc = M->d_partial.alloc(5 * (size_t)P + 2))) return rc;
  if ((rc = M->d_f.alloc((size_t)M->n_inst))) return rc;
  if ((rc = M->d_theta.alloc((size_t)std::max(P, 1)))) return rc;
  return PSY_OK;
}

int psy_fit(psy_model* M, const double* theta, int skip_update, double* m2ll, double* latent_cm, double* pred_cm) {
  if (!M || !m2ll || (M->P > 0 && !theta)) return fail(PSY_ERR_ARG, "psy_fit: bad arguments");
  if (!M->d_tidx.p) return fail(PSY_ERR_STATE, "psy_fit: call psy_upload and psy_set_slots first");
  int rc;
  if (M->P > 0) CU_TRY(cudaMemcpy(M->d_theta.p, theta, sizeof(double) * M->P, cudaMemcpyHostToDevice));
  double *dl = nullptr, *dp = nullptr;
  if (latent_cm) { if ((rc = M->d_latent.alloc((size_t)M->rows * M->n * 2))) return rc; dl = M->d_latent.p; }
  if (pred_cm) { if ((rc = M->d_pred.alloc((size_t)M->rows * M->m * 2))) return rc; dp = M->d_pred.p; }
  // base instances only; results land in d_f[0..N)
  if ((rc = launch_filter(M, M->d_inst_base.p, nullptr, M->N, 0.0, skip_update, dl, dp))) return rc;
  CU_TRY(cudaMemcpy(m2ll, M->d_f.p, sizeof(double) * M->N, cudaMemcpyDeviceToHost));
  if (latent_cm) {
    const long long tot = (long long)M->rows * M->n;
    psy_to_colmajor<<<(unsigned)((tot + 255) / 256), 256>>>(dl, dl + tot, M->rows, M->n);
    g_launches++;
    CU_TRY(cudaMemcpy(latent_cm, dl + tot, sizeof(double) * tot, cudaMemcpyDeviceToHost));
  }
  if (pred_cm) {
    const long long tot = (long long)M->rows * M->m;
    psy_to_colmajor<<<(unsigned)((tot + 255) / 256), 256>>>(dp, dp + tot, M->rows, M->m);
    g_launches++;
    CU_TRY(cudaMemcpy(pred_cm, dp + tot, sizeof(double) * tot, cudaMemcpyDeviceToHost));
  }
  return PSY_OK;
}

// Several parameter vectors in ONE launch: N x n_sets unperturbed instances (a line search's candidate steps, random
// restarts, ...).  totals[k] = Σ_p -2LL at thetas[k*P ..], nonfinite[k] = number of persons whose -2LL is not finite.
// Person sums are formed on the host in person order (deterministic).
int psy_fit_many(psy_model* M, const double* thetas, int n_sets, int skip_update, double* totals, double* nonfinite) {
  if (!M || !thetas || n_sets < 1 || !totals || !nonfinite) return fail(PSY_ERR_ARG, "psy_fit_many: bad arguments");
  if (!M->d_tidx.p) return fail(PSY_ERR_STATE, "psy_fit_many: call psy_upload and psy_set_slots first");
  const int N = M->N, P = M->P;
  const int64_t cnt = (int64_t)N * n_sets;
  if (cnt > 2147483000LL) return fail(PSY_ERR_ARG, "psy_fit_many: too many instances");
  int rc;
  if ((rc = M->d_theta.alloc((size_t)std::max(P, 1) * n_sets))) return rc;
  if (P > 0) CU_TRY(cudaMemcpy(M->d_theta.p, thetas, sizeof(double) * P * n_sets, cudaMemcpyHostToDevice));
  if ((rc = M->d_f.alloc((size_t)std::max<int64_t>(cnt, M->n_inst)))) return rc;
  // set-major blocks of persons: lanes of a warp share their person's neighbours' dt pattern as in psy_fit
  std::vector<PsyInst> list((size_t)cnt);
  for (int k = 0; k < n_sets; k++)
    for (int p = 0; p < N; p++) list[(size_t)k * N + p] = PsyInst{p, -1 - k};
  if ((rc = M->d_inst_sub.upload(list.data(), list.size()))) return rc;
  if ((rc = launch_filter(M, M->d_inst_sub.p, nullptr, cnt, 0.0, skip_update, nullptr, nullptr))) return rc;
  std::vector<double> f((size_t)cnt);
  CU_TRY(cudaMemcpy(f.data(), M->d_f.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
  for (int k = 0; k < n_sets; k++) {
    double s = 0.0, bad = 0.0;
    for (int p = 0; p < N; p++) {
      const double v = f[(size_t)k * N + p];
      if (std::isfinite(v)) s += v; else bad += 1.0;
    }
    totals[k] = s;
    nonfinite[k] = bad;
  }
  return PSY_OK;
}

int64_t psy_partial_len(const psy_model* M) { return M ? 5 * (int64_t)M->P + 2 : 0; }
int64_t psy_n_instances(const psy_model* M) { return M ? M->n_inst_real : 0; }

int psy_grad_level(psy_model* M, const double* theta, double eps, int direction, const unsigned char* need, void** partial_device) {
  if (!M || (M->P > 0 && !theta)) return fail(PSY_ERR_ARG, "psy_grad_level: bad arguments");
  if (direction < 0 || direction > 2) return fail(PSY_ERR_ARG, "Unknown direction argument. Possible are central, left and right");
  if (!M->d_tidx.p) return fail(PSY_ERR_STATE, "psy_grad_level: call psy_upload and psy_set_slots first");
  int rc;
  if (M->P > 0) CU_TRY(cudaMemcpy(M->d_theta.p, theta, sizeof(double) * M->P, cudaMemcpyHostToDevice));
  const bool all = (direction == PSY_DIR_CENTRAL) && !need;
  if (all) {
    if ((rc = launch_filter(M, M->d_inst.p, nullptr, M->n_inst, eps, 0, nullptr, nullptr))) return rc;
  } else {
    // restricted sweep (one-sided direction, or the eps[] fallback for the sides that failed): build the sub-list
    std::vector<PsyInst> sub;
    std::vector<int> oidx;
    const bool with_base = !need;  // the first level of a one-sided sweep also needs f0
    for (int p = 0; p < M->N; p++) {
      const int64_t o = M->h_inst_start[p];
      if (with_base) { sub.push_back(PsyInst{p, -1}); oidx.push_back((int)o); }
      int64_t j = 0;
      for (int64_t q = M->h_dl_off[p]; q < M->h_dl_off[p + 1]; q++, j++) {
        const int t = M->h_dl[q];
        for (int side = 0; side < 2; side++) {
          bool want = need ? need[2 * t + side] != 0
                           : (direction == PSY_DIR_CENTRAL || (direction == PSY_DIR_LEFT && side == 0) || (direction == PSY_DIR_RIGHT && side == 1));
          if (!want) continue;
          sub.push_back(PsyInst{p, (t << 1) | side});
          oidx.push_back((int)(o + 1 + 2 * j + side));
        }
      }
    }
    if ((rc = M->d_inst_sub.upload(sub.data(), sub.size()))) return rc;
    if ((rc = M->d_out_idx.upload(oidx.data(), oidx.size()))) return rc;
    if (!need) CU_TRY(cudaMemset(M->d_f.p, 0, sizeof(double) * M->n_inst));  // unused sides read as 0 (never consumed)
    if ((rc = launch_filter(M, M->d_inst_sub.p, M->d_out_idx.p, (int64_t)sub.size(), eps, 0, nullptr, nullptr))) return rc;
  }
  if ((rc = run_reduction(M))) return rc;
  if (partial_device) *partial_device = M->d_partial.p;
  return PSY_OK;
}

psy_sweep* psy_sweep_create(int n_theta) {
  if (n_theta < 0) { fail(PSY_ERR_ARG, "psy_sweep_create: bad arguments"); return nullptr; }
  psy_sweep* S = new psy_sweep;
  S->P = n_theta;
  return S;
}
void psy_sweep_destroy(psy_sweep* S) { delete S; }

int psy_sweep_resolve(psy_sweep* S, const double* part, double eps, int level, int n_eps, int direction, unsigned char* need, int* done) {
  if (!S || !part || !need || !done) return fail(PSY_ERR_ARG, "psy_sweep_resolve: bad arguments");
  if (direction < 0 || direction > 2) return fail(PSY_ERR_ARG, "Unknown direction argument. Possible are central, left and right");
  const int P = S->P;
  if (level == 0) {
    S->base_total = part[0];
    S->base_nonfinite = part[1];
    S->Dval.assign((size_t)P * 2, 0.0);
    S->eps_used.assign(P, 0.0);
    S->resolved.assign((size_t)P * 2, 0);
    S->wanted.assign((size_t)P * 2, 0);
    for (int t = 0; t < P; t++) {
      S->wanted[2 * t + 0] = (direction == PSY_DIR_CENTRAL || direction == PSY_DIR_LEFT);
      S->wanted[2 * t + 1] = (direction == PSY_DIR_CENTRAL || direction == PSY_DIR_RIGHT);
    }
  }
  if (S->wanted.size() != (size_t)P * 2) return fail(PSY_ERR_STATE, "psy_sweep_resolve: level 0 has not been resolved");
  const double* Dp = part + 2;
  const double* Dm = part + 2 + (size_t)P;
  const double* nfp = part + 2 + 2 * (size_t)P;
  const double* nfm = part + 2 + 3 * (size_t)P;
  const double* nbt = part + 2 + 4 * (size_t)P;
  int pending = 0;
  for (int t = 0; t < P; t++) {
    for (int side = 0; side < 2; side++) {
      const size_t k = 2 * (size_t)t + side;
      need[k] = 0;
      if (!S->wanted[k] || S->resolved[k]) continue;
      const double D = side ? Dp[t] : Dm[t];
      const double nf = side ? nfp[t] : nfm[t];
      // Σ -2LL over ALL persons is finite  <=>  untouched persons finite at base, touched persons finite here
      const bool ok = (S->base_nonfinite - nbt[t] == 0.0) && nf == 0.0 && std::isfinite(D);
      if (ok) {
        S->resolved[k] = 1;
        S->Dval[k] = D;
        S->eps_used[t] += eps;  // R/modelTemplate.R:893, 914
      } else if (level + 1 < n_eps) {
        need[k] = 1;
        pending++;
      }
    }
  }
  *done = (pending == 0);
  return PSY_OK;
}

int psy_sweep_finish(psy_sweep* S, int direction, double* grad, double* total) {
  if (!S || !grad) return fail(PSY_ERR_ARG, "psy_sweep_finish: bad arguments");
  if (S->wanted.size() != (size_t)S->P * 2) return fail(PSY_ERR_STATE, "psy_sweep_finish: nothing resolved");
  (void)direction;
  const double nan = std::nan("");
  for (int t = 0; t < S->P; t++) {
    const bool wl = S->wanted[2 * t], wr = S->wanted[2 * t + 1];
    if ((wl && !S->resolved[2 * t]) || (wr && !S->resolved[2 * t + 1]) || (!wl && !wr)) { grad[t] = nan; continue; }
    if (!(wl && wr) && S->base_nonfinite != 0.0) { grad[t] = nan; continue; }  // one-sided: the other side is Σ f0
    const double dplus = wr ? S->Dval[2 * t + 1] : 0.0, dminus = wl ? S->Dval[2 * t] : 0.0;
    grad[t] = (dplus - dminus) / S->eps_used[t];  // R/modelTemplate.R:926
  }
  if (total) *total = S->base_total;
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
    CU_TRY(cudaMemcpy(part.data(), dptr, sizeof(double) * part.size(), cudaMemcpyDeviceToHost));
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
  int rc = load_driver();
  if (rc) return rc;
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, 0));
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
  if (!M || !M->fn) return fail(PSY_ERR_STATE, "psy_kernel_info: no module loaded");
  int v = 0;
  if (regs) { g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, M->fn); *regs = v; }
  if (local_bytes) { g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, M->fn); *local_bytes = v; }
  if (max_threads) { g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, M->fn); *max_threads = v; }
  return PSY_OK;
}

}  // extern "C"
