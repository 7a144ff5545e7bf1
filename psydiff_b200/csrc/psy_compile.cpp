// psy_compile.cpp — compileModel (R/prepareModel.R:553-562) for the device: the generated CUDA translation unit is compiled
// for sm_100a and cached by content hash, with an `nvcc -cubin` subprocess (default) or in-process through a dlopen'ed libnvrtc
// (PSY_COMPILER=nvrtc).  Host-only code; no GPU is needed for it.
#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include "psy_internal.h"

#define fail psy_fail

namespace {

thread_local std::string g_err;

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


// mkdir -p; returns false (errno set) when a component cannot be created
bool make_dirs(const std::string& path) {
  if (path.empty()) return false;
  for (size_t i = 1; i <= path.size(); i++) {
    if (i != path.size() && path[i] != '/') continue;
    const std::string part = path.substr(0, i);
    if (mkdir(part.c_str(), 0755) != 0 && errno != EEXIST) return false;
  }
  struct stat st;
  return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

// Runs argv[0] with the given arguments (no shell: paths with spaces and hostile flag strings stay single arguments),
// stdout + stderr appended to `log_path`.  Returns the exit status, or -1 when the process could not be started.
int run_process(const std::vector<std::string>& argv, const std::string& log_path) {
  std::vector<char*> av;
  for (auto& a : argv) av.push_back(const_cast<char*>(a.c_str()));
  av.push_back(nullptr);
  const pid_t pid = fork();
  if (pid < 0) return -1;
  if (pid == 0) {
    FILE* lf = fopen(log_path.c_str(), "w");
    if (lf) { dup2(fileno(lf), 1); dup2(fileno(lf), 2); }
    execvp(av[0], av.data());
    fprintf(stderr, "cannot execute %s: %s\n", av[0], strerror(errno));
    _exit(127);
  }
  int status = 0;
  while (waitpid(pid, &status, 0) < 0)
    if (errno != EINTR) return -1;
  return WIFEXITED(status) ? WEXITSTATUS(status) : -1;
}

}  // namespace

int psy_fail(int code, const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
// error reporting for psy_codegen.cpp
int psy_set_error(int code, const char* msg) { return psy_fail(code, "%s", msg); }

extern "C" {

const char* psy_last_error(void) { return g_err.c_str(); }
int psy_version(void) { return 200; }

int psy_model_compile(const char* cuda_source, const char* include_dir, const char* cache_dir, const char* extra_flags,
                      char* cubin_path, size_t cap) {
  if (!cuda_source || !include_dir || !cache_dir || !cubin_path) return fail(PSY_ERR_ARG, "psy_model_compile: null argument");
  const std::string inc(include_dir), cache(cache_dir), flags(extra_flags ? extra_flags : "");
  uint64_t h = fnv1a(cuda_source);
  for (const char* hdr : {"/psy_filter.cuh", "/psy_filter_pair.cuh", "/psy_lang.cuh", "/psy_launch.h"}) {
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
  if (!make_dirs(cache)) return fail(PSY_ERR_COMPILE, "cannot create the kernel cache directory %s: %s", cache.c_str(), strerror(errno));
  const std::string cu = cache + "/" + name + ".cu", cubin = cache + "/" + name + ".cubin", log = cache + "/" + name + ".log";
  if (cubin.size() + 1 > cap) return fail(PSY_ERR_ARG, "cubin_path buffer too small");
  if (!file_exists(cubin)) {
    {
      std::ofstream f(cu);
      f << cuda_source;
      if (!f) return fail(PSY_ERR_COMPILE, "cannot write %s", cu.c_str());
    }
    const std::string tmp = cubin + ".tmp" + std::to_string((long long)getpid());
    if (use_nvrtc) {
      std::string image, clog;
      if (nvrtc_compile(cuda_source, inc, flags, image, clog) != 0) {
        if (clog.size() > 3000) clog = clog.substr(0, 3000) + "\n...";
        return fail(PSY_ERR_COMPILE,
                    "model does not compile for the device (statements outside the supported equation subset are "
                    "unsupported on device):\n%s", clog.c_str());
      }
      std::ofstream f(tmp, std::ios::binary);
      f.write(image.data(), (std::streamsize)image.size());
      if (!f) return fail(PSY_ERR_COMPILE, "cannot write %s", tmp.c_str());
    } else {
      const char* nvcc_env = getenv("PSY_NVCC");
      const std::string nvcc = nvcc_env ? nvcc_env : (file_exists("/usr/local/cuda/bin/nvcc") ? "/usr/local/cuda/bin/nvcc" : "nvcc");
      std::vector<std::string> argv = {nvcc, "-cubin", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                                       "-Xptxas", "-v", "-I" + inc};
      std::stringstream ss(flags);
      for (std::string tok; ss >> tok;) argv.push_back(tok);
      argv.insert(argv.end(), {"-o", tmp, cu});
      const int rc = run_process(argv, log);
      if (rc != 0 || !file_exists(tmp)) {
        std::string msg = slurp(log);
        if (msg.size() > 3000) msg = msg.substr(0, 3000) + "\n...";
        remove(tmp.c_str());
        return fail(PSY_ERR_COMPILE,
                    "model does not compile for the device (statements outside the supported equation subset are "
                    "unsupported on device):\n%s", msg.c_str());
      }
    }
    if (rename(tmp.c_str(), cubin.c_str()) != 0) return fail(PSY_ERR_COMPILE, "cannot move %s into the cache", tmp.c_str());
  }
  strcpy(cubin_path, cubin.c_str());
  return PSY_OK;
}

}  // extern "C"
