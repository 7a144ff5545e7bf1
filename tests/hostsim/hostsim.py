"""hostsim — TEST INFRASTRUCTURE ONLY: the kernel SOURCE (psy_filter.cuh + psy_lang.cuh + the model's generated translation
unit, exactly the text compileModel hands to nvcc) compiled by g++ through tests/hostsim/cuda_host_shim.h and run one
simulated CUDA thread at a time on the host.

Purpose: the `-m "not gpu"` suite can check the kernel's own arithmetic, indexing and instance-list conventions against
the oracle on a machine without a GPU (and a kernel change can be checked before GPU time is spent on it).  It is not a
fallback: nothing under psydiff_b200/ imports it, and the product path still fails loudly without a device.

The launch block (PsyLaunch, psy_launch.h), the instance list (person-major, base instance first, then (theta<<1)|side
pairs per distinct theta of the person) and the central-difference bookkeeping (differences formed per person before
summing) follow psy_capi.cu: psy_set_slots / launch_filter / psy_reduce_*.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "psydiff_b200", "csrc")
_BUILD = os.path.join(_HERE, "_build")

_DRIVER = r"""
extern "C" void psy_hostsim_run(const PsyLaunch* L, int block) {
  const int grid = (L->n_inst + block - 1) / block;
  if (psy_lanes_per_instance_in_rk4 == 1) {
    blockDim.x = (unsigned)block; blockDim.y = blockDim.z = 1;
    gridDim.x = (unsigned)grid; gridDim.y = gridDim.z = 1;
    for (int b = 0; b < grid; b++)
      for (int t = 0; t < block; t++) {
        blockIdx.x = (unsigned)b; blockIdx.y = blockIdx.z = 0;
        threadIdx.x = (unsigned)t; threadIdx.y = threadIdx.z = 0;
        psy_filter_kernel(*L);
      }
    return;
  }
  // lane-pair kernel: one simulated warp (PSY_SIM_WARP fibers, cuda_host_shim.h) at a time; warps that hold no instance are
  // skipped (they would only compute on zeros)
  static const PsyLaunch* launch;
  launch = L;
  blockDim.x = (unsigned)block; blockDim.y = blockDim.z = 1;
  gridDim.x = (unsigned)grid; gridDim.y = gridDim.z = 1;
  threadIdx.y = threadIdx.z = 0;
  for (int b = 0; b < grid; b++)
    for (int w0 = 0; w0 < block; w0 += PSY_SIM_WARP) {
      if (b * block + w0 >= L->n_inst) continue;
      blockIdx.x = (unsigned)b; blockIdx.y = blockIdx.z = 0;
      psy_sim_run_warp([](int) { psy_filter_kernel(*launch); }, (unsigned)w0);
    }
}
extern "C" int psy_hostsim_block(void) { return psy_block_size; }
extern "C" int psy_hostsim_lanes(void) { return psy_lanes_per_instance_in_rk4; }
#ifdef PSY_COUNT_OPS
extern "C" void psy_hostsim_ops(unsigned long long* out, int reset) {
  out[0] = psy_ops::add; out[1] = psy_ops::mul; out[2] = psy_ops::fma_; out[3] = psy_ops::div_; out[4] = psy_ops::special;
  if (reset) psy_ops::add = psy_ops::mul = psy_ops::fma_ = psy_ops::div_ = psy_ops::special = 0;
}
#endif
extern "C" const int* psy_hostsim_info(void) { return psy_module_info; }
"""


class PsyInst(C.Structure):
    _fields_ = [("person", C.c_int), ("code", C.c_int)]


class PsyLaunch(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("dt", C.c_void_p), ("ksteps", C.c_void_p), ("row_off", C.c_void_p),
                ("person_id", C.c_void_p), ("theta_index", C.c_void_p), ("theta", C.c_void_p), ("inst", C.c_void_p),
                ("f_out", C.c_void_p), ("out_idx", C.c_void_p), ("latent", C.c_void_p), ("pred", C.c_void_p),
                ("eps", C.c_double), ("h", C.c_double), ("c", C.c_double), ("sqrtc", C.c_double), ("wm0", C.c_double),
                ("wm1", C.c_double), ("wc0", C.c_double), ("wc1", C.c_double), ("n_inst", C.c_int), ("skip_update", C.c_int),
                ("n_theta", C.c_int)]


def _ptr(a):
    return a.ctypes.data if a is not None else None


def build(source: str, extra_flags: str = "") -> str:
    """g++-compile one model TU (+ the simulation driver) into tests/hostsim/_build/sim_<hash>.so."""
    hdrs = "".join(open(os.path.join(_CSRC, h)).read() for h in ("psy_filter.cuh", "psy_filter_pair.cuh", "psy_lang.cuh", "psy_launch.h"))
    shim = open(os.path.join(_HERE, "cuda_host_shim.h")).read()
    key = hashlib.sha1((source + hdrs + shim + _DRIVER + extra_flags).encode()).hexdigest()[:16]
    so = os.path.join(_BUILD, f"sim_{key}.so")
    if not os.path.exists(so):
        os.makedirs(_BUILD, exist_ok=True)
        cpp = os.path.join(_BUILD, f"sim_{key}_{os.getpid()}.cpp")   # per process: pytest-xdist workers may build the same model
        with open(cpp, "w") as f:
            f.write('#include "cuda_host_shim.h"\n' + source + _DRIVER)
        # -ffp-contract=fast + -mfma: a*b+c contracts to fma like nvcc's default -fmad=true (not necessarily the same
        # choices: agreement with the device is to rounding, not bit for bit)
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-mfma", "-ffp-contract=fast", "-Wno-unknown-pragmas",
               "-I", _HERE, "-I", _CSRC] + extra_flags.split() + ["-o", so + f".tmp{os.getpid()}", cpp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("hostsim: g++ failed:\n" + r.stderr[:4000])
        os.replace(so + f".tmp{os.getpid()}", so)
    return so


class HostSim:
    """The model's kernel, host-compiled, bound to the model's data and slot map."""

    def __init__(self, model, extra_flags: str = ""):
        from psydiff_b200 import _lib, model as _model
        _model._sync_source(model)
        self.model = model
        self.lib = C.CDLL(build(model["cppCode"], extra_flags))
        self.lib.psy_hostsim_info.restype = C.POINTER(C.c_int)
        self.block = int(self.lib.psy_hostsim_block())
        self.lanes = int(self.lib.psy_hostsim_lanes())     # 2 = the lane-pair kernel (psy_filter_pair.cuh) is what was compiled
        info = self.lib.psy_hostsim_info()
        self.info = [int(info[i]) for i in range(6)]
        spec = model["_spec"]
        self.n, self.m, self.F = spec.nlatent, spec.nmanifest, spec.n_slots
        pt = model["pars"]["parameterTable"]
        self.P = len(pt.theta_labels)
        self.row_off = np.ascontiguousarray(model["_row_off"], dtype=np.int64)
        self.obs = np.ascontiguousarray(model["data"]["observations"], dtype=np.float64)   # [rows][m] row-major
        self.dt = np.ascontiguousarray(model["data"]["dt"], dtype=np.float64)
        self.pid = np.ascontiguousarray(np.asarray(model["_persons_unique"], dtype=np.float64).astype(np.int32))
        self.N = len(self.pid)
        self.rows = int(self.row_off[-1])
        self.tidx = np.ascontiguousarray(pt.theta_index, dtype=np.int32).reshape(self.N, self.F)   # N x 0 for a model without free parameters
        self._rk = _lib.lib().psy_rk4_step_count          # host-only exported helper of the C ABI (rule T1)
        self._rk.argtypes, self._rk.restype = [C.c_double, C.c_double], C.c_int

    # ---- the instance list of psy_set_slots (without warp-alignment padding unless pad=True) ----------------------
    def instance_list(self, pad=False):
        inst, start = [], np.zeros(self.N, dtype=np.int64)
        distinct = []
        for p in range(self.N):
            dl = list(dict.fromkeys(int(t) for t in self.tidx[p])) if self.F else []
            distinct.append(dl)
            cnt = 1 + 2 * len(dl)
            padded = (cnt + 31) // 32 * 32
            align = pad and (padded - cnt) * 10 <= cnt
            if align:
                while len(inst) % 32:
                    inst.append((-1, -1))
            start[p] = len(inst)
            inst.append((p, -1))
            for t in dl:
                inst.append((p, (t << 1) | 0))
                inst.append((p, (t << 1) | 1))
            if align:
                while len(inst) % 32:
                    inst.append((-1, -1))
        arr = (PsyInst * len(inst))(*[PsyInst(a, b) for a, b in inst])
        return arr, start, distinct

    def _launch(self, inst, n_inst, thetas, eps, skip_update=False, latent=None, pred=None):
        mdl = self.model
        n = self.n
        kappa = float(mdl["kappa"])
        alpha, beta = float(mdl["alpha"]), float(mdl["beta"])
        if n + kappa == 0:
            kappa -= 1
        lam = alpha * alpha * (n + kappa) - n
        h = float(np.atleast_1d(mdl["timeStep"])[0])
        ks = np.array([self._rk(float(d), h) for d in self.dt], dtype=np.int32)
        f = np.full(n_inst, np.nan)
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        if th.size == 0:
            th = np.zeros(1)                               # the library allocates max(P, 1) doubles as well
        L = PsyLaunch()
        L.obs, L.dt, L.ksteps, L.row_off = _ptr(self.obs), _ptr(self.dt), _ptr(ks), _ptr(self.row_off)
        L.person_id, L.theta_index, L.theta = _ptr(self.pid), _ptr(self.tidx), _ptr(th)
        L.inst = C.addressof(inst)
        L.f_out, L.out_idx, L.latent, L.pred = _ptr(f), None, _ptr(latent), _ptr(pred)
        L.eps, L.h = float(eps), h
        L.c = alpha * alpha * (n + kappa)
        L.sqrtc = float(np.sqrt(L.c))
        L.wm0 = lam / (n + lam)
        L.wm1 = 1 / (2 * (n + lam))
        L.wc0 = lam / (n + lam) + (1 - alpha * alpha + beta)
        L.wc1 = L.wm1
        L.n_inst, L.skip_update, L.n_theta = n_inst, 1 if skip_update else 0, self.P
        self.lib.psy_hostsim_run(C.byref(L), self.block)
        return f

    def theta(self):
        return np.ascontiguousarray(self.model["pars"]["parameterTable"].theta_values, dtype=np.float64)

    def ops(self, reset=True):
        """Floating-point operations the kernel SOURCE executed since the last reset (only with extra_flags "-DPSY_COUNT_OPS"):
        dict(add, mul, fma, div, special); flops = add + mul + 2 fma + div + special."""
        out = (C.c_ulonglong * 5)()
        self.lib.psy_hostsim_ops(out, 1 if reset else 0)
        d = dict(zip(("add", "mul", "fma", "div", "special"), (int(v) for v in out)))
        d["flops"] = d["add"] + d["mul"] + 2 * d["fma"] + d["div"] + d["special"]
        return d

    def fit(self, theta=None, skip_update=False, states=True):
        """psy_fit: base instances only -> per-person -2LL (+ latentScores rows x n, predictedManifest rows x m)."""
        th = self.theta() if theta is None else np.asarray(theta, dtype=np.float64)
        inst = (PsyInst * self.N)(*[PsyInst(p, -1) for p in range(self.N)])
        lat = np.full((self.rows, self.n), np.nan) if states else None
        pred = np.full((self.rows, self.m), np.nan) if states else None
        f = self._launch(inst, self.N, th, 0.0, skip_update, lat, pred)
        return {"m2LL": f, "latentScores": lat, "predictedManifest": pred}

    def fit_many(self, thetas, skip_update=False):
        """psy_fit_many: set-major blocks of persons, code = -1 - set."""
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        thetas = thetas.reshape(-1, self.P) if self.P else np.zeros((max(len(thetas), 1), 1))
        k = thetas.shape[0]
        inst = (PsyInst * (self.N * k))(*[PsyInst(p, -1 - s) for s in range(k) for p in range(self.N)])
        f = self._launch(inst, self.N * k, thetas, 0.0, skip_update).reshape(k, self.N)
        return f

    def gradients(self, theta=None, eps=1e-6, pad=True):
        """One central-difference level of psy_gradients: g_t = Σ_p [(f+ − f0) − (f− − f0)] / (2 eps), plus the base -2LL."""
        th = self.theta() if theta is None else np.asarray(theta, dtype=np.float64)
        inst, start, distinct = self.instance_list(pad)
        f = self._launch(inst, len(inst), th, eps)
        dplus, dminus = np.zeros(self.P), np.zeros(self.P)
        base = np.empty(self.N)
        for p in range(self.N):
            o = int(start[p])
            base[p] = f[o]
            for j, t in enumerate(distinct[p]):
                dminus[t] += f[o + 1 + 2 * j] - f[o]
                dplus[t] += f[o + 2 + 2 * j] - f[o]
        return {"grad": (dplus - dminus) / (2 * eps), "m2LL": base, "f": f}
