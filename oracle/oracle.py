"""Python driver of the CPU oracle (oracle/psy_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

PARITY STATUS — see the header of psy_oracle.cpp: the reference holds no golden vectors for this path; the oracle is pinned to
the reference's OWN C++ compiled and run here against stand-in third-party headers (oracle/reference_build.py,
tests/test_reference_build.py, tests/golden/*_reference.json) and further anchored on the exact linear Kalman filter,
SR == covariance for linear h, primitive identities, a numpy/LAPACK restatement and its own 80-bit build.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It takes the model object built by ``psydiff_b200.newPsydiff`` (the shared INPUT description: equations, containers,
slot map, data) and evaluates it with the restated reference algorithm on the CPU; it never touches the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
CXX = os.environ.get("CXX", "g++")
CXXFLAGS = ["-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread"]


def _compile(src_path: str, out: str, extra=()):
    os.makedirs(_BUILD, exist_ok=True)
    tmp = out + f".tmp{os.getpid()}"
    subprocess.run([CXX] + CXXFLAGS + list(extra) + ["-o", tmp, src_path], check=True)
    os.replace(tmp, out)


_EXT = ['-DPSY_REAL=long double', "-DPSY_EXT"]


def build(force=False, count_flops=False, extended=False) -> str:
    """Compile the oracle library (g++ -O2 -ffp-contract=off).  ``extended`` builds the same source with x87 80-bit
    long double arithmetic (used to bound the double build's own rounding noise).  Returns the shared object path."""
    src = os.path.join(_HERE, "psy_oracle.cpp")
    name = "libpsy_oracle" + ("_flops" if count_flops else "") + ("_ext" if extended else "") + ".so"
    out = os.path.join(_BUILD, name)
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        _compile(src, out, (["-DPSY_COUNT_FLOPS"] if count_flops else []) + (_EXT if extended else []))
    return out


class psy_oracle_problem(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("m", C.c_int), ("alpha", C.c_double), ("beta", C.c_double), ("kappa", C.c_double),
        ("stepper", C.c_int), ("h", C.c_double), ("sr", C.c_int), ("skip_update", C.c_int),
        ("off_m0", C.c_int), ("off_A0", C.c_int), ("off_L", C.c_int), ("off_R", C.c_int), ("plist_len", C.c_int),
        ("drift", C.c_void_p), ("meas", C.c_void_p),
        ("N", C.c_int), ("F", C.c_int), ("P", C.c_int),
        ("base_plist", C.POINTER(C.c_double)), ("slot_off", C.POINTER(C.c_int)), ("theta_index", C.POINTER(C.c_int)),
        ("row_off", C.POINTER(C.c_int64)), ("obs", C.POINTER(C.c_double)), ("dt", C.POINTER(C.c_double)),
        ("person_id", C.POINTER(C.c_int)),
    ]


_libs = {}


def _lib(count_flops=False, extended=False):
    key = (bool(count_flops), bool(extended))
    if key not in _libs:
        L = C.CDLL(build(count_flops=count_flops, extended=extended))
        L.psy_oracle_fit.restype = C.c_int
        L.psy_oracle_fit.argtypes = [C.POINTER(psy_oracle_problem), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.psy_oracle_gradients.restype = C.c_int
        L.psy_oracle_gradients.argtypes = [C.POINTER(psy_oracle_problem), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                           C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.psy_oracle_rk4_steps.restype = C.c_int
        L.psy_oracle_rk4_steps.argtypes = [C.c_double, C.c_double, C.c_double]
        L.psy_oracle_flops_reset.restype = C.c_ulonglong
        _libs[key] = L
    return _libs[key]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def emit_host_cpp(spec) -> tuple:
    """Wrap the user's statements the way newPsydiff does (R/prepareModel.R:479-505), against psy_oracle_lang.hpp.
    Every container referenced in the text is bound from the person's flat parameter list (prepareEquations, :737-770).
    Returns (source, offsets) where offsets[name] is the container's position in the flat list (column-major)."""
    off, cur = {}, 0
    for c in spec.containers:
        off[c.name] = cur
        cur += c.values.size
    names = [c.name for c in spec.containers]

    def binds(eq):
        toks = [t for t in re.split(r"[^A-Za-z0-9_]", eq) if t]
        out, seen = [], set()
        for t in toks:
            if t in names and t not in seen:
                seen.add(t)
                c = spec.container(t)
                if c.kind == "double":
                    out.append(f"  const real {t} = plist[{off[t]}];\n")
                else:
                    out.append(f"  ol::Mat {t}(plist + {off[t]}, {c.shape[0]}, {c.shape[1]});\n")
        return "".join(out)

    n, m = spec.nlatent, spec.nmanifest
    src = f"""// generated by oracle/oracle.py — user equations for the CPU oracle
#include "{os.path.join(_HERE, 'psy_oracle_lang.hpp')}"
using std::exp; using std::log; using std::pow; using std::sqrt; using std::sin; using std::cos; using std::tanh; using std::fabs;
extern "C" void psy_user_drift(const real* x_, real* dx_, const real* plist, int person, real t) {{
  const ol::Vec x(x_, {n}); ol::Vec dx({n});
{binds(spec.latentEquations)}  {spec.latentEquations.strip()}
  for (int i = 0; i < {n}; i++) dx_[i] = dx.v[i];
}}
extern "C" void psy_user_meas(const real* x_, real* y_, const real* plist, int person, real t) {{
  const ol::Vec x(x_, {n}); ol::Vec y({m});
{binds(spec.manifestEquations)}  {spec.manifestEquations.strip()}
  for (int i = 0; i < {m}; i++) y_[i] = y.v[i];
}}
"""
    return src, off, cur


class Oracle:
    """CPU oracle bound to one model object (built by psydiff_b200.newPsydiff)."""

    def __init__(self, model, count_flops=False, extended=False):
        spec = model["_spec"]
        self.model, self.spec = model, spec
        self.L = _lib(count_flops, extended)
        src, off, plen = emit_host_cpp(spec)
        h = hashlib.sha1(src.encode()).hexdigest()[:16] + ("_ext" if extended else "")
        so = os.path.join(_BUILD, f"user_{h}.so")
        if not os.path.exists(so):
            os.makedirs(_BUILD, exist_ok=True)
            cpp = os.path.join(_BUILD, f"user_{h}.cpp")
            with open(cpp, "w") as f:
                f.write(src)
            _compile(cpp, so, _EXT if extended else [])
        self.user = C.CDLL(so)
        self.off, self.plen = off, plen
        base = np.zeros(plen)
        slot_off = np.zeros(max(spec.n_slots, 1), dtype=np.int32)
        for c in spec.containers:
            base[off[c.name]:off[c.name] + c.values.size] = c.values.reshape(-1, order="F")
            r = c.shape[0]
            for i in range(c.shape[0]):
                for j in range(c.shape[1]):
                    k = int(c.slots[i, j])
                    if k >= 0:
                        slot_off[k] = off[c.name] + i + j * r
        pt = model["pars"]["parameterTable"]
        self.base, self.slot_off = base, slot_off
        self.tidx = np.ascontiguousarray(pt.theta_index, dtype=np.int32)
        self.row_off = np.ascontiguousarray(model["_row_off"], dtype=np.int64)
        self.obs = np.ascontiguousarray(model["data"]["observations"], dtype=np.float64)  # rows x m ROW-major
        self.dt = np.ascontiguousarray(model["data"]["dt"], dtype=np.float64)
        self.pid = np.ascontiguousarray(np.asarray(model["_persons_unique"], dtype=np.float64).astype(np.int32))
        self.N, self.F, self.P = self.tidx.shape[0], spec.n_slots, len(pt.theta_labels)

    def _problem(self, skip_update=False, persons=None):
        m = self.model
        q = psy_oracle_problem()
        q.n, q.m = self.spec.nlatent, self.spec.nmanifest
        q.alpha, q.beta, q.kappa = float(m["alpha"]), float(m["beta"]), float(m["kappa"])
        q.stepper = {"rk4": 0, "runge_kutta_dopri5": 1}.get(m["integrateFunction"], 2)
        q.h = float(np.atleast_1d(m["timeStep"])[0])
        q.sr = 1 if m.get("srUpdate", True) else 0
        q.skip_update = 1 if skip_update else 0
        q.off_m0, q.off_A0, q.off_L, q.off_R = (self.off[k] for k in ("m0", "A0LogChol", "LLogChol", "RLogChol"))
        q.plist_len = self.plen
        q.drift = C.cast(self.user.psy_user_drift, C.c_void_p)
        q.meas = C.cast(self.user.psy_user_meas, C.c_void_p)
        if persons is None:
            tidx, row_off, pid = self.tidx, self.row_off, self.pid
        else:  # a person subset shares obs/dt through row_off
            persons = np.asarray(persons)
            tidx = np.ascontiguousarray(self.tidx[persons])
            pid = np.ascontiguousarray(self.pid[persons])
            # re-pack rows of the subset contiguously
            rows = np.concatenate([np.arange(self.row_off[p], self.row_off[p + 1]) for p in persons])
            lens = np.array([self.row_off[p + 1] - self.row_off[p] for p in persons], dtype=np.int64)
            row_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            self._sub_obs = np.ascontiguousarray(self.obs[rows])
            self._sub_dt = np.ascontiguousarray(self.dt[rows])
        q.N, q.F, q.P = len(pid), self.F, self.P
        q.base_plist = _dp(self.base)
        q.slot_off = _ip(self.slot_off)
        self._keep = (tidx, row_off, pid)
        q.theta_index = _ip(tidx)
        q.row_off = row_off.ctypes.data_as(C.POINTER(C.c_int64))
        q.obs = _dp(self.obs if persons is None else self._sub_obs)
        q.dt = _dp(self.dt if persons is None else self._sub_dt)
        q.person_id = _ip(pid)
        return q

    def theta(self):
        return np.ascontiguousarray(self.model["pars"]["parameterTable"].theta_values, dtype=np.float64)

    def fit(self, theta=None, skip_update=False, states=True, nthreads=1, persons=None):
        """fitModel: returns dict(m2LL[N], latentScores rows x n, predictedManifest rows x m)."""
        th = self.theta() if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        q = self._problem(skip_update, persons)
        rows = int(self._keep[1][-1])
        m2 = np.empty(q.N)
        lat = np.full((rows, q.n), np.nan) if states else None
        pred = np.full((rows, q.m), np.nan) if states else None
        self.L.psy_oracle_fit(C.byref(q), _dp(th), _dp(m2), _dp(lat) if states else None, _dp(pred) if states else None, nthreads)
        return {"m2LL": m2, "latentScores": lat, "predictedManifest": pred}

    def gradients(self, theta=None, eps=None, direction=None, nthreads=1, persons=None):
        """getGradients: returns dict(grad_ref, grad_clean, m2LL) (see psy_oracle_gradients)."""
        th = self.theta() if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        eps = np.ascontiguousarray(self.model["eps"] if eps is None else eps, dtype=np.float64)
        d = {"central": 0, "left": 1, "right": 2}[self.model["direction"] if direction is None else direction]
        q = self._problem(False, persons)
        g1, g2, base = np.empty(self.P), np.empty(self.P), np.empty(q.N)
        self.L.psy_oracle_gradients(C.byref(q), _dp(th), _dp(eps), len(eps), d, _dp(g1), _dp(g2), _dp(base), nthreads)
        return {"grad_ref": g1, "grad_clean": g2, "m2LL": base}

    def flops_reset(self):
        return int(self.L.psy_oracle_flops_reset())
