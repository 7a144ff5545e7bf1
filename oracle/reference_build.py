"""oracle/_ref — the REFERENCE'S OWN C++ for this path, compiled here and run as the anchor of the oracle.  TEST INFRASTRUCTURE.

psydiff's fitModel / getGradients / getGradient are C++ held as a string in R/modelTemplate.R:5-1011; newPsydiff splices the
user's equations into it (R/prepareModel.R:479-512, 737-770) and compileModel hands the text to Rcpp::sourceCpp.  R, Rcpp,
Armadillo and Boost are absent from this image, so the package cannot be built the way its authors build it.  This
module does the next best thing, entirely from the sources where they lie under /root/reference (nothing of them is copied
into the repository; everything generated goes to oracle/_ref/, which is git-ignored):

  1. reads the template text out of R/modelTemplate.R (the SR or the covariance variant + the gradient functions),
  2. splices the model's equations in exactly as newPsydiff / prepareEquations do (restated below, the only R code involved),
  3. compiles the result — which #includes the reference's inst/include/psydiff.h, psydiffBasic.h, psydiffInternals.h — with
     g++ against oracle/refbuild/include: stand-in headers for the THIRD-PARTY interfaces only (Rcpp containers with R's
     reference semantics, the Armadillo operations the reference uses, Boost.odeint's integrate_const<runge_kutta4> and
     integrate = controlled dopri5, restated from their published algorithms: SURVEY §8c T1-T3),
  4. drives it through the small C ABI of oracle/refbuild/ref_driver.inc.

So the person loop, the time loop, the parameter plumbing (setParameterList_C and its `changed` flags), every filter primitive,
the finite-difference bookkeeping and every quirk of the reference run as their authors wrote them; only LAPACK-level numerics
(QR, triangular inverse, Cholesky, LU) and the ODE stepper are restatements.  tests/test_reference_build.py checks the oracle
(oracle/psy_oracle.cpp) against it, and tests/golden/make_reference_golden.py writes its outputs into the committed fixtures the GPU
tests read (the GPU box has no /root/reference).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("PSY_REFERENCE", "/root/reference")
_OUT = os.path.join(_HERE, "_ref")
_INC = os.path.join(_HERE, "refbuild", "include")
_DRIVER = os.path.join(_HERE, "refbuild", "ref_driver.inc")


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE, "R", "modelTemplate.R")) and \
        os.path.exists(os.path.join(REFERENCE, "inst", "include", "psydiffInternals.h"))


def template(srUpdate: bool = True) -> str:
    """modelTemplate(srUpdate) (R/modelTemplate.R:5-1011): the text of the selected `mod <- '...'` literal followed by the
    `gradients <- '...'` literal, read from the reference's file at build time."""
    text = open(os.path.join(REFERENCE, "R", "modelTemplate.R")).read()
    mods = [m.end() for m in re.finditer(r"mod <- '", text)]
    grad = re.search(r"gradients <- '", text)
    if len(mods) != 2 or not grad:
        raise RuntimeError("R/modelTemplate.R does not have the expected layout (two `mod <- '` literals and `gradients <- '`)")

    def literal(start):            # an R single-quoted literal; the template contains no single quotes of its own
        end = text.index("'", start)
        return text[start:end]

    return literal(mods[0] if srUpdate else mods[1]) + literal(grad.end())


def prepareEquations(equations: str, spec) -> str:
    """prepareEquations (R/prepareModel.R:737-770), restated: one `arma::mat X = modelPars["X"];` / `double X = ...` line per
    entry of `parameters` the statements name (first appearance, no repeats), then the statements."""
    tokens = [t for t in re.split(r"[^A-Za-z0-9|^_]", equations) if t]       # stringr::str_split(equations, '[^[:alnum:]|^_]')
    out = "Rcpp::List modelPars = pars.parameterList;\n  "
    targets = [c.name for c in spec.containers]
    for tok in tokens:
        if tok in targets:
            c = spec.container(tok)
            kind = "double" if c.kind == "double" else "arma::mat"
            out += f'\n               {kind} {tok} = modelPars["{tok}"]; \n'
            targets.remove(tok)
    return out + "\n                              " + equations


def model_source(spec) -> str:
    """newPsydiff's assembly (R/prepareModel.R:479-512): wrappers around the prepared equations, pasted over the placeholders."""
    manifest = ("\n  arma::colvec measurementEquation(const arma::colvec x, const odeintpars &pars,\n"
                "                                   const int &person, const double &t){\n"
                "    arma::colvec y(pars.nmanifest, arma::fill::zeros);\n  "
                + prepareEquations(spec.manifestEquations, spec) + "\n    return(y);\n  }\n  ")
    latent = ("\n  arma::colvec latentEquation(const arma::colvec x, const odeintpars &pars,\n"
              "                            const int &person, const double &t){\n"
              "    arma::colvec dx(pars.nlatent, arma::fill::zeros);\n  "
              + prepareEquations(spec.latentEquations, spec) + "\n    return(dx);\n  }\n  ")
    code = template(spec.srUpdate)
    code = code.replace("MEASUREMENTEQUATIONPLACEHOLDER", manifest).replace("LATENTEQUATIONPLACEHOLDER", latent)
    return code + '\n#include "ref_driver.inc"\n'


def _manifest_key(spec) -> str:
    """Identifies a model without reading the reference: its equations, containers' shapes/kinds and update variant."""
    desc = repr((spec.nlatent, spec.nmanifest, spec.latentEquations, spec.manifestEquations, bool(spec.srUpdate),
                 [(c.name, c.kind, tuple(c.values.shape)) for c in spec.containers]))
    return hashlib.sha1(desc.encode()).hexdigest()[:16]


def prebuilt(spec):
    """Path of the binary built earlier for this model (oracle/_ref/manifest.json), or None.  What the GPU box — which has the
    prebuilt files but no /root/reference — uses."""
    import json
    mf = os.path.join(_OUT, "manifest.json")
    if not os.path.exists(mf):
        return None
    so = json.load(open(mf)).get(_manifest_key(spec))
    return os.path.join(_OUT, so) if so and os.path.exists(os.path.join(_OUT, so)) else None


def build(spec) -> str:
    """Compile the model's reference translation unit into oracle/_ref/ref_<hash>.so (g++ only; needs /root/reference; where it
    is absent the binary recorded in oracle/_ref/manifest.json for this model is returned)."""
    if not available():
        so = prebuilt(spec)
        if so:
            return so
        raise RuntimeError(f"reference sources not found under {REFERENCE} and no prebuilt binary for this model in oracle/_ref")
    src = model_source(spec)
    deps = "".join(open(p).read() for p in (
        _DRIVER, os.path.join(_INC, "RcppArmadillo.h"), os.path.join(_INC, "boost", "numeric", "odeint.hpp"),
        os.path.join(REFERENCE, "inst", "include", "psydiffBasic.h"), os.path.join(REFERENCE, "inst", "include", "psydiffInternals.h")))
    key = hashlib.sha1((src + deps).encode()).hexdigest()[:16]
    so = os.path.join(_OUT, f"ref_{key}.so")
    if not os.path.exists(so):
        os.makedirs(_OUT, exist_ok=True)
        # the assembled translation unit (reference text + spliced equations) is piped to the compiler, never written to disk:
        # only the binary lands in oracle/_ref/.  -ffp-contract=off: R's default flags (-O2, no FMA contraction on x86-64)
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", _INC,
               "-I", os.path.join(REFERENCE, "inst", "include"), "-I", os.path.join(_HERE, "refbuild"), "-o", so + ".tmp", "-x", "c++", "-"]
        r = subprocess.run(cmd, input=src, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference translation unit does not compile against the stand-in headers:\n" + r.stderr[:6000])
        os.replace(so + ".tmp", so)
    import json
    mf = os.path.join(_OUT, "manifest.json")
    man = json.load(open(mf)) if os.path.exists(mf) else {}
    if man.get(_manifest_key(spec)) != os.path.basename(so):
        man[_manifest_key(spec)] = os.path.basename(so)
        json.dump(man, open(mf, "w"), indent=1)
    return so


def prune():
    """Delete binaries in oracle/_ref that the manifest no longer points at (left behind when a header or a model changed)."""
    import json
    mf = os.path.join(_OUT, "manifest.json")
    if not os.path.exists(mf):
        return 0
    keep = set(json.load(open(mf)).values())
    n = 0
    for f in os.listdir(_OUT):
        if f.startswith("ref_") and f.endswith(".so") and f not in keep:
            os.remove(os.path.join(_OUT, f))
            n += 1
    return n


class _Container(C.Structure):
    _fields_ = [("name", C.c_char_p), ("is_scalar", C.c_int), ("rows", C.c_int), ("cols", C.c_int), ("values", C.POINTER(C.c_double))]


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("N", C.c_int), ("rows", C.c_longlong),
                ("obs", C.POINTER(C.c_double)), ("dt", C.POINTER(C.c_double)), ("person", C.POINTER(C.c_double)),
                ("alpha", C.c_double), ("beta", C.c_double), ("kappa", C.c_double),
                ("time_step", C.POINTER(C.c_double)), ("n_time_step", C.c_int), ("integrate_function", C.c_char_p),
                ("break_early", C.c_int), ("verbose", C.c_int), ("eps", C.POINTER(C.c_double)), ("n_eps", C.c_int),
                ("direction", C.c_char_p), ("plist", C.POINTER(_Container)), ("n_plist", C.c_int), ("n_ptab", C.c_int),
                ("label", C.POINTER(C.c_char_p)), ("value", C.POINTER(C.c_double)), ("target", C.POINTER(C.c_char_p)),
                ("target_class", C.POINTER(C.c_char_p)), ("row", C.POINTER(C.c_double)), ("col", C.POINTER(C.c_double)),
                ("ptab_person", C.POINTER(C.c_double)), ("changed", C.POINTER(C.c_int))]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Reference:
    """The reference's fitModel / getGradients / getGradient bound to one model object (built by psydiff_b200.newPsydiff)."""

    def __init__(self, model):
        from psydiff_b200 import model as _model
        _model._sync_source(model)
        self.model, self.spec = model, model["_spec"]
        self.lib = C.CDLL(build(self.spec))

    def _problem(self, theta=None, changed=None):
        m, spec = self.model, self.spec
        pt = m["pars"]["parameterTable"]
        keep = []

        def arr(a, dtype=np.float64):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a

        def strs(values):
            a = (C.c_char_p * len(values))(*[str(v).encode() for v in values])
            keep.append(a)
            return a

        q = _Problem()
        obs = np.asfortranarray(m["data"]["observations"], dtype=np.float64)
        keep.append(obs)
        q.n, q.m, q.N, q.rows = spec.nlatent, spec.nmanifest, len(m["_persons_unique"]), obs.shape[0]
        q.obs, q.dt = _dp(obs), _dp(arr(m["data"]["dt"]))
        q.person = _dp(arr(np.asarray(m["data"]["person"], dtype=float)))
        q.alpha, q.beta, q.kappa = float(m["alpha"]), float(m["beta"]), float(m["kappa"])
        ts = arr(np.atleast_1d(m["timeStep"]))
        q.time_step, q.n_time_step = _dp(ts), len(ts)
        q.integrate_function = m["integrateFunction"].encode()
        q.break_early, q.verbose = int(bool(m["breakEarly"])), 0
        eps = arr(np.atleast_1d(m["eps"]))
        q.eps, q.n_eps = _dp(eps), len(eps)
        q.direction = m["direction"].encode()
        cs = (_Container * len(spec.containers))()
        for k, c in enumerate(spec.containers):
            vals = arr(np.asfortranarray(c.values, dtype=np.float64).reshape(-1, order="F"))
            cs[k] = _Container(c.name.encode(), 1 if c.kind == "double" else 0, c.values.shape[0], c.values.shape[1], _dp(vals))
        keep.append(cs)
        q.plist, q.n_plist = cs, len(spec.containers)
        # model$pars$parameterTable, person-major (R/prepareModel.R:471-476 + makeGroups)
        values = np.asarray(pt.value, dtype=float) if theta is None else np.asarray(theta, dtype=float)[pt.theta_index.reshape(-1)]
        npers = len(pt.persons)
        q.n_ptab = len(pt)
        q.label, q.value = strs(list(pt.label)), _dp(arr(values))
        q.target = strs(list(pt.target))
        q.target_class = strs(["double" if spec.container(r["target"]).kind == "double" else "matrix" for r in pt.slot_rows] * npers)
        q.row = _dp(arr([float(r["row"]) for r in pt.slot_rows] * npers))
        q.col = _dp(arr([float(r["col"]) for r in pt.slot_rows] * npers))
        q.ptab_person = _dp(arr(np.asarray(pt.person, dtype=float)))
        ch = arr(np.ones(len(pt)) if changed is None else changed, dtype=np.int32)
        q.changed = ch.ctypes.data_as(C.POINTER(C.c_int))
        self._keep = keep
        return q

    def fit(self, theta=None, skip_update=False, changed=None):
        """fitModel(psydiffModel, skipUpdate): dict(m2LL[N] in unique(person) order, latentScores, predictedManifest, changed)."""
        q = self._problem(theta, changed)
        m2 = np.empty(q.N)
        lat = np.empty((q.rows, q.n), order="F")
        pred = np.empty((q.rows, q.m), order="F")
        chg = np.empty(q.n_ptab, dtype=np.int32)
        err = C.create_string_buffer(2048)
        rc = self.lib.psy_ref_fit(C.byref(q), int(skip_update), _dp(m2), _dp(lat), _dp(pred), chg.ctypes.data_as(C.POINTER(C.c_int)), err, 2048)
        if rc:
            raise RuntimeError(err.value.decode())
        return {"m2LL": m2, "latentScores": np.ascontiguousarray(lat), "predictedManifest": np.ascontiguousarray(pred), "changed": chg.astype(bool)}

    def gradients(self, theta=None):
        """getGradients(psydiffModel): dict label -> value, in getParameterValues order."""
        q = self._problem(theta)
        cap = len(self.model["pars"]["parameterTable"].theta_labels) + 8
        g = np.empty(cap)
        n = C.c_int(0)
        names = C.create_string_buffer(64 * cap + 64)
        err = C.create_string_buffer(2048)
        rc = self.lib.psy_ref_gradients(C.byref(q), _dp(g), cap, C.byref(n), names, len(names), err, 2048)
        if rc:
            raise RuntimeError(err.value.decode())
        return dict(zip(names.value.decode().split("\n"), g[:n.value].tolist()))

    def gradient(self, par: int, theta=None):
        q = self._problem(theta)
        g = C.c_double(0)
        err = C.create_string_buffer(2048)
        rc = self.lib.psy_ref_gradient(C.byref(q), int(par), C.byref(g), err, 2048)
        if rc:
            raise RuntimeError(err.value.decode())
        return g.value

    def primitive(self, name, a=None, b=None, c=None, scalars=(), text="", out_shape=(1,)):
        """One function of inst/include/psydiffInternals.h on column-major buffers (see psy_ref_primitive)."""
        def prep(x):
            if x is None:
                return None, 0, 0, None
            x = np.asarray(x, dtype=np.float64)
            x2 = x.reshape(-1, 1) if x.ndim == 1 else x
            f = np.ascontiguousarray(x2.reshape(-1, order="F"))
            return _dp(f), x2.shape[0], x2.shape[1], f
        pa, ar, ac, ka = prep(a)
        pb, br, bc, kb = prep(b)
        pc, cr, cc, kc = prep(c)
        s = np.ascontiguousarray(list(scalars) + [0.0] * 4, dtype=np.float64)
        out = np.empty(int(np.prod(out_shape)))
        err = C.create_string_buffer(2048)
        rc = self.lib.psy_ref_primitive(name.encode(), pa, ar, ac, pb, br, bc, pc, cr, cc, _dp(s), text.encode(), _dp(out), err, 2048)
        if rc:
            raise RuntimeError(err.value.decode())
        return out.reshape(out_shape, order="F") if len(out_shape) > 1 else out
