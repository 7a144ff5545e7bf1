"""r/psydiff_b200_shim.c — the .Call layer of the R drop-in — compiled against tests/r_stub (a functional stand-in for the R C
API; R itself is not in this image) and EXECUTED: the registration table against the reference's (src/RcppExports.cpp:292-316),
every primitive, the by-reference parameter plumbing (inst/include/psydiffBasic.h), the source generator entry point, and —
on a GPU — compile / fit / gradients through the same SEXP interface an R session would use.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import synth, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")
REALSXP, INTSXP, LGLSXP, STRSXP, VECSXP = 14, 13, 10, 16, 19

# src/RcppExports.cpp:292-316
REFERENCE_TABLE = {"_psydiff_setParameterValues": 3, "_psydiff_getParameterValues": 1, "_psydiff_setParameterList": 3,
                   "_psydiff_getSigmaPoints": 4, "_psydiff_getMeanWeights": 4, "_psydiff_getCovWeights": 4, "_psydiff_getWMatrix": 2,
                   "_psydiff_computeMeanFromSigmaPoints": 2, "_psydiff_getAFromSigmaPoints": 3, "_psydiff_computePFromSigmaPoints": 3,
                   "_psydiff_getPhi": 1, "_psydiff_predictMu": 2, "_psydiff_predictS": 3, "_psydiff_predictC": 3, "_psydiff_computeK": 2,
                   "_psydiff_updateM": 3, "_psydiff_updateP": 3, "_psydiff_computeIndividualM2LL": 4, "_psydiff_cholupdate": 4,
                   "_psydiff_qr_": 1, "_psydiff_logChol2Chol": 1, "_psydiff_clonePsydiffModel": 1}


class R:
    """Just enough of an R session: builds stand-in SEXPs from Python values, .Call()s registered routines, reads results."""

    def __init__(self, mock_device=False):
        """mock_device=True links the shim against tests/r_stub/mock_device.c ahead of the real library: the device entry points
        (compile, upload, fit, gradients) become trivial host stand-ins so the whole flow runs without a GPU."""
        _lib.lib()                                         # libpsydiff_b200.so built (conftest) before linking against it
        out = os.path.join(STUB, "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libpsydiff_rshim_mock.so" if mock_device else "libpsydiff_rshim.so")
        srcs = [os.path.join(ROOT, "r", "psydiff_b200_shim.c"), os.path.join(STUB, "fake_r.c")]
        if mock_device:
            srcs.append(os.path.join(STUB, "mock_device.c"))   # same link unit: its definitions of the device entry points win
        deps = srcs + [os.path.join(STUB, h) for h in ("Rinternals.h", "R.h", "R_ext/Rdynload.h")] + \
            [os.path.join(ROOT, "include", "psydiff_b200.h"), _lib.LIB_PATH]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            libdir = os.path.dirname(_lib.LIB_PATH)
            cmd = ["gcc", "-std=gnu11", "-O1", "-g", "-shared", "-fPIC", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-I", STUB, "-I",
                   os.path.join(ROOT, "include"), "-o", so] + srcs + ["-L", libdir, "-lpsydiff_b200", f"-Wl,-rpath,{libdir}"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
        L = self.L = C.CDLL(so)
        for name, res, args in (("Rf_allocVector", C.c_void_p, [C.c_uint, C.c_ssize_t]), ("Rf_allocMatrix", C.c_void_p, [C.c_uint, C.c_int, C.c_int]),
                                ("Rf_mkChar", C.c_void_p, [C.c_char_p]), ("SET_STRING_ELT", None, [C.c_void_p, C.c_ssize_t, C.c_void_p]),
                                ("SET_VECTOR_ELT", C.c_void_p, [C.c_void_p, C.c_ssize_t, C.c_void_p]), ("STRING_ELT", C.c_void_p, [C.c_void_p, C.c_ssize_t]),
                                ("VECTOR_ELT", C.c_void_p, [C.c_void_p, C.c_ssize_t]), ("CHAR", C.c_char_p, [C.c_void_p]),
                                ("REAL", C.POINTER(C.c_double), [C.c_void_p]), ("INTEGER", C.POINTER(C.c_int), [C.c_void_p]),
                                ("LOGICAL", C.POINTER(C.c_int), [C.c_void_p]), ("TYPEOF", C.c_int, [C.c_void_p]), ("LENGTH", C.c_int, [C.c_void_p]),
                                ("Rf_nrows", C.c_int, [C.c_void_p]), ("Rf_ncols", C.c_int, [C.c_void_p]),
                                ("Rf_getAttrib", C.c_void_p, [C.c_void_p, C.c_void_p]), ("Rf_setAttrib", C.c_void_p, [C.c_void_p, C.c_void_p, C.c_void_p]),
                                ("R_ExternalPtrAddr", C.c_void_p, [C.c_void_p]),
                                ("psy_rstub_call", C.c_void_p, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
                                ("psy_rstub_error", C.c_char_p, []), ("psy_rstub_warning", C.c_char_p, []), ("psy_rstub_warnings", C.c_int, []),
                                ("psy_rstub_n_registered", C.c_int, []), ("psy_rstub_registered_name", C.c_char_p, [C.c_int]),
                                ("psy_rstub_registered_arity", C.c_int, [C.c_int]), ("psy_rstub_dynamic_symbols", C.c_int, []),
                                ("psy_rstub_run_finalizer", None, [C.c_void_p]), ("R_init_psydiff", None, [C.c_void_p])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.names_sym = C.c_void_p.in_dll(L, "R_NamesSymbol").value
        self.nil = C.c_void_p.in_dll(L, "R_NilValue").value
        L.R_init_psydiff(None)                              # what R does when the package's shared object is loaded

    # ---- Python -> SEXP ------------------------------------------------------------------------------------------
    def num(self, a):
        a = np.asarray(a, dtype=np.float64)
        s = self.L.Rf_allocMatrix(REALSXP, a.shape[0], a.shape[1]) if a.ndim == 2 else self.L.Rf_allocVector(REALSXP, a.size)
        flat = np.asfortranarray(a).reshape(-1, order="F")
        C.memmove(self.L.REAL(s), flat.ctypes.data, flat.nbytes)
        return s

    def int(self, a, logical=False):
        a = np.asarray(a, dtype=np.int32)
        t = LGLSXP if logical else INTSXP
        s = self.L.Rf_allocMatrix(t, a.shape[0], a.shape[1]) if a.ndim == 2 else self.L.Rf_allocVector(t, a.size)
        flat = np.asfortranarray(a).reshape(-1, order="F")
        C.memmove(self.L.INTEGER(s), flat.ctypes.data, flat.nbytes)
        return s

    def lgl(self, a):
        return self.int(np.atleast_1d(np.asarray(a, dtype=bool)).astype(np.int32), logical=True)

    def chr(self, strings):
        strings = [strings] if isinstance(strings, str) else list(strings)
        s = self.L.Rf_allocVector(STRSXP, len(strings))
        for i, v in enumerate(strings):
            self.L.SET_STRING_ELT(s, i, self.L.Rf_mkChar(v.encode()))
        return s

    def list(self, items, names=None):
        items = list(items.items()) if isinstance(items, dict) else [(None, v) for v in items]
        s = self.L.Rf_allocVector(VECSXP, len(items))
        for i, (_, v) in enumerate(items):
            self.L.SET_VECTOR_ELT(s, i, v)
        if items and items[0][0] is not None:
            self.L.Rf_setAttrib(s, self.names_sym, self.chr([k for k, _ in items]))
        return s

    # ---- SEXP -> Python ------------------------------------------------------------------------------------------
    def py(self, s):
        L = self.L
        t, n = L.TYPEOF(s), L.LENGTH(s)
        if t == REALSXP:
            v = np.ctypeslib.as_array(L.REAL(s), shape=(max(n, 1),))[:n].copy()
        elif t in (INTSXP, LGLSXP):
            v = np.ctypeslib.as_array(L.INTEGER(s), shape=(max(n, 1),))[:n].copy()
            v = v.astype(bool) if t == LGLSXP else v
        elif t == STRSXP:
            return [L.CHAR(L.STRING_ELT(s, i)).decode() for i in range(n)]
        elif t == VECSXP:
            vals = [self.py(L.VECTOR_ELT(s, i)) for i in range(n)]
            nm = L.Rf_getAttrib(s, self.names_sym)
            return dict(zip(self.py(nm), vals)) if nm != self.nil else vals
        else:
            return s
        if L.Rf_ncols(s) > 1 or L.Rf_nrows(s) != n:
            v = v.reshape((L.Rf_nrows(s), L.Rf_ncols(s)), order="F")
        return v

    def names(self, s):
        return self.py(self.L.Rf_getAttrib(s, self.names_sym))

    def call(self, name, *args):
        arr = (C.c_void_p * max(len(args), 1))(*args)
        out = self.L.psy_rstub_call(name.encode(), len(args), arr)
        if not out:
            raise RuntimeError(self.L.psy_rstub_error().decode())
        return out


@pytest.fixture(scope="module")
def r():
    return R()


def test_registration_table_matches_the_reference(r):
    got = {r.L.psy_rstub_registered_name(k).decode(): r.L.psy_rstub_registered_arity(k) for k in range(r.L.psy_rstub_n_registered())}
    for name, arity in REFERENCE_TABLE.items():
        assert got.get(name) == arity, name
    assert set(got) - set(REFERENCE_TABLE) == {"R_psy_emit_source", "R_psy_compile", "R_psy_is_current", "R_psy_upload", "R_psy_load_module",
                                               "R_psy_set_devices", "R_psy_fit", "R_psy_gradients", "R_psy_gradient"}
    assert r.L.psy_rstub_dynamic_symbols() == 0          # R_useDynamicSymbols(dll, FALSE), src/RcppExports.cpp:320
    with pytest.raises(RuntimeError, match="Incorrect number of arguments"):
        r.call("_psydiff_getPhi", r.num(np.eye(2)), r.num([1.0]))
    with pytest.raises(RuntimeError, match="not in load table"):
        r.call("_psydiff_rcpparma_hello_world")


def test_primitives_through_dot_call(r):
    rng = np.random.default_rng(3)
    n, m = 3, 2
    s = 2 * n + 1
    alpha, beta, kappa = 0.2, 2.0, 0.0
    lam = alpha ** 2 * (n + kappa) - n
    wm = r.py(r.call("_psydiff_getMeanWeights", r.int([n]), r.num([alpha]), r.num([beta]), r.num([kappa])))
    wc = r.py(r.call("_psydiff_getCovWeights", r.int([n]), r.num([alpha]), r.num([beta]), r.num([kappa])))
    assert np.allclose(wm, [lam / (n + lam)] + [1 / (2 * (n + lam))] * (2 * n)) and np.isclose(wc[0], wm[0] + 1 - alpha ** 2 + beta)
    W = r.py(r.call("_psydiff_getWMatrix", r.num(wm), r.num(wc)))
    IW = np.eye(s) - np.outer(wm, np.ones(s))
    assert W.shape == (s, s) and np.allclose(W, IW @ np.diag(wc) @ IW.T)
    A = np.tril(rng.standard_normal((n, n))) + 2 * np.eye(n)
    mu = rng.standard_normal(n)
    c = n + lam
    X = r.py(r.call("_psydiff_getSigmaPoints", r.num(mu), r.num(A), r.lgl(True), r.num([c])))
    assert np.allclose(X, np.column_stack([mu, mu[:, None] + np.sqrt(c) * A, mu[:, None] - np.sqrt(c) * A]))
    X2 = r.py(r.call("_psydiff_getSigmaPoints", r.num(mu), r.num(A @ A.T), r.lgl(False), r.num([c])))
    assert np.allclose(X2, X)
    with pytest.raises(RuntimeError, match="decomposition failed"):
        r.call("_psydiff_getSigmaPoints", r.num(mu), r.num(-np.eye(n)), r.lgl(False), r.num([c]))
    assert np.allclose(r.py(r.call("_psydiff_computeMeanFromSigmaPoints", r.num(X), r.num(wm))).ravel(), X @ wm)
    assert np.allclose(r.py(r.call("_psydiff_getAFromSigmaPoints", r.num(X), r.num(wm), r.num([c]))), A)
    assert np.allclose(r.py(r.call("_psydiff_computePFromSigmaPoints", r.num(X), r.num(wm), r.num([c]))), A @ A.T)
    Mx = rng.standard_normal((n, n))
    assert np.allclose(r.py(r.call("_psydiff_getPhi", r.num(Mx))), np.tril(Mx, -1) + 0.5 * np.diag(np.diag(Mx)))
    Y = rng.standard_normal((m, s))
    Rm = np.array([[0.5, 0.1], [0.1, 0.4]])
    assert np.allclose(r.py(r.call("_psydiff_predictMu", r.num(Y), r.num(wm))).ravel(), Y @ wm)
    S = r.py(r.call("_psydiff_predictS", r.num(Y), r.num(W), r.num(Rm)))
    assert np.allclose(S, Y @ W @ Y.T + Rm)
    Cm = r.py(r.call("_psydiff_predictC", r.num(X), r.num(Y), r.num(W)))
    assert np.allclose(Cm, X @ W @ Y.T)
    Sp = Rm + np.eye(m)
    K = r.py(r.call("_psydiff_computeK", r.num(Cm), r.num(Sp)))
    assert np.allclose(K, Cm @ np.linalg.inv(Sp))
    res = rng.standard_normal(m)
    assert np.allclose(r.py(r.call("_psydiff_updateM", r.num(mu), r.num(K), r.num(res))).ravel(), mu + K @ res)
    P = A @ A.T
    assert np.allclose(r.py(r.call("_psydiff_updateP", r.num(P), r.num(K), r.num(Sp))), P - K @ Sp @ K.T)
    raw, em = rng.standard_normal(m), rng.standard_normal(m)
    ll = r.py(r.call("_psydiff_computeIndividualM2LL", r.int([m]), r.num(raw), r.num(em), r.num(Sp)))[0]
    d = raw - em
    assert np.isclose(ll, m * np.log(2 * np.pi) + np.log(np.linalg.det(Sp)) + d @ np.linalg.solve(Sp, d))
    Lc = np.linalg.cholesky(P)
    x = rng.standard_normal(n)
    up = r.py(r.call("_psydiff_cholupdate", r.num(Lc), r.num(x), r.num([1.0]), r.chr("update")))
    assert np.allclose(up @ up.T, P + np.outer(x, x)) and np.allclose(np.triu(up, 1), 0)
    with pytest.raises(RuntimeError, match="Unknown direction in cholupdate"):
        r.call("_psydiff_cholupdate", r.num(Lc), r.num(x), r.num([1.0]), r.chr("sideways"))
    T = rng.standard_normal((7, 3))
    Rq = r.py(r.call("_psydiff_qr_", r.num(T)))
    assert np.allclose(Rq.T @ Rq, T.T @ T) and np.allclose(np.tril(Rq, -1), 0)
    lc = np.array([[0.3, 0.0], [0.2, -0.4]])
    assert np.allclose(r.py(r.call("_psydiff_logChol2Chol", r.num(lc))), [[np.exp(0.3), 0], [0.2, np.exp(-0.4)]])


def _parameter_table(r, model):
    """model$pars$parameterTable as the data.frame newPsydiff builds (R/prepareModel.R:471-476)."""
    pt = model["pars"]["parameterTable"]
    return r.list({"label": r.chr(list(pt.label)), "value": r.num(np.asarray(pt.value, dtype=float)), "target": r.chr(list(pt.target)),
                   "targetClass": r.chr(["double" if c.kind == "double" else "matrix" for c in
                                         (model["_spec"].container(t) for t in pt.target)]),
                   "row": r.num([float(x["row"]) for x in pt.slot_rows] * len(pt.persons)),
                   "col": r.num([float(x["col"]) for x in pt.slot_rows] * len(pt.persons)),
                   "changed": r.lgl(np.ones(len(pt), dtype=bool)),
                   "person": r.num(np.asarray(pt.person, dtype=float))})


def test_parameter_plumbing_by_reference(r):
    model = pd.newPsydiff(**synth.config_c4(N=3, T=4, n_groups=2))
    ptab = _parameter_table(r, model)
    rmodel = r.list({"pars": r.list({"parameterTable": ptab})})
    theta = r.call("_psydiff_getParameterValues", rmodel)
    labels = r.names(theta)
    assert labels == model["pars"]["parameterTable"].theta_labels            # first-appearance order (quirk Q6)
    assert np.array_equal(r.py(theta), model["pars"]["parameterTable"].theta_values)
    # setParameterValues mutates the data.frame it is given and raises `changed` only where the value differs
    C.memset(r.L.LOGICAL(r.L.VECTOR_ELT(ptab, 6)), 0, 4 * len(model["pars"]["parameterTable"]))
    r.call("_psydiff_setParameterValues", ptab, r.num([0.123, 1.0]), r.chr(["mu1_p2", "om1"]))      # om1 already is 1.0
    df = r.py(ptab)
    hit = np.array(df["label"]) == "mu1_p2"
    assert hit.sum() == 1 and np.all(df["value"][hit] == 0.123) and np.array_equal(df["changed"], hit)
    assert r.py(r.call("_psydiff_getParameterValues", rmodel))[labels.index("mu1_p2")] == 0.123
    w0 = r.L.psy_rstub_warnings()
    r.call("_psydiff_setParameterValues", ptab, r.num([1.0]), r.chr(["nosuch"]))
    assert r.L.psy_rstub_warnings() == w0 + 1 and b"not found in the parameterTable: nosuch" in r.L.psy_rstub_warning()
    with pytest.raises(RuntimeError, match="Length of parameterValues and parameterLabels does not match"):
        r.call("_psydiff_setParameterValues", ptab, r.num([1.0, 2.0]), r.chr(["om1"]))
    # setParameterList copies the person's CHANGED rows into the shared list at (target, row, col)
    plist = r.list({c.name: (r.num([c.values[0, 0]]) if c.kind == "double" else r.num(c.values)) for c in model["_spec"].containers})
    person2 = int(model["_persons_unique"][1])
    r.call("_psydiff_setParameterList", ptab, plist, r.int([person2]))
    after = r.py(plist)
    assert after["mu1"][0] == 0.123                                       # the one changed row of person 2
    assert after["om1"][0] == 1.0 and np.array_equal(after["LLogChol"], model["_spec"].container("LLogChol").values)
    with pytest.raises(RuntimeError, match="Person 99 not found in data set."):
        r.call("_psydiff_setParameterList", ptab, plist, r.int([99]))
    # quirk Q5: extractParameters tags vectors "colvec", which setParameterList rejects (psydiffBasic.h:102-111)
    r.L.SET_STRING_ELT(r.L.VECTOR_ELT(ptab, 3), int(np.flatnonzero(hit)[0]), r.L.Rf_mkChar(b"colvec"))
    with pytest.raises(RuntimeError, match="Parameter class not found"):
        r.call("_psydiff_setParameterList", ptab, plist, r.int([person2]))
    # clonePsydiffModel is a deep copy
    clone = r.call("_psydiff_clonePsydiffModel", rmodel)
    r.call("_psydiff_setParameterValues", ptab, r.num([9.0]), r.chr(["om1"]))
    assert r.py(r.call("_psydiff_getParameterValues", clone))[labels.index("om1")] == 1.0
    assert r.py(r.call("_psydiff_getParameterValues", rmodel))[labels.index("om1")] == 9.0


def _emit_args(r, model):
    """The arguments psydiffCudaSource (r/psydiff_b200.R) passes: names, scalar flags, value matrices, slot matrices."""
    spec = model["_spec"]
    return (r.int([spec.nlatent]), r.int([spec.nmanifest]), r.int([spec.n_slots]), r.chr(spec.latentEquations), r.chr(spec.manifestEquations),
            r.chr([c.name for c in spec.containers]), r.lgl([c.kind == "double" for c in spec.containers]),
            r.list([r.num(c.values) for c in spec.containers]), r.list([r.int(np.asarray(c.slots, dtype=np.int32)) for c in spec.containers]),
            r.lgl(spec.srUpdate), r.int([spec.stepper]))


@pytest.mark.parametrize("kw", [synth.config_c1(N=3, T=4), synth.config_c4(N=3, T=4), synth.config_c3(N=3, T=4),
                                synth.config_c1(N=3, T=4, srUpdate=False)], ids=["c1", "c4", "c3-dopri5", "c1-cov"])
def test_source_generation_through_the_shim_equals_the_python_front_end(r, kw):
    model = pd.newPsydiff(**kw)
    src = r.py(r.call("R_psy_emit_source", *_emit_args(r, model)))
    assert src == [model["cppCode"]]


def test_source_generation_errors_become_r_errors(r):
    kw = synth.config_c1(N=3, T=4)
    model = pd.newPsydiff(**kw)
    args = list(_emit_args(r, model))
    args[0] = r.int([3])                                                    # nlatent that does not match A0 / L
    with pytest.raises(RuntimeError, match="A0 and L must be nlatent x nlatent"):
        r.call("R_psy_emit_source", *args)


def test_compute_entry_points_fail_loudly_without_a_device(r):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    model = pd.newPsydiff(**synth.config_c1(N=3, T=4))
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        _compile(r, model)
    with pytest.raises(RuntimeError, match="model is not compiled"):
        r.call("R_psy_gradients", r.L.Rf_allocVector(VECSXP, 0), r.num([.2, 2, 0]), r.int([0]), r.num([1 / 30]), r.lgl(True),
               r.num([0.0]), r.num([1e-6]), r.int([0]))


def _compile(r, model):
    """compileModel of r/psydiff_b200.R: compile + load only; the data follow through .psy_sync."""
    spec = model["_spec"]
    return r.call("R_psy_compile", r.chr(model["cppCode"]), r.chr(_lib.CSRC_DIR), r.chr(_lib.CACHE_DIR), r.int([spec.nlatent]),
                  r.int([spec.nmanifest]), r.int([spec.n_slots]))


def _r_model(r, model):
    """The psydiffModel list as R holds it (the fields the wrappers and the shim read)."""
    obs = np.asfortranarray(model["data"]["observations"])
    N, rows = len(model["_persons_unique"]), obs.shape[0]
    return r.list({"pars": r.list({"parameterTable": _parameter_table(r, model)}),
                   "nlatent": r.int([model["nlatent"]]), "nmanifest": r.int([model["nmanifest"]]),
                   "data": r.list({"person": r.num(np.asarray(model["data"]["person"], dtype=float)), "observations": r.num(obs),
                                   "dt": r.num(model["data"]["dt"])}),
                   "m2LL": r.num(np.zeros(N)), "predictedManifest": r.num(np.full((rows, model["nmanifest"]), -99999.99)),
                   "latentScores": r.num(np.full((rows, model["nlatent"]), -99999.99))})


def _sync(r, h, rm, model):
    """.psy_sync of r/psydiff_b200.R; returns whether an upload happened."""
    L = r.L
    data, pt = L.VECTOR_ELT(rm, 3), L.VECTOR_ELT(L.VECTOR_ELT(rm, 0), 0)
    person, obs, dt, labels = L.VECTOR_ELT(data, 0), L.VECTOR_ELT(data, 1), L.VECTOR_ELT(data, 2), L.VECTOR_ELT(pt, 0)
    if r.py(r.call("R_psy_is_current", h, obs, dt, person, labels))[0]:
        return False
    ptab = model["pars"]["parameterTable"]
    r.call("R_psy_upload", h, obs, dt, person, r.num(np.asarray(model["_row_off"], dtype=float)),
           r.int(np.asarray(model["_persons_unique"], dtype=np.int32)), r.int([len(ptab.theta_labels)]),
           r.int(np.ascontiguousarray(ptab.theta_index, dtype=np.int32).reshape(-1)), labels)
    return True


def test_compile_sync_fit_gradient_flow_with_a_mock_device():
    """The whole R-side flow of r/psydiff_b200.R through the shim, the device entry points replaced by tests/r_stub/mock_device.c:
    compile -> sync (upload once per dataset object) -> fit with the reference's by-reference side effects -> named gradients ->
    a second model object of the same structure -> finalizer."""
    r = R(mock_device=True)
    L = r.L
    L.psy_mock_live_models.restype = C.c_int
    model = pd.newPsydiff(**synth.config_c4(N=6, T=8))
    h = _compile(r, model)
    assert L.psy_mock_live_models() == 1
    rm = _r_model(r, model)
    settings = (r.num([model["alpha"], model["beta"], model["kappa"]]), r.int([0]), r.num(model["timeStep"]), r.lgl(True))
    theta = r.call("_psydiff_getParameterValues", rm)
    with pytest.raises(RuntimeError, match="psy_upload|psy_set_slots"):
        r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(False))
    assert _sync(r, h, rm, model) is True and _sync(r, h, rm, model) is False
    obs = model["data"]["observations"]
    th = model["pars"]["parameterTable"].theta_values
    want = np.nansum(obs) + np.arange(6) + th.sum()                                    # the mock's arithmetic
    out = r.py(r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(False)))
    assert list(out) == ["latentScores", "predictedManifest", "m2LL"] and np.allclose(out["m2LL"], want, rtol=1e-13)
    assert out["latentScores"].shape == (48, 6) and out["latentScores"][1, 0] == 1001.0 and out["latentScores"][0, 1] == 1048.0   # column-major
    after = r.py(rm)
    assert np.array_equal(after["m2LL"], out["m2LL"]) and np.array_equal(after["predictedManifest"], out["predictedManifest"])
    assert not after["pars"]["parameterTable"]["changed"].any()
    assert np.all(r.py(r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(True)))["m2LL"] == 0.0)       # skipUpdate reaches the library
    gr = r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([2]))
    assert r.names(gr) == model["pars"]["parameterTable"].theta_labels
    assert np.allclose(r.py(gr), 2 * th + 6 + 0.002 + model["eps"][0])
    one = r.call("R_psy_gradient", h, *settings, theta, r.int([4]), r.num(model["eps"]), r.int([2]))
    assert r.py(one)[0] == r.py(gr)[4] and r.names(one) == [model["pars"]["parameterTable"].theta_labels[4]]
    with pytest.raises(RuntimeError, match="Unknown direction"):
        r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([7]))
    with pytest.raises(RuntimeError, match="parameter index out of range"):
        r.call("R_psy_gradient", h, *settings, theta, r.int([999]), r.num(model["eps"]), r.int([0]))
    # compileParallelGradients(model, nCores): the workers are devices of the handle — capped at the visible ones (4 in the mock)
    assert list(r.py(r.call("R_psy_set_devices", h, r.int([2])))) == [0, 1]
    assert list(r.py(r.call("R_psy_set_devices", h, r.int([64])))) == [0, 1, 2, 3]
    assert list(r.py(r.call("R_psy_set_devices", h, r.int([0])))) == [0]                  # stopParallelGradients: back to one
    # a changed integrateFunction / srUpdate: .psy_variant_sync hands the regenerated source to R_psy_load_module
    L.psy_mock_module_loads.restype = C.c_int
    L.psy_mock_module_loads.argtypes = [C.c_void_p]
    model["integrateFunction"] = "runge_kutta_dopri5"
    pd.model._sync_source(model)
    r.call("R_psy_load_module", h, r.chr(model["cppCode"]), r.chr(_lib.CSRC_DIR), r.chr(_lib.CACHE_DIR))
    assert L.psy_mock_module_loads(r.L.R_ExternalPtrAddr(h)) == 1
    # a different model object of the same structure: re-upload, results follow the new data; handing the first one back re-uploads again
    model2 = pd.newPsydiff(**synth.config_c4(N=4, T=6, seed=77))
    rm2 = _r_model(r, model2)
    assert _sync(r, h, rm2, model2) is True
    out2 = r.py(r.call("R_psy_fit", h, rm2, *settings, r.call("_psydiff_getParameterValues", rm2), r.lgl(False)))
    assert np.allclose(out2["m2LL"], np.nansum(model2["data"]["observations"]) + np.arange(4) + model2["pars"]["parameterTable"].theta_values.sum(), rtol=1e-13)
    assert _sync(r, h, rm, model) is True and _sync(r, h, rm, model) is False
    r.L.psy_rstub_run_finalizer(h)
    assert L.psy_mock_live_models() == 0 and not r.L.R_ExternalPtrAddr(h)
    with pytest.raises(RuntimeError, match="model is not compiled"):
        r.call("R_psy_is_current", h, r.num([0.0]), r.num([0.0]), r.num([0.0]), r.chr(["a"]))


@pytest.mark.gpu
def test_compile_fit_and_gradients_through_the_shim_on_the_gpu(r):
    model = pd.newPsydiff(**synth.config_c4(N=6, T=8))
    pd.compileModel(model)
    fit = pd.fitModel(model)
    g = pd.getGradients(model)
    h = _compile(r, model)
    rm = _r_model(r, model)
    settings = (r.num([model["alpha"], model["beta"], model["kappa"]]), r.int([0]), r.num(model["timeStep"]), r.lgl(True))
    theta = r.call("_psydiff_getParameterValues", rm)
    with pytest.raises(RuntimeError, match="psy_upload|psy_set_slots"):               # nothing resident yet
        r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(False))
    assert _sync(r, h, rm, model) is True and _sync(r, h, rm, model) is False          # uploads once for the same R objects
    out = r.py(r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(False)))
    assert list(out) == ["latentScores", "predictedManifest", "m2LL"]                   # R/modelTemplate.R:416-419
    assert np.array_equal(out["m2LL"], fit["m2LL"]) and np.array_equal(out["latentScores"], fit["latentScores"])
    assert np.array_equal(out["predictedManifest"], fit["predictedManifest"])
    after = r.py(rm)                                                                    # by-reference side effects on the model
    assert np.array_equal(after["m2LL"], fit["m2LL"]) and np.array_equal(after["latentScores"], fit["latentScores"])
    assert not after["pars"]["parameterTable"]["changed"].any()
    gr = r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([0]))
    assert r.names(gr) == list(g) and np.array_equal(r.py(gr), np.array(list(g.values())))
    one = r.call("R_psy_gradient", h, *settings, theta, r.int([4]), r.num(model["eps"]), r.int([0]))
    assert r.py(one)[0] == r.py(gr)[4] and r.names(one) == [list(g)[4]]
    with pytest.raises(RuntimeError, match="Unknown direction"):
        r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([7]))
    # a different dataset object of the same structure: the wrappers re-upload and the functions work on it (as the reference's do)
    model2 = pd.newPsydiff(**synth.config_c4(N=4, T=6, seed=77))
    pd.compileModel(model2)
    rm2 = _r_model(r, model2)
    assert _sync(r, h, rm2, model2) is True
    out2 = r.py(r.call("R_psy_fit", h, rm2, *settings[:2], r.num(model2["timeStep"]), settings[3], r.call("_psydiff_getParameterValues", rm2), r.lgl(False)))
    assert np.array_equal(out2["m2LL"], pd.fitModel(model2)["m2LL"])
    r.L.psy_rstub_run_finalizer(h)                                                      # what R's GC does with the external pointer
    assert not r.L.R_ExternalPtrAddr(h)
    with pytest.raises(RuntimeError, match="model is not compiled"):
        r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([0]))


@pytest.mark.gpu
def test_get_gradients_parallel_through_the_shim_uses_several_gpus(r):
    """compileParallelGradients / getGradientsParallel of r/psydiff_b200.R on a box with >= 2 GPUs: R_psy_set_devices spreads the
    handle's persons over the devices; fit, states and gradient through the .Call interface equal the one-device results
    (per-person values exactly, parameter sums to 1e-12) — no Python, no process per GPU."""
    if pd.deviceCount() < 2:
        pytest.skip("needs at least two GPUs")
    model = pd.newPsydiff(**synth.config_c4(N=301, T=12))
    h = _compile(r, model)
    rm = _r_model(r, model)
    settings = (r.num([model["alpha"], model["beta"], model["kappa"]]), r.int([0]), r.num(model["timeStep"]), r.lgl(True))
    theta = r.call("_psydiff_getParameterValues", rm)
    assert _sync(r, h, rm, model) is True
    one = r.py(r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(False)))
    g1 = r.py(r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([0])))
    n = min(pd.deviceCount(), 8)
    assert list(r.py(r.call("R_psy_set_devices", h, r.int([n])))) == list(range(n))
    many = r.py(r.call("R_psy_fit", h, rm, *settings, theta, r.lgl(False)))
    gn = r.py(r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([0])))
    for k in ("m2LL", "latentScores", "predictedManifest"):
        assert np.array_equal(one[k], many[k]), k
    assert np.max(np.abs(gn - g1) / np.maximum(np.abs(g1), 1e-300)) < 1e-12
    assert list(r.py(r.call("R_psy_set_devices", h, r.int([1])))) == [0]
    assert np.array_equal(r.py(r.call("R_psy_gradients", h, *settings, theta, r.num(model["eps"]), r.int([0]))), g1)
    r.L.psy_rstub_run_finalizer(h)
