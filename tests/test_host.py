"""Host logic and the C-ABI boundary without a GPU: symbol export, loud failure without a device, the model front end
(newPsydiff parsing, log-Cholesky, grouping -> slot map), code generation + nvcc cross-compile, sweep bookkeeping,
and the exported filter primitives."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import _lib, codegen, synth
from psydiff_b200._lib import PsydiffError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    from psydiff_b200 import build
    build.build_library()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "psydiff_b200.h")).read()
    declared = set(re.findall(r"\b(psy_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 40
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/psydiff_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_compute_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    model = pd.newPsydiff(**synth.config_c1(N=3, T=5))
    with pytest.raises(PsydiffError, match="no CUDA device|CPU fallback"):
        pd.compileModel(model)
    with pytest.raises(PsydiffError, match="not compiled"):
        pd.fitModel(model)
    tf = C.c_double()
    assert L.psy_measure_fp64_peak(C.byref(tf)) == 2   # PSY_ERR_CUDA


def test_newpsydiff_parameter_table_and_logchol():
    m = pd.newPsydiff(**synth.config_c1(N=4, T=6))
    pt = m["pars"]["parameterTable"]
    assert pt.theta_labels == ["drift_eta1", "drift_eta1_eta2", "drift_eta2_eta1", "drift_eta2", "mm_Y1", "mm_Y2",
                               "lnT0var_eta1", "T0var_eta2_eta1", "lnT0var_eta2", "lneta1_eta1", "lneta2_eta2"]
    v = pd.getParameterValues(m)
    assert v["lneta1_eta1"] == pytest.approx(np.log(np.sqrt(.5)))     # chol2LogChol: log on the diagonal, "ln" prefix
    assert v["lnT0var_eta1"] == 0.0 and v["T0var_eta2_eta1"] == 0.0
    assert pt.theta_index.shape == (4, 11) and (pt.theta_index == np.arange(11)).all()
    assert len(pt) == 44 and list(pt.label[:2]) == ["drift_eta1", "drift_eta1_eta2"]
    assert m["timeStep"][0] == pytest.approx(1 / 30)
    assert m["m2LL"].shape == (4,) and (m["latentScores"] == -99999.99).all()
    # default timeStep (R/prepareModel.R:435-438)
    kw = synth.config_c1(N=2, T=4)
    kw.pop("timeStep")
    assert np.allclose(pd.newPsydiff(**kw)["timeStep"], np.linspace(1 / 30, 1 / 10, 5))


def test_grouping_person_and_group_labels():
    kw = synth.config_c1(N=6, T=4)
    kw["grouping"] = "mm_Y1 | person;\n mm_Y2 | group1;"
    kw["groupingvariables"] = {"group1": np.array([1, 1, 1, 2, 2, 2])}
    m = pd.newPsydiff(**kw)
    pt = m["pars"]["parameterTable"]
    labs = pt.theta_labels
    assert "mm_Y1_p1" in labs and "mm_Y1_p6" in labs and "mm_Y2_p1" in labs and "mm_Y2_p2" in labs and "mm_Y1" not in labs
    assert len(labs) == 9 + 6 + 2
    # first-appearance order: person 1's rows first
    assert labs[:6] == ["drift_eta1", "drift_eta1_eta2", "drift_eta2_eta1", "drift_eta2", "mm_Y1_p1", "mm_Y2_p1"]
    k1, k2 = 4, 5
    assert [labs[i] for i in pt.theta_index[:, k1]] == [f"mm_Y1_p{p}" for p in range(1, 7)]
    assert [labs[i] for i in pt.theta_index[:, k2]] == ["mm_Y2_p1"] * 3 + ["mm_Y2_p2"] * 3
    with pytest.raises(PsydiffError, match="not found in groupingvariables"):
        pd.newPsydiff(**dict(kw, grouping="mm_Y1 | nope;"))


def test_newpsydiff_errors_mirror_reference():
    kw = synth.config_c1(N=2, T=4)
    with pytest.raises(PsydiffError, match="to end with ;"):
        pd.newPsydiff(**dict(kw, latentEquations="dx = DRIFT*x"))
    with pytest.raises(PsydiffError, match="Unknown field in dataset"):
        pd.newPsydiff(**dict(kw, dataset=dict(kw["dataset"], extra=1)))
    with pytest.raises(PsydiffError, match="dataset must be a list"):
        pd.newPsydiff(**dict(kw, dataset=[1, 2]))
    with pytest.raises(PsydiffError, match="differ in their number of manifest"):
        pd.newPsydiff(**dict(kw, Rchol=np.eye(3)))


def test_set_get_parameter_values_and_changed_flags():
    m = pd.newPsydiff(**synth.config_c1(N=3, T=4))
    pt = m["pars"]["parameterTable"]
    pt.changed[:] = False
    pd.setParameterValues(pt, [0.25, -0.3], ["mm_Y1", "drift_eta1"])   # second value unchanged -> not flagged
    assert pd.getParameterValues(m)["mm_Y1"] == 0.25
    assert pt.changed[:, 4].all() and not pt.changed[:, 0].any()
    with pytest.warns(UserWarning, match="not found in the parameterTable"):
        pd.setParameterValues(pt, [1.0], ["nope"])
    with pytest.raises(PsydiffError, match="does not match"):
        pd.setParameterValues(pt, [1.0, 2.0], ["mm_Y1"])
    c = pd.clonePsydiffModel(m)
    pd.setParameterValues(c["pars"]["parameterTable"], {"mm_Y1": 9.0})
    assert pd.getParameterValues(m)["mm_Y1"] == 0.25


def test_codegen_and_cross_compile_for_sm100a():
    m = pd.newPsydiff(**synth.config_c1(N=2, T=3))
    src = m["cppCode"]
    assert "psy_filter.cuh" in src and "dx = DRIFT*x;" in src and "PSY_INSTANTIATE" in src
    assert "L_STRUCT = psy::L_DIAG" in src and "R_DIAG = true" in src
    cubin = pd.compileModelSource(m)
    assert os.path.getsize(cubin) > 10000
    log = open(cubin.replace(".cubin", ".log")).read() if os.path.exists(cubin.replace(".cubin", ".log")) else ""
    assert "sm_100a" in log or log == ""
    # element-wise syntax variants of the reference examples (R/prepareModel.R:172-179, 345-355)
    kw = synth.config_c1(N=2, T=3)
    kw["parameters"] = {"a": -0.3, "b": 0.2, "lam": np.array([["1"], ["l2 = 0.5"]], dtype=object)}
    kw["latentEquations"] = "dx(0) = a*x(0);\n dx[1] = b*x[0] - 0.5*x[1]*exp(-x(0)*x(0));"
    kw["manifestEquations"] = "y(0) = x(0);\n y(1) = lam(1,0)*x(1) + sin(t)*0.0;"
    assert os.path.exists(pd.compileModelSource(pd.newPsydiff(**kw)))


def test_unsupported_equation_is_reported():
    kw = synth.config_c1(N=2, T=3)
    kw["latentEquations"] = "arma::mat E = arma::expmat(DRIFT); dx = E*x;"
    with pytest.raises(PsydiffError, match="unsupported on device"):
        pd.compileModelSource(pd.newPsydiff(**kw))
    with pytest.raises(ValueError, match="lower triangular"):
        pd.newPsydiff(**dict(synth.config_c1(N=2, T=3), A0=np.array([[1.0, 0.3], [0.0, 1.0]])))


def _sweep(P, levels, direction=0):
    """levels: list of (eps, partial vector).  Returns grad, total."""
    L = _lib.lib()
    s = L.psy_sweep_create(P)
    need = np.zeros(2 * P + 1, dtype=np.uint8)
    done = C.c_int(0)
    for lv, (eps, part) in enumerate(levels):
        part = np.ascontiguousarray(part, dtype=np.float64)
        _lib.check(L.psy_sweep_resolve(s, _lib.dptr(part), eps, lv, len(levels), direction, _lib.u8ptr(need), C.byref(done)))
        if done.value:
            break
    g = np.zeros(P)
    tot = C.c_double()
    _lib.check(L.psy_sweep_finish(s, direction, _lib.dptr(g), C.byref(tot)))
    L.psy_sweep_destroy(s)
    return g, tot.value, need


def test_sweep_bookkeeping_eps_fallback():
    P = 3

    def part(total, nb, Dp, Dm, nfp, nfm, nbt):
        return np.concatenate([[total, nb], Dp, Dm, nfp, nfm, nbt])
    # level 0: parameter 1's left side fails (one non-finite person); level 1 repairs it with a larger eps
    l0 = part(100.0, 0, [1e-6, 2e-6, 3e-6], [-1e-6, np.nan, -3e-6], [0, 0, 0], [0, 1, 0], [0, 0, 0])
    l1 = part(100.0, 0, [0, 0, 0], [0, -2e-5, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0])
    g, tot, _ = _sweep(P, [(1e-8, l0), (1e-6, l1)])
    assert tot == 100.0
    assert g[0] == pytest.approx(2e-6 / 2e-8) and g[2] == pytest.approx(6e-6 / 2e-8)
    assert g[1] == pytest.approx((2e-6 + 2e-5) / (1e-8 + 1e-6))       # denominators: eps actually used on each side
    # exhausted eps[] -> NaN entry, others unaffected
    g, _, _ = _sweep(P, [(1e-8, l0)])
    assert np.isnan(g[1]) and np.isfinite(g[[0, 2]]).all()
    # a non-finite base person outside the touched set poisons every total -> all NaN
    lb = part(np.nan, 1, [1e-6] * 3, [-1e-6] * 3, [0] * 3, [0] * 3, [0] * 3)
    assert np.isnan(_sweep(P, [(1e-8, lb)])[0]).all()
    # one-sided directions
    gl, _, _ = _sweep(P, [(1e-6, l1)], direction=1)
    assert gl[1] == pytest.approx(2e-5 / 1e-6)
    with pytest.raises(PsydiffError, match="Unknown direction"):
        _sweep(P, [(1e-6, l1)], direction=7)


def test_exported_primitives_match_oracle_primitives():
    from oracle.oracle import _lib as orc
    O, L = orc(), _lib.lib()
    rng = np.random.default_rng(1)
    n, m = 3, 2
    s = 2 * n + 1
    wm, wc, wm2 = np.zeros(s), np.zeros(s), np.zeros(s)
    L.psy_getMeanWeights(n, .2, 2., 0., _lib.dptr(wm)); L.psy_getCovWeights(n, .2, 2., 0., _lib.dptr(wc))
    O.psy_oracle_getMeanWeights(C.c_int(n), C.c_double(.2), C.c_double(2.), C.c_double(0.), _lib.dptr(wm2))
    assert np.allclose(wm, wm2) and abs(wm.sum() - 1) < 1e-12
    Y = np.asfortranarray(rng.standard_normal((m, s)))
    mu = np.zeros(m)
    L.psy_predictMu(m, s, _lib.dptr(Y), _lib.dptr(wm), _lib.dptr(mu))
    assert np.allclose(mu, Y @ wm)
    Rc = np.asfortranarray(np.diag([.5, .7]))
    a, b = np.zeros((m, m), order="F"), np.zeros((m, m), order="F")
    L.psy_predictSquareRootOfS(m, s, _lib.dptr(Y), _lib.dptr(mu), _lib.dptr(Rc), wc[0], wc[1], _lib.dptr(a))
    O.psy_oracle_predictSquareRootOfS(C.c_int(m), C.c_int(s), _lib.dptr(Y), _lib.dptr(mu), _lib.dptr(Rc), C.c_double(wc[0]), C.c_double(wc[1]), _lib.dptr(b))
    assert np.allclose(a, b, atol=1e-12)
    # computeIndividualM2LL (R/RcppExports.R:214) against the closed form
    S = a @ a.T
    raw = rng.standard_normal(m)
    out = C.c_double()
    L.psy_computeIndividualM2LL(m, _lib.dptr(raw), _lib.dptr(mu), _lib.dptr(np.asfortranarray(S)), C.byref(out))
    r = raw - mu
    assert out.value == pytest.approx(m * np.log(2 * np.pi) + np.linalg.slogdet(S)[1] + r @ np.linalg.solve(S, r))
    out2 = C.c_double()
    L.psy_computeIndividualM2LLChol(m, _lib.dptr(raw), _lib.dptr(mu), _lib.dptr(a), C.byref(out2))
    assert out2.value == pytest.approx(out.value)
    assert L.psy_cholupdate(m, _lib.dptr(a), _lib.dptr(raw), 1.0, b"sideways") == 1     # Rcpp::stop on unknown direction
    X = np.asfortranarray(rng.standard_normal((7, 3)))
    R = np.zeros((3, 3), order="F")
    L.psy_qr(7, 3, _lib.dptr(X), _lib.dptr(R))
    assert np.allclose(R.T @ R, X.T @ X)


def test_drift_jacobian_sparsity_is_conservative():
    """codegen.drift_dependency: exact for plain element statements, all-ones for anything it cannot read."""
    full2 = [0b11, 0b11]
    m = pd.newPsydiff(**synth.config_c4(N=2, T=3))
    assert codegen.drift_dependency(m["_spec"]) == [0b10, 0b10111, 0b1000, 0b11101, 0b100000, 0b110101]
    assert "dep(int i)" in m["cppCode"]
    base = synth.config_c1(N=2, T=3)

    def dep(latent, **kw):
        return codegen.drift_dependency(pd.newPsydiff(**dict(base, latentEquations=latent, **kw))["_spec"])
    assert dep("dx = DRIFT*x;") == full2                                   # every DRIFT entry is a free parameter
    fixed = {"DRIFT": np.array([["a = -0.3", "0"], ["b = 0.2", "c = -0.5"]], dtype=object), "LAMBDA": base["parameters"]["LAMBDA"],
             "MANIFESTMEANS": base["parameters"]["MANIFESTMEANS"]}
    assert dep("dx = DRIFT*x;", parameters=fixed) == [0b01, 0b11]           # structural zero in DRIFT[0,1]
    sc = dict(base["parameters"], a=-0.3, b=0.2)
    assert dep("dx(0) = a*x(0);\n dx[1] = b*x[0] - 0.5*x[1]*exp(-x(0)*x(0)) + 2.5e-1*t;", parameters=sc) == [0b01, 0b11]
    assert dep("dx(1) = a*x(1);", parameters=sc) == [0, 0b10]               # dx(0) never assigned: stays zero
    # not understood -> everything depends on everything
    assert dep("double tmp = x(0)*x(1); dx(0) = tmp; dx(1) = a*x(1);", parameters=sc) == full2
    assert dep("dx(0) = a*x(0); dx(0) = dx(0) + x(1); dx(1) = b;", parameters=sc) == full2
    assert dep("dx = DRIFT*x; dx(0) = dx(0) + a;", parameters=sc) == full2
    assert dep("dx(0) = arma::accu(x); dx(1) = a*x(1);", parameters=sc) == full2
    assert dep("dx(0) = DRIFT(0,0)*x(0); dx(1) = a*x(1);", parameters=sc) == full2   # matrix element access: not analysed


def test_r_shim_only_uses_declared_abi():
    """r/psydiff_b200_shim.c (compiled by an R maintainer, not here): every psy_* it calls is declared in the header."""
    hdr = open(os.path.join(ROOT, "include", "psydiff_b200.h")).read()
    declared = set(re.findall(r"\b(psy_[A-Za-z0-9_]+)\s*\(", hdr)) | {"psy_model", "psy_container"}
    shim = open(os.path.join(ROOT, "r", "psydiff_b200_shim.c")).read()
    used = set(re.findall(r"\b(psy_[A-Za-z0-9_]+)\b", shim))
    assert used - declared == set(), used - declared
    assert {"psy_model_compile", "psy_model_create", "psy_upload", "psy_set_slots", "psy_fit", "psy_gradients", "psy_gradient"} <= used


def _literal_parameter_table(slot_labels, persons, grouping, groupingvariables):
    """R/prepareModel.R:471-476 + makeGroups :704-734, row by row, the slow way: replicate the slot rows per person, relabel
    grouped parameters, then unique() in order of appearance (quirk Q6: first-appearance order)."""
    rows = [[lab, p] for p in persons for lab in slot_labels]
    if grouping:
        for grp in [g for g in grouping.replace("\n", "").replace(" ", "").split(";") if g]:
            par, var = grp.split("|")
            for r in rows:
                if r[0] == par:
                    if var in ("person", "persons", "individual", "individuals"):
                        r[0] = f"{par}_p{int(r[1])}"
                    else:
                        r[0] = f"{par}_p{int(groupingvariables[var][int(r[1]) - 1])}"
    labels = []
    for lab, _ in rows:
        if lab not in labels:
            labels.append(lab)
    F = len(slot_labels)
    idx = [[labels.index(rows[i * F + k][0]) for k in range(F)] for i in range(len(persons))]
    return labels, idx


def test_slot_map_equals_literal_parameter_table_on_random_groupings():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 9), st.lists(st.sampled_from(["none", "person", "g1", "g2"]), min_size=11, max_size=11), st.integers(0, 10 ** 6))
    def check(N, kinds, seed):
        rng = np.random.default_rng(seed)
        kw = synth.config_c1(N=N, T=2)
        gv = {"g1": rng.integers(1, 4, size=N).astype(float), "g2": rng.integers(1, 3, size=N).astype(float)}
        slot_labels = ["drift_eta1", "drift_eta1_eta2", "drift_eta2_eta1", "drift_eta2", "mm_Y1", "mm_Y2", "lnT0var_eta1",
                       "T0var_eta2_eta1", "lnT0var_eta2", "lneta1_eta1", "lneta2_eta2"]
        grouping = "".join(f"{lab} | {k};\n" for lab, k in zip(slot_labels, kinds) if k != "none") or None
        m = pd.newPsydiff(**dict(kw, grouping=grouping, groupingvariables=gv if grouping else None))
        pt = m["pars"]["parameterTable"]
        labels, idx = _literal_parameter_table(slot_labels, list(range(1, N + 1)), grouping, gv)
        assert pt.theta_labels == labels
        assert pt.theta_index.tolist() == idx
    check()


def test_nvrtc_compile_path_matches_nvcc(monkeypatch):
    """PSY_COMPILER=nvrtc: in-process compile of the same TU (no nvcc subprocess); same kernel resources."""
    import subprocess
    m = pd.newPsydiff(**synth.config_c1(N=2, T=3))
    a = pd.compileModelSource(m)
    monkeypatch.setenv("PSY_COMPILER", "nvrtc")
    b = pd.compileModelSource(m)
    assert a != b and os.path.getsize(b) > 10000

    def res(path):
        out = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        return re.search(r"REG:(\d+) STACK:(\d+)", out).groups()
    assert res(a)[1] == res(b)[1] and abs(int(res(a)[0]) - int(res(b)[0])) <= 8
    kw = synth.config_c1(N=2, T=3)
    kw["latentEquations"] = "arma::mat E = arma::expmat(DRIFT); dx = E*x;"
    with pytest.raises(PsydiffError, match="unsupported on device"):
        pd.compileModelSource(pd.newPsydiff(**kw))


def test_all_exported_primitives_against_numpy():
    """The 21 host primitives of the C ABI (src/exposeInlineFunctions.cpp names) against their numpy definitions."""
    from oracle import numpy_ref as nr
    L = _lib.lib()
    rng = np.random.default_rng(11)
    n, m = 4, 3
    s = 2 * n + 1
    F = np.asfortranarray
    d = _lib.dptr
    wm, wc = np.zeros(s), np.zeros(s)
    L.psy_getMeanWeights(n, .3, 2., 1., d(wm)); L.psy_getCovWeights(n, .3, 2., 1., d(wc))
    assert np.allclose(wm, nr.getMeanWeights(n, .3, 2., 1.)) and np.allclose(wc, nr.getCovWeights(n, .3, 2., 1.))
    W = F(np.zeros((s, s)))
    L.psy_getWMatrix(s, d(wm), d(wc), d(W))
    assert np.allclose(W, nr.getWMatrix(wm, wc))
    c = .09 * (n + 1)
    A = F(np.tril(rng.standard_normal((n, n))) + 2 * np.eye(n))
    mvec = rng.standard_normal(n)
    X = F(np.zeros((n, s)))
    assert L.psy_getSigmaPoints(n, d(mvec), d(A), 1, c, d(X)) == 0
    assert np.allclose(X, nr.getSigmaPoints(mvec, A, c))
    P = F(A @ A.T)
    X2 = F(np.zeros((n, s)))
    assert L.psy_getSigmaPoints(n, d(mvec), d(P), 0, c, d(X2)) == 0 and np.allclose(X2, X)      # chol(P) inside
    assert L.psy_getSigmaPoints(n, d(mvec), d(F(-np.eye(n))), 0, c, d(X2)) != 0                   # not positive definite
    mean = np.zeros(n)
    L.psy_computeMeanFromSigmaPoints(n, s, d(X), d(wm), d(mean))
    assert np.allclose(mean, X @ wm)
    A2, P2 = F(np.zeros((n, n))), F(np.zeros((n, n)))
    L.psy_getAFromSigmaPoints(n, d(X), d(wm), c, d(A2)); L.psy_computePFromSigmaPoints(n, d(X), d(wm), c, d(P2))
    assert np.allclose(A2, A) and np.allclose(P2, A @ A.T)
    Mx, Phi = F(rng.standard_normal((n, n))), F(np.zeros((n, n)))
    L.psy_getPhi(n, d(Mx), d(Phi))
    assert np.allclose(Phi, nr.getPhi(Mx))
    Y = F(rng.standard_normal((m, s)))
    R = F(np.diag([.3, .4, .5]))
    mu, S, Cm = np.zeros(m), F(np.zeros((m, m))), F(np.zeros((n, m)))
    L.psy_predictMu(m, s, d(Y), d(wm), d(mu)); L.psy_predictS(m, s, d(Y), d(W), d(R), d(S)); L.psy_predictC(n, m, s, d(X), d(Y), d(W), d(Cm))
    assert np.allclose(mu, Y @ wm) and np.allclose(S, Y @ W @ Y.T + R) and np.allclose(Cm, X @ W @ Y.T)
    Spd = F(S @ S.T + np.eye(m))
    K = F(np.zeros((n, m)))
    assert L.psy_computeK(n, m, d(Cm), d(Spd), d(K)) == 0 and np.allclose(K, Cm @ np.linalg.inv(Spd))
    res = rng.standard_normal(m)
    mu2, Pu = np.zeros(n), F(np.zeros((n, n)))
    L.psy_updateM(n, m, d(mvec), d(K), d(res), d(mu2)); L.psy_updateP(n, m, d(P), d(K), d(Spd), d(Pu))
    assert np.allclose(mu2, mvec + K @ res) and np.allclose(Pu, P - K @ Spd @ K.T)
    srS = F(np.linalg.cholesky(Spd))
    Ksr = F(np.zeros((n, m)))
    assert L.psy_computeK_SR(n, m, d(Cm), d(srS), d(Ksr)) == 0 and np.allclose(Ksr, K)
    out = C.c_double()
    L.psy_computeIndividualM2LLChol(m, d(res), d(mu), d(srS), C.byref(out))
    assert out.value == pytest.approx(nr.m2ll_chol(m, res, mu, srS))
    sr2, want = F(np.zeros((m, m))), nr.predictSquareRootOfS(Y, mu, np.asarray(R), wc[0], wc[1])
    L.psy_predictSquareRootOfS(m, s, d(Y), d(mu), d(R), wc[0], wc[1], d(sr2))
    assert np.allclose(sr2, want)                      # Givens here, LAPACK Householder there: signs normalised by cholupdate (T3)
    Lc = F(np.linalg.cholesky(Spd)); x = rng.standard_normal(m) * .2
    want = nr.cholupdate(np.array(Lc), x, 2.0, "downdate")
    assert L.psy_cholupdate(m, d(Lc), d(x), 2.0, b"downdate") == 0 and np.allclose(np.tril(Lc), np.tril(want))
    lc, ch = F(rng.standard_normal((n, n))), F(np.zeros((n, n)))
    L.psy_logChol2Chol(n, d(lc), d(ch))
    exp = np.array(lc); exp[np.diag_indices(n)] = np.exp(np.diag(lc))
    assert np.allclose(ch, exp)


def test_python_mirror_exports_all_22_reference_functions():
    """R/RcppExports.R:12-257 lists 22 exported functions; the mirror offers every one under the same name and arguments."""
    import inspect
    ref = {"setParameterValues": ["parameterTable", "parameterValues", "parameterLabels"], "getParameterValues": ["psydiffModel"],
           "setParameterList": ["parameterTable", "parameterList", "person"], "getSigmaPoints": ["m", "A", "covarianceIsRoot", "c"],
           "getMeanWeights": ["n", "alpha", "beta", "kappa"], "getCovWeights": ["n", "alpha", "beta", "kappa"],
           "getWMatrix": ["meanWeights", "covWeights"], "computeMeanFromSigmaPoints": ["sigmaPoints", "meanWeights"],
           "getAFromSigmaPoints": ["sigmaPoints", "meanWeights", "c"], "computePFromSigmaPoints": ["sigmaPoints", "meanWeights", "c"],
           "getPhi": ["squareMatrix"], "predictMu": ["Y_", "meanWeights"], "predictS": ["Y_", "W", "R"], "predictC": ["sigmaPoints", "Y_", "W"],
           "computeK": ["C_", "S"], "updateM": ["m_", "K", "residual"], "updateP": ["P_", "K", "S"],
           "computeIndividualM2LL": ["nObservedVariables", "rawData", "expectedMeans", "expectedCovariance"],
           "cholupdate": ["L", "x", "v", "direction"], "qr_": ["X"], "logChol2Chol": ["logChol"], "clonePsydiffModel": ["model"]}
    assert len(ref) == 22
    for name, args in ref.items():
        fn = getattr(pd, name)
        assert list(inspect.signature(fn).parameters)[:len(args)] == args, name
    # a few calls through the mirror
    A = np.array([[2.0, 0.0], [0.5, 1.5]])
    X = pd.getSigmaPoints([1.0, 2.0], A, True, 0.08)
    wm = pd.getMeanWeights(2, .2, 2., 0.)
    assert np.allclose(pd.getAFromSigmaPoints(X, wm, 0.08), A) and np.allclose(pd.computePFromSigmaPoints(X, wm, 0.08), A @ A.T)
    assert np.allclose(pd.logChol2Chol(np.array([[np.log(2.0), 0.0], [0.5, np.log(1.5)]])), A)
    L2 = pd.cholupdate(A, [0.3, 0.1], 1.0, "update")
    assert np.allclose(np.tril(L2) @ np.tril(L2).T, A @ A.T + np.outer([0.3, 0.1], [0.3, 0.1]))
    with pytest.raises(PsydiffError, match="Unknown direction"):
        pd.cholupdate(A, [0.3, 0.1], 1.0, "sideways")
    S = A @ A.T
    assert pd.computeIndividualM2LL(2, [0.1, -0.2], [0.0, 0.0], S) == pytest.approx(
        2 * np.log(2 * np.pi) + np.linalg.slogdet(S)[1] + np.array([0.1, -0.2]) @ np.linalg.solve(S, [0.1, -0.2]))
    m = pd.newPsydiff(**synth.config_c1(N=3, T=4))
    pt, plist = m["pars"]["parameterTable"], m["pars"]["parameterList"]
    pd.setParameterValues(pt, {"drift_eta1": -0.9, "mm_Y2": 0.7})
    pd.setParameterList(pt, plist, 2)
    assert plist["DRIFT"][0, 0] == -0.9 and plist["MANIFESTMEANS"][1, 0] == 0.7
    with pytest.raises(PsydiffError, match="not found in data set"):
        pd.setParameterList(pt, plist, 99)


def test_model_template_and_helpers():
    t = pd.modelTemplate(True)
    assert "LATENTEQUATIONPLACEHOLDER" in t and "MEASUREMENTEQUATIONPLACEHOLDER" in t and "SR = true" in t
    assert "SR = false" in pd.modelTemplate(False)
    m = pd.newPsydiff(**synth.config_c1(N=2, T=3))
    pre = pd.prepareEquations("dx = DRIFT*x;", m["_spec"])
    assert "psy::Mat<2,2> DRIFT;" in pre and pre.rstrip().endswith("dx = DRIFT*x;")
    assert "PLACEHOLDER" not in m["cppCode"]
    mx = {"values": np.array([[-0.3, 0.0], [0.2, -0.5]]), "labels": np.array([["a", None], ["b", "c"]], dtype=object)}
    out = pd.fromMxMatrix(mx)
    assert out[0, 0] == "a = -0.3" and out[0, 1] == "0.0" and out[1, 1] == "c = -0.5"
    s = pd.sepNameFromValue(out)
    assert s["labels"][1, 0] == "b" and s["values"][1, 0] == 0.2


# ---- the C++ code generator behind the C ABI (psy_emit_model_source, psy_codegen.cpp) -------------------------------------
def test_cpp_code_generator_equals_the_python_one_on_many_models():
    """newPsydiff's source assembly lives in the shared library (callable from R and Python alike); the Python generator kept
    in codegen.py is its cross-check: identical text for the BASELINE configs and 60 random model structures."""
    from test_kernel_fuzz import random_model
    kws = [synth.config_c1(N=3, T=4), synth.config_c1(N=3, T=4, srUpdate=False), synth.config_c3(N=3, T=4), synth.config_c4(N=3, T=4),
           synth.config_coupled(K=4, N=3, T=4), synth.config_coupled(K=5, N=3, T=4)]
    kws += [random_model(7000 + s, stepper=("rk4", "runge_kutta_dopri5")[s % 2], srUpdate=(s % 3 != 0))[0] for s in range(60)]
    for kw in kws:
        m = pd.newPsydiff(**kw)
        spec = m["_spec"]
        assert m["cppCode"] == codegen.emit_cuda(spec) == codegen.emit_cuda_py(spec)
        assert codegen.drift_dependency(spec) == codegen.drift_dependency_py(spec)
    assert codegen.modelTemplate(True).startswith("// generated by psydiff_b200.codegen")
    L = _lib.lib()
    need = C.c_size_t(0)
    for sr in (0, 1):
        assert L.psy_model_template(sr, None, 0, C.byref(need)) == 0
        buf = C.create_string_buffer(need.value)
        assert L.psy_model_template(sr, buf, need.value, None) == 0 and buf.value.decode() == codegen.modelTemplate(bool(sr))


def test_cpp_code_generator_literals_round_trip_like_repr():
    """Fixed values become C++ literals that read back bit for bit (shortest round-trip digits, Python-repr layout)."""
    import copy
    rng = np.random.default_rng(0)
    vals = [0.0, -0.0, 1.0, -1.0, 1e-4, 1e-5, 9.999e-5, 123456789012345678.0, 1e15, 1e16, 1e17, 0.1, 1 / 3, 2.5e-300,
            1.7976931348623157e308, 5e-324, 100.0, 1234.5, 1e22, 1e23, float("nan"), float("inf"), float("-inf")]
    vals += list(rng.standard_normal(300) * 10.0 ** rng.integers(-30, 30, 300))
    spec0 = pd.newPsydiff(**synth.config_c1(N=3, T=4))["_spec"]
    for v in vals:
        sp = copy.deepcopy(spec0)
        c = sp.container("m0")
        c.values[0, 0], c.slots[0, 0] = v, -1
        src = codegen.emit_cuda(sp)
        assert src == codegen.emit_cuda_py(sp)
        text = [l for l in src.splitlines() if l.strip().startswith("m0[0] =")][0].split("=")[1].strip(" ;")
        if np.isfinite(v):
            assert float(text) == v and np.signbit(float(text)) == np.signbit(v)


def test_cpp_code_generator_errors():
    L = _lib.lib()
    m = pd.newPsydiff(**synth.config_c1(N=3, T=4))
    spec = m["_spec"]
    arr, keep = codegen._c_containers(spec)
    args = lambda **o: (o.get("n", 2), 2, spec.n_slots, spec.latentEquations.encode(), spec.manifestEquations.encode(), arr,
                        o.get("nc", len(spec.containers)), 1, o.get("stepper", 0))
    need = C.c_size_t(0)
    assert L.psy_emit_model_source(*args(), None, 0, C.byref(need)) == 0 and need.value == len(m["cppCode"].encode()) + 1
    small = C.create_string_buffer(16)
    assert L.psy_emit_model_source(*args(), small, 16, None) == 1 and b"too small" in L.psy_last_error()
    assert L.psy_emit_model_source(*args(n=3), None, 0, C.byref(need)) == 1 and b"nlatent x nlatent" in L.psy_last_error()
    assert L.psy_emit_model_source(*args(stepper=2), None, 0, C.byref(need)) == 1 and b"stepper" in L.psy_last_error()
    assert L.psy_emit_model_source(*args(nc=1), None, 0, C.byref(need)) == 1 and b"must include A0LogChol" in L.psy_last_error()
    with pytest.raises(ValueError, match="lower triangular"):          # the Python front end turns the C error into ValueError
        pd.newPsydiff(**synth.config_c1(N=3, T=4, A0=np.array([[1.0, 0.3], [0.0, 1.0]])))


def test_product_code_never_touches_the_checkers():
    """The package (Python and native sources) must not import, include or load anything under oracle/ or tests/: the oracle,
    the reference build, the host-compiled kernel and the R stand-in are checkers, not fallbacks."""
    pkg = os.path.join(ROOT, "psydiff_b200")
    bad = []
    for base, _, files in os.walk(pkg):
        if "_kernel_cache" in base or "__pycache__" in base:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                continue
            text = open(os.path.join(base, f)).read()
            code = "\n".join(l.split("#")[0].split("//")[0] for l in text.splitlines())       # comments may name them
            for pat in (r"\bimport\s+oracle\b", r"\bfrom\s+oracle\b", r"\bhostsim\b", r"reference_build", r"#include\s*[\"<].*oracle",
                        r"#include\s*[\"<].*tests/", r"libpsy_oracle", r"r_stub"):
                if re.search(pat, code):
                    bad.append((f, pat))
    assert bad == [], bad
    # the one hook the kernel header has for the host compile is inert on the device
    hdr = open(os.path.join(pkg, "csrc", "psy_filter.cuh")).read()
    assert "#ifdef __CUDA_ARCH__" in hdr and hdr.count("psy_sim_rcp_seed") == 1


def test_r_wrapper_file_is_lexically_well_formed_and_binds_registered_routines():
    """r/psydiff_b200.R cannot be executed here (no R): at least its brackets and strings balance, and every routine it .Call()s
    is registered by the shim with the number of arguments the call passes."""
    text = open(os.path.join(ROOT, "r", "psydiff_b200.R")).read()
    stack, i, n = [], 0, len(text)
    pairs = {")": "(", "]": "[", "}": "{"}
    while i < n:
        ch = text[i]
        if ch == "#":
            while i < n and text[i] != "\n":
                i += 1
            continue
        if ch in "\"'":
            q = ch
            i += 1
            while i < n and text[i] != q:
                i += 2 if text[i] == "\\" else 1
            assert i < n, "unterminated string"
        elif ch in "([{":
            stack.append(ch)
        elif ch in ")]}":
            assert stack and stack.pop() == pairs[ch], f"unbalanced {ch} at offset {i}"
        i += 1
    assert not stack
    shim = open(os.path.join(ROOT, "r", "psydiff_b200_shim.c")).read()
    registered = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(\w+)", \(DL_FUNC\)&\w+, (\d+)\}', shim)}

    def count_args(s, start):                         # arguments of the call whose "(" is at s[start]
        depth, args, j, seen = 0, 0, start, False
        while j < len(s):
            c = s[j]
            if c in "\"'":
                q = c
                j += 1
                while s[j] != q:
                    j += 2 if s[j] == "\\" else 1
            elif c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    return args + (1 if seen else 0)
            elif c == "," and depth == 1:
                args += 1
            elif depth >= 1 and not c.isspace():
                seen = True
            j += 1
        raise AssertionError("unterminated call")

    code = "\n".join(l.split("#")[0] for l in text.splitlines())
    calls = [(m.group(1), count_args(code, m.start(0) + len(".Call"))) for m in re.finditer(r'\.Call\("(\w+)"', code) if "..." not in code[m.start(0):m.start(0) + 80]]
    assert {c for c, _ in calls} >= {"R_psy_emit_source", "R_psy_compile", "R_psy_is_current", "R_psy_upload", "R_psy_fit", "R_psy_gradients", "R_psy_gradient"}
    for name, nargs in calls:
        assert name in registered, name
        assert nargs - 1 == registered[name], (name, nargs - 1, registered[name])      # the first argument is the routine's name


def test_typed_temporaries_and_expm_compile_for_the_device():
    from test_kernel_hostsim import vignette_expm_model
    m = pd.newPsydiff(**vignette_expm_model())
    assert m["cppCode"] == codegen.emit_cuda_py(m["_spec"])                   # both generators rewrite the declarations alike
    assert codegen.drift_dependency(m["_spec"]) == [3, 3]                     # temporaries: the sparsity analysis gives up (all ones)
    cubin = pd.compileModelSource(m)
    assert os.path.getsize(cubin) > 1000
    got = codegen.device_statements("arma :: colvec v=x; arma::mat T = A*B; if (a == b) arma::vec w = v;")
    assert got == "auto v =x; auto T = A*B; if (a == b) auto w = v;"
    bad = dict(vignette_expm_model())
    bad["latentEquations"] = "arma::mat T(2, 2);\n" + bad["latentEquations"]   # a size-only declaration has no device counterpart
    with pytest.raises(PsydiffError, match="unsupported on device"):
        pd.compileModelSource(pd.newPsydiff(**bad))


def test_named_vector_reads_like_a_dict():
    """getGradients returns a NamedVector (R's named numeric): dict semantics without per-entry objects until asked for."""
    from psydiff_b200.model import NamedVector
    v = NamedVector(["a", "b", "c"], np.array([1.0, np.nan, 3.0]))
    assert v["a"] == 1.0 and np.isnan(v["b"]) and len(v) == 3 and "c" in v and "z" not in v
    assert list(v) == ["a", "b", "c"] and list(v.keys()) == ["a", "b", "c"]
    assert v.values()[2] == 3.0 and isinstance(v.values()[0], float)
    assert dict(v.items())["c"] == 3.0 and set(dict(v)) == {"a", "b", "c"}
    assert np.array_equal(np.asarray(v), [1.0, np.nan, 3.0], equal_nan=True)
    assert np.array_equal(np.array(list(v.values())), np.asarray(v), equal_nan=True)
    w = NamedVector(["x", "y"], [2.0, 4.0], {"x": 0, "y": 1})
    assert w == {"x": 2.0, "y": 4.0} and {"x": 2.0, "y": 4.0} == w and w.get("q", 7.0) == 7.0
    with pytest.raises(KeyError):
        w["q"]
    with pytest.raises(TypeError):
        w["x"] = 1.0                                   # read-only, like the named vector R hands back


def test_synthetic_data_can_be_drawn_with_torch():
    """simulate_on = a torch device: the same simulation scheme with torch ops (examples/c5_gist_fit.py and bench.py --simulate-on-gpu
    use it on the GPU, where numpy needs minutes); another random stream, the same distribution."""
    for cfg in (synth.config_c4, synth.config_c5):
        a = cfg(N=400, T=12, seed=3)["dataset"]["observations"]
        b = cfg(N=400, T=12, seed=3, simulate_on="cpu")["dataset"]["observations"]
        assert a.shape == b.shape and np.isfinite(b).all()
        assert np.allclose(a.mean(0), b.mean(0), atol=0.15) and np.allclose(a.std(0), b.std(0), rtol=0.15)
