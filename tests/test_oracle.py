"""The CPU oracle against its anchors (no GPU).  The reference holds no golden vectors; what pins the oracle is the reference's
OWN C++ compiled and run here (tests/test_reference_build.py where /root/reference exists, and its committed outputs
tests/golden/*_reference.json everywhere).  This file holds the independent anchors: (i) the exact discrete-time Kalman filter
for linear models, (ii) SR == covariance variant for linear measurement equations, (iii) primitive identities, (iv) its own
80-bit build, (v) a numpy/LAPACK restatement — and the comparison with the reference-run fixtures."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import synth
from oracle import exact_kf
from oracle.oracle import Oracle, _lib as orc_lib

HERE = os.path.dirname(os.path.abspath(__file__))


def _exact_c1(kw):
    A = np.array([[-.3, 0], [.2, -.5]])
    G = np.diag([np.sqrt(.5)] * 2)
    R = np.diag([np.sqrt(.5)] * 2)
    persons = kw["dataset"]["person"]
    out = []
    for p in np.unique(persons):
        sel = persons == p
        out.append(exact_kf.kalman_m2ll(A, np.zeros(2), G, np.eye(2), np.zeros(2), R, np.array([1., 2.]), np.eye(2),
                                        kw["dataset"]["observations"][sel], kw["dataset"]["dt"][sel]))
    return np.array(out)


def test_oracle_vs_exact_kalman_filter_c1():
    kw = synth.config_c1()
    ex = _exact_c1(kw)
    r30 = Oracle(pd.newPsydiff(**kw)).fit()["m2LL"]
    r100 = Oracle(pd.newPsydiff(**synth.config_c1(timeStep=np.array([1 / 100])))).fit()["m2LL"]
    e30, e100 = np.max(np.abs(r30 - ex) / np.abs(ex)), np.max(np.abs(r100 - ex) / np.abs(ex))
    assert e30 < 1e-9 and e100 < 1e-11     # RK4: O(h^4) convergence to the exact filter
    assert e100 < e30 / 50


def test_oracle_missing_data_vs_exact():
    kw = synth.config_c1(N=10, T=30, missing=0.3, seed=3)
    kw["dataset"]["observations"][7] = np.nan
    ex = _exact_c1(kw)
    r = Oracle(pd.newPsydiff(**kw)).fit()["m2LL"]
    assert np.max(np.abs(r - ex) / np.abs(ex)) < 1e-9


def test_sr_equals_covariance_variant_for_linear_measurement():
    kw = synth.config_c1(N=10, T=30, missing=0.2, seed=4)
    a = Oracle(pd.newPsydiff(**kw)).fit()
    b = Oracle(pd.newPsydiff(**dict(kw, srUpdate=False))).fit()
    assert np.max(np.abs(a["m2LL"] - b["m2LL"]) / np.abs(a["m2LL"])) < 1e-12
    assert np.max(np.abs(a["latentScores"] - b["latentScores"])) < 1e-11


def test_dopri5_vs_exact_kalman_filter():
    kw = synth.config_c1(N=6, T=20, integrateFunction="runge_kutta_dopri5")
    ex = _exact_c1(kw)
    r = Oracle(pd.newPsydiff(**kw)).fit()["m2LL"]
    assert np.max(np.abs(r - ex) / np.abs(ex)) < 1e-5      # controller tolerance eps_abs = eps_rel = 1e-6


def test_integrate_const_step_rule_T1():
    L = orc_lib()
    assert L.psy_oracle_rk4_steps(0.0, 1.0, 1.0 / 30) == 30
    assert L.psy_oracle_rk4_steps(0.0, 0.77, 1.0 / 30) == 23     # remainder dropped
    assert L.psy_oracle_rk4_steps(0.0, 0.01, 1.0 / 30) == 0      # dt < h: no step at all
    assert L.psy_oracle_rk4_steps(0.0, 0.5, 0.1) == 5
    from psydiff_b200 import flops
    for t1, h in [(1.0, 1 / 30), (0.77, 1 / 30), (0.01, 1 / 30), (0.5, 0.1), (1.5, 0.0166), (0.3, 0.1)]:
        assert flops.rk4_step_count(t1, h) == L.psy_oracle_rk4_steps(0.0, t1, h)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_primitive_identities():
    L = orc_lib()
    rng = np.random.default_rng(0)
    n = 4
    wm, wc = np.zeros(2 * n + 1), np.zeros(2 * n + 1)
    L.psy_oracle_getMeanWeights(C.c_int(n), C.c_double(.2), C.c_double(2.), C.c_double(0.), _dp(wm))
    L.psy_oracle_getCovWeights(C.c_int(n), C.c_double(.2), C.c_double(2.), C.c_double(0.), _dp(wc))
    assert abs(wm.sum() - 1) < 1e-12
    W = np.zeros((2 * n + 1, 2 * n + 1), order="F")
    L.psy_oracle_getWMatrix(C.c_int(2 * n + 1), _dp(wm), _dp(wc), _dp(W))
    D = np.eye(2 * n + 1) - np.outer(wm, np.ones(2 * n + 1))
    assert np.allclose(W, D @ np.diag(wc) @ D.T, atol=1e-12)
    # cholupdate(L, x, 1, "update"): L'L'^T = L L^T + x x^T
    Lm = np.asfortranarray(np.linalg.cholesky(np.cov(rng.standard_normal((n, 50)))))
    x = rng.standard_normal(n) * 0.3
    L2 = Lm.copy(order="F")
    L.psy_oracle_cholupdate(C.c_int(n), _dp(L2), _dp(x), C.c_double(1.0), C.c_int(1))
    L2 = np.tril(L2)
    assert np.allclose(L2 @ L2.T, Lm @ Lm.T + np.outer(x, x), atol=1e-12)
    L3 = L2.copy(order="F")
    L.psy_oracle_cholupdate(C.c_int(n), _dp(L3), _dp(x), C.c_double(1.0), C.c_int(-1))
    assert np.allclose(np.tril(L3), Lm, atol=1e-12)
    # quirk Q3: the weight enters as v^(1/4)
    L4 = Lm.copy(order="F")
    L.psy_oracle_cholupdate(C.c_int(n), _dp(L4), _dp(x), C.c_double(16.0), C.c_int(1))
    L4 = np.tril(L4)
    assert np.allclose(L4 @ L4.T, Lm @ Lm.T + 4.0 * np.outer(x, x), atol=1e-12)
    # qr_: R^T R = X^T X
    X = np.asfortranarray(rng.standard_normal((9, n)))
    R = np.zeros((n, n), order="F")
    L.psy_oracle_qr(C.c_int(9), C.c_int(n), _dp(X), _dp(R))
    assert np.allclose(R.T @ R, X.T @ X, atol=1e-12) and np.allclose(R, np.triu(R))
    # sigma points <-> A
    m, A = rng.standard_normal(n), np.asfortranarray(np.tril(rng.standard_normal((n, n))) + 2 * np.eye(n))
    Xs = np.zeros((n, 2 * n + 1), order="F")
    c = .04 * n
    L.psy_oracle_getSigmaPoints(C.c_int(n), _dp(m), _dp(A), C.c_double(c), _dp(Xs))
    A2 = np.zeros((n, n), order="F")
    L.psy_oracle_getAFromSigmaPoints(C.c_int(n), _dp(Xs), _dp(wm), C.c_double(c), _dp(A2))
    assert np.allclose(A2, A, atol=1e-12)


def test_gradient_definitions_agree():
    o = Oracle(pd.newPsydiff(**synth.config_c1(N=6, T=12)))
    g = o.gradients()
    assert np.allclose(g["grad_ref"], g["grad_clean"], rtol=1e-5, atol=1e-5)
    gl, gr = o.gradients(direction="left")["grad_clean"], o.gradients(direction="right")["grad_clean"]
    assert np.allclose(0.5 * (gl + gr), g["grad_clean"], rtol=1e-6, atol=1e-6)


def test_double_oracle_fd_noise_on_nonlinear_model():
    """Documents WHY gradient parity on C4 is checked against the 80-bit build: in double the reference formulation's
    FD gradient at eps = 1e-6 carries rounding noise far above the 1e-6 gate (weighted mean with wm0 = -24)."""
    m = pd.newPsydiff(**synth.config_c4(N=4, T=10))
    d, e = Oracle(m), Oracle(m, extended=True)
    fd, fe = d.fit()["m2LL"], e.fit()["m2LL"]
    assert np.max(np.abs(fd - fe) / np.abs(fe)) < 1e-9          # the likelihood itself is fine
    gd, ge = d.gradients(nthreads=4)["grad_clean"], e.gradients(nthreads=4)["grad_clean"]
    noise = np.max(np.abs(gd - ge) / np.abs(ge))
    assert 1e-6 < noise < 1e-2, noise


def test_golden_vectors():
    """Regression pins generated by tests/golden/make_golden.py (oracle outputs; see that script's header)."""
    g = json.load(open(os.path.join(HERE, "golden", "c1_oracle.json")))
    o = Oracle(pd.newPsydiff(**synth.config_c1()))
    assert np.allclose(o.fit()["m2LL"], g["m2LL"], rtol=1e-12)
    assert np.allclose(o.gradients()["grad_clean"], g["grad_clean"], rtol=1e-7, atol=1e-6)
    assert np.allclose(g["exact_kf_m2LL"], g["m2LL"], rtol=1e-9)


def test_cpp_oracle_vs_numpy_lapack_restatement():
    """Third implementation: the numpy/LAPACK restatement (oracle/numpy_ref.py) and the C++ oracle agree on a nonlinear
    model with a nonlinear measurement equation, missing data and irregular dt — and LAPACK's QR sign convention does not
    leak."""
    from oracle import numpy_ref as nr
    rng = np.random.default_rng(3)
    N, T = 3, 9
    obs = 0.8 + rng.standard_normal((N * T, 2))
    obs[rng.random(obs.shape) < 0.2] = np.nan
    obs[4] = np.nan
    dt = rng.choice([0.3, 0.5, 0.77], size=N * T)
    dt[::T] = 0.0
    kw = dict(dataset={"person": np.repeat(np.arange(1, N + 1), T), "observations": obs, "dt": dt},
              latentEquations="dx(0) = -0.4*x(0) + c1*x(1) + 0.1*t;\ndx(1) = -0.3*x(1) - 0.2*x(0)*x(0)*x(1);",
              manifestEquations="y(0) = x(0) + q*x(0)*x(0);\ny(1) = exp(0.3*x(1)) + 0.1*sin(t);",
              parameters={"c1": 0.2, "q": 0.15}, L=np.array([[0.5, 0.0], [0.1, 0.4]]), Rchol=np.array([[0.6, 0.0], [0.1, 0.5]]),
              A0=np.array([[0.8, 0.0], [0.1, 0.7]]), m0=np.array([[1.0], [0.5]]), integrateFunction="rk4", timeStep=np.array([0.05]))
    model = pd.newPsydiff(**kw)
    ref = Oracle(model).fit()

    def drift(x, p, person, t):
        return np.array([-0.4 * x[0] + 0.2 * x[1] + 0.1 * t, -0.3 * x[1] - 0.2 * x[0] * x[0] * x[1]])

    def meas(x, p, person, t):
        return np.array([x[0] + 0.15 * x[0] * x[0], np.exp(0.3 * x[1]) + 0.1 * np.sin(t)])
    ref5 = Oracle(pd.newPsydiff(**dict(kw, integrateFunction="runge_kutta_dopri5"))).fit()      # rule T2: adaptive dopri5
    refc = Oracle(pd.newPsydiff(**dict(kw, srUpdate=False))).fit()                              # covariance variant (a24)
    assert np.max(np.abs(ref5["m2LL"] - ref["m2LL"]) / np.abs(ref["m2LL"])) > 1e-9               # a different integrator ...
    for p in range(N):
        sl = slice(p * T, (p + 1) * T)
        f, lat, pred = nr.fit_person(obs[sl], dt[sl], drift, meas, None, p + 1, np.array([1.0, 0.5]), np.array(kw["A0"]),
                                     np.array(kw["L"]), np.array(kw["Rchol"]), h=0.05)
        assert abs(f - ref["m2LL"][p]) / abs(f) < 1e-11
        assert np.max(np.abs(lat - ref["latentScores"][sl])) < 1e-10
        assert np.max(np.abs(pred - ref["predictedManifest"][sl])) < 1e-10
        f5, lat5, _ = nr.fit_person(obs[sl], dt[sl], drift, meas, None, p + 1, np.array([1.0, 0.5]), np.array(kw["A0"]),
                                    np.array(kw["L"]), np.array(kw["Rchol"]), h=0.05, stepper="dopri5")
        assert abs(f5 - ref5["m2LL"][p]) / abs(f5) < 1e-11                                       # ... restated twice, same numbers
        assert np.max(np.abs(lat5 - ref5["latentScores"][sl])) < 1e-10
        fc, latc = nr.fit_person_cov(obs[sl], dt[sl], drift, meas, None, p + 1, np.array([1.0, 0.5]), np.array(kw["A0"]),
                                     np.array(kw["L"]), np.array(kw["Rchol"]), h=0.05)
        assert abs(fc - refc["m2LL"][p]) / abs(fc) < 1e-11 and np.max(np.abs(latc - refc["latentScores"][sl])) < 1e-10


# ---- the oracle against the committed outputs of the reference's own C++ (oracle/_ref build, tests/golden/*_reference.json) ----
@pytest.mark.parametrize("name", ["c1", "c1_missing_cov", "c3", "c4", "c4_T100", "n8"])
def test_oracle_equals_reference_run_fixtures(name):
    import json
    import reference_fixtures as rf
    ref = rf.load(name)
    model = pd.newPsydiff(**rf.CASES[name]())
    assert model["pars"]["parameterTable"].theta_labels == ref["labels"]
    o = Oracle(model)
    rel = rf.check_fit(o.fit(), ref)
    if name == "c1":
        assert rel < 1e-12
    if "getGradients" in ref:
        g = o.gradients(nthreads=8)
        rf.check_gradient(g["grad_ref"], ref)


def test_dormand_prince_tableau_equals_scipys_in_all_three_implementations():
    """The Dormand-Prince 5(4) coefficients of the kernel (psy_filter.cuh), the oracle (psy_oracle.cpp) and the odeint stand-in
    of the reference build (oracle/refbuild) against an independent transcription: scipy.integrate's RK45 tableau."""
    import re
    from scipy.integrate._ivp.rk import RK45
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = {"a21": RK45.A[1][0], "a31": RK45.A[2][0], "a32": RK45.A[2][1], "a41": RK45.A[3][0], "a42": RK45.A[3][1], "a43": RK45.A[3][2],
            "a51": RK45.A[4][0], "a52": RK45.A[4][1], "a53": RK45.A[4][2], "a54": RK45.A[4][3],
            "a61": RK45.A[5][0], "a62": RK45.A[5][1], "a63": RK45.A[5][2], "a64": RK45.A[5][3], "a65": RK45.A[5][4],
            "c1": RK45.B[0], "c3": RK45.B[2], "c4": RK45.B[3], "c5": RK45.B[4], "c6": RK45.B[5]}
    err = {1: RK45.E[0], 3: RK45.E[2], 4: RK45.E[3], 5: RK45.E[4], 6: RK45.E[5], 7: RK45.E[6]}     # scipy: 5th minus 4th order weights, negated
    for path, eprefix in (("psydiff_b200/csrc/psy_filter.cuh", "e"), ("oracle/psy_oracle.cpp", "dc"),
                          ("oracle/refbuild/include/boost/numeric/odeint.hpp", "dc")):
        text = open(os.path.join(root, path)).read()
        env = {}
        for name, expr in re.findall(r"\b([ace]\d+|dc\d)\s*=\s*([-+0-9./ c()]+?)\s*[,;]", text):
            if re.fullmatch(r"[-+0-9./ c()]+", expr) and name not in env:
                env[name] = eval(expr, {}, env)            # plain rational arithmetic, later ones refer to c1..c6
        for k, v in want.items():
            assert env[k] == pytest.approx(v, rel=1e-15), (path, k)
        for i, v in err.items():
            assert abs(env[f"{eprefix}{i}"]) == pytest.approx(abs(v), rel=1e-12), (path, i)
        assert np.sign(env[f"{eprefix}1"]) * np.sign(env[f"{eprefix}7"]) == np.sign(err[1]) * np.sign(err[7])
