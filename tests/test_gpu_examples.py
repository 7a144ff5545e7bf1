"""The model structures of the reference's own worked examples (R/prepareModel.R:159-395) through the CUDA path vs the
oracle: n != m, x(i) and x[i] syntax, scalar assignment to a 1-vector, alpha = .81 (positive wc0 -> cholupdate "update"
branch), zero diffusion hitting the sanity check, nonlinear measurement (quirk Q3 term active).  GPU only."""
import warnings

import numpy as np
import pytest

import psydiff_b200 as pd

pytestmark = pytest.mark.gpu


def _data(N, T, m, seed, dt=1.0, scale=1.0, offset=0.0):
    rng = np.random.default_rng(seed)
    obs = offset + scale * rng.standard_normal((N * T, m))
    obs[rng.random(obs.shape) < 0.1] = np.nan
    return {"person": np.repeat(np.arange(1, N + 1), T), "observations": obs,
            "dt": np.tile(np.concatenate([[0.0], np.full(T - 1, dt)]), N)}


def _compare(model, grad=True, extended=False):
    """-2LL / predictions against the double oracle (1e-9); the FD gradient against the double oracle, or — where the
    double restatement's own rounding noise exceeds the gate (tiny diffusion -> large A^-1) — against its 80-bit build."""
    from oracle.oracle import Oracle
    pd.compileModel(model)
    fit = pd.fitModel(model)
    ref = Oracle(model).fit()
    orc = Oracle(model, extended=extended)
    assert np.array_equal(np.isfinite(fit["m2LL"]), np.isfinite(ref["m2LL"])) and np.isfinite(ref["m2LL"]).all()
    assert np.max(np.abs(fit["m2LL"] - ref["m2LL"]) / np.abs(ref["m2LL"])) < 1e-9
    assert np.nanmax(np.abs(fit["predictedManifest"] - ref["predictedManifest"])) < 1e-8 * max(1.0, np.nanmax(np.abs(ref["predictedManifest"])))
    if grad:
        g = np.array(list(pd.getGradients(model).values()))
        gr = orc.gradients()["grad_clean"]
        ftot = abs(float(ref["m2LL"].sum()))
        atol = 64 * np.finfo(float).eps * ftot / (2 * float(model["eps"][0]))
        assert np.all(np.abs(g - gr) <= 1e-6 * np.abs(gr) + atol), list(zip(model["pars"]["parameterTable"].theta_labels, g, gr))
    return fit


def test_ctexample3_structure_n1_m3_alpha081():
    # R/prepareModel.R:172-205: one latent, three manifests, alpha = .81, labelled Rchol diagonal and L
    kw = dict(
        dataset=_data(8, 25, 3, 1, dt=0.5, scale=2.0, offset=8.0),
        latentEquations="dx(0) = cint + driftEta1*x(0);",
        manifestEquations="y(0) = x(0);\ny(1) = manifestmean2 + lambda2*x(0);\ny(2) = manifestmean3 + lambda3*x(0);",
        parameters={"driftEta1": -1.150227824, "cint": 11.214338733, "lambda2": 0.480098989, "lambda3": 0.959200513,
                    "manifestmean2": 2.824753235, "manifestmean3": 5.606085485},
        m0=np.array([["9.5"]], dtype=object), A0=np.array([[1.2]]),
        Rchol=np.array([["mvar1 = 0.9", "0", "0"], ["0", "mvar2 = 0.7", "0"], ["0", "0", "mvar3 = 1.1"]], dtype=object),
        L=np.array([["lvar = 1.3"]], dtype=object), alpha=.81, beta=2, kappa=0, integrateFunction="rk4", eps=np.array([1e-6]))
    m = pd.newPsydiff(**kw)
    assert m["pars"]["parameterTable"].theta_labels[-4:] == ["lnlvar", "lnmvar1", "lnmvar2", "lnmvar3"]
    _compare(m)


def test_second_order_structure_scalar_assignment_and_sanity_check():
    # R/prepareModel.R:253-279: y = x(0) (scalar -> 1-vector), numeric A0/L, character m0 and Rchol, tiny diffusion -> .01
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = pd.newPsydiff(dataset=_data(5, 40, 1, 2, scale=3.0), latentEquations="dx(0) = x(1);\ndx(1) = cint + a*x(0) +b*x(1);",
                          manifestEquations="y = x(0);", L=np.diag([0.001, 0.001]), Rchol=np.array([["1"]], dtype=object),
                          A0=np.eye(2), m0=np.array([["0"], ["1"]], dtype=object), parameters={"cint": 1.0, "a": -.5, "b": -.5},
                          integrateFunction="rk4", eps=np.array([1e-6]))
    assert any("replaced by .01" in str(x.message) for x in w)
    assert m["timeStep"][0] == pytest.approx(1 / 30)
    _compare(m)


def test_predator_prey_structure_bracket_indexing_n2_m4():
    # R/prepareModel.R:345-375: x[0]-style indexing, four manifests from two latents, L = 0 (sanity check -> .01)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = pd.newPsydiff(
            dataset=_data(6, 30, 4, 3, dt=0.18, scale=0.5, offset=1.0),
            latentEquations="dx[0] = x[0]*(alpha - beta*x[1]);\ndx[1] = -x[1]*(gamma - delta*x[0]);",
            manifestEquations="y[0] = 1*x[0];\ny[1] = a1*x[0];\ny[2] = 1*x[1];\ny[3] = a2*x[1];",
            parameters={"alpha": 0.1, "beta": .1, "gamma": .1, "delta": .1, "a1": 1.0, "a2": 1.0},
            m0=np.array([["m01 = 1"], ["m02 = 1"]], dtype=object),
            A0=np.array([["a01 = 1", "0"], ["a012 = .2", "a02 = 1"]], dtype=object),
            Rchol=np.array([["r1 = 0.1", "0", "0", "0"], ["0", "r2 = 0.1", "0", "0"], ["0", "0", "r3 = 0.1", "0"],
                            ["0", "0", "0", "r4 = 0.1"]], dtype=object),
            L=np.array([["0", "0"], ["0", "0"]], dtype=object), integrateFunction="rk4", eps=np.array([1e-6]))
    assert m["timeStep"][0] == pytest.approx(0.18 / 30)      # dt < 1 is integrated (quirk Q1: fabs semantics)
    _compare(m, extended=True)


def test_nonlinear_measurement_reproduces_quirk_q3():
    # with a nonlinear h the column-0 deviation Y0 - mu is non-zero, so the v^(1/4)-weighted up/down-date matters
    m = pd.newPsydiff(dataset=_data(6, 20, 2, 4, scale=1.0, offset=1.0),
                      latentEquations="dx(0) = -0.4*x(0) + c1*x(1);\ndx(1) = -0.3*x(1);",
                      manifestEquations="y(0) = x(0) + q*x(0)*x(0);\ny(1) = exp(0.3*x(1));",
                      parameters={"c1": 0.2, "q": 0.15}, L=np.diag([0.5, 0.4]), Rchol=np.diag([0.6, 0.5]), A0=np.diag([0.8, 0.7]),
                      m0=np.array([[1.0], [0.5]]), integrateFunction="rk4", eps=np.array([1e-6]))
    fit = _compare(m)
    # and the covariance variant (textbook weights) differs from the SR variant here, as in the reference
    m2 = pd.newPsydiff(**{**dict(dataset=m["data"], latentEquations="dx(0) = -0.4*x(0) + c1*x(1);\ndx(1) = -0.3*x(1);",
                                 manifestEquations="y(0) = x(0) + q*x(0)*x(0);\ny(1) = exp(0.3*x(1));", parameters={"c1": 0.2, "q": 0.15},
                                 L=np.diag([0.5, 0.4]), Rchol=np.diag([0.6, 0.5]), A0=np.diag([0.8, 0.7]), m0=np.array([[1.0], [0.5]]),
                                 integrateFunction="rk4", eps=np.array([1e-6]), srUpdate=False)})
    fit2 = _compare(m2)
    assert np.max(np.abs(fit["m2LL"] - fit2["m2LL"]) / np.abs(fit["m2LL"])) > 1e-6


def test_time_conventions_and_operator_surface():
    """Drift sees the interval-LOCAL time (integration restarts at 0 for every interval, R/modelTemplate.R:317-318),
    the measurement equation the ABSOLUTE time (timeSum, :353-355); also exercises matrix containers, whole-vector
    statements mixed with element statements, % and unary minus."""
    kw = dict(
        dataset=_data(6, 18, 2, 5, dt=0.7, scale=1.0, offset=0.5),
        latentEquations="dx = DRIFT*x + CINT;\ndx(1) = dx(1) - 0.1*x(1)*x(1)*x(1) + 0.3*t;\ndx = dx - 0.01*(x % x);",
        manifestEquations="y = LAMBDA*x;\ny = y + MANIFESTMEANS;\ny(0) = y(0) + 0.2*sin(0.5*t);",
        parameters={"DRIFT": np.array([["d11 = -0.5", "0.1"], ["d21 = 0.2", "d22 = -0.6"]], dtype=object),
                    "CINT": np.array([["c1 = 0.3"], ["0.1"]], dtype=object),
                    "LAMBDA": np.array([["1", "0"], ["l21 = 0.4", "1"]], dtype=object),
                    "MANIFESTMEANS": np.array([["0.2"], ["mm2 = -0.1"]], dtype=object)},
        L=np.array([["l1 = 0.5", "0"], ["l21x = 0.1", "l2 = 0.4"]], dtype=object),       # lower-triangular (not diagonal) diffusion
        Rchol=np.array([["r1 = 0.6", "0"], ["r21 = 0.1", "r2 = 0.5"]], dtype=object),     # non-diagonal measurement root
        A0=np.array([["0.8", "0"], ["a21 = 0.1", "0.7"]], dtype=object), m0=np.array([[1.0], [0.5]]),
        integrateFunction="rk4", eps=np.array([1e-6]))
    m = pd.newPsydiff(**kw)
    assert "L_STRUCT = psy::L_FULL" in m["cppCode"] and "R_DIAG = false" in m["cppCode"]
    _compare(m, extended=True)          # nonlinear drift: gradient against the 80-bit oracle
    # the same model under dopri5 and in the covariance variant
    m["integrateFunction"] = "runge_kutta_dopri5"
    _compare(m, grad=False)
    m3 = pd.newPsydiff(**dict(kw, srUpdate=False))
    _compare(m3, extended=True)
