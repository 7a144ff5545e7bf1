"""Seeded random model structures through the host-compiled kernel source (tests/hostsim) and the oracle.

What varies: n and m (n != m included), Jacobian sparsity of the drift (exercises the code generator's dependency masks, whose
errors would be silently wrong numbers), literal / labelled / grouped parameters, `t` and `person` in the equations, nonlinear
measurement (quirk Q3 active), diagonal and full lower-triangular L and Rchol, alpha (negative and positive wc0), missing data,
ragged persons, irregular dt including dt < h, both steppers and both update variants.  Runs without a GPU.
"""
import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import codegen
from hostsim.hostsim import HostSim


def _lower(rng, k, diag_only, label, lo, hi, labelled):
    out = np.full((k, k), "0", dtype=object)
    for i in range(k):
        for j in range(i + 1):
            if i == j:
                v = rng.uniform(lo, hi)
            elif diag_only:
                continue
            else:
                v = rng.uniform(-0.15, 0.15)
            out[i, j] = f"{label}{i}{j} = {v:.6f}" if (labelled and rng.random() < 0.7) else f"{v:.6f}"
    return out


def random_model(seed, stepper="rk4", srUpdate=True):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 6)), int(rng.integers(1, 5))
    N, T = int(rng.integers(3, 7)), int(rng.integers(5, 12))
    params, lat, man = {}, [], []

    def coef(name, lo, hi):
        v = rng.uniform(lo, hi)
        if rng.random() < 0.5:
            params[name] = float(np.round(v, 6))
            return name
        return f"{v:.6f}"

    expected_dep = []
    for i in range(n):
        terms = [f"-{coef(f'd{i}', 0.2, 0.9)}*x({i})"]
        dep = 1 << i
        for k in range(n):
            if k != i and rng.random() < 0.35:
                terms.append(f"+ {coef(f'c{i}{k}', -0.3, 0.3)}*x[{k}]")
                dep |= 1 << k
        kind = rng.integers(0, 4)
        if kind == 0:
            k = int(rng.integers(0, n))
            terms.append(f"- 0.05*x({k})*x({k})*x({i})")
            dep |= 1 << k
        elif kind == 1:
            k = int(rng.integers(0, n))
            terms.append(f"+ 0.2*tanh(x({k}))")
            dep |= 1 << k
        elif kind == 2:
            terms.append("+ 0.1*sin(0.7*t)")
        if rng.random() < 0.2:
            terms.append("+ 0.01*person")
        lat.append(f"dx({i}) = " + " ".join(terms) + ";")
        expected_dep.append(dep)
    if n > 1 and rng.random() < 0.25:      # a statement the dependency analysis cannot read: must fall back to "all"
        lat.append("dx = dx - 0.01*(x % x);")
        expected_dep = [(1 << n) - 1] * n
    for a in range(m):
        k = int(rng.integers(0, n))
        terms = [f"{coef(f'lam{a}', 0.6, 1.2)}*x({k})"]
        if rng.random() < 0.4:
            k2 = int(rng.integers(0, n))
            terms.append(f"+ 0.1*x({k2})*x({k})")
        if rng.random() < 0.3:
            terms.append(f"+ {coef(f'mm{a}', -0.5, 0.5)}")
        if rng.random() < 0.2:
            terms.append("+ 0.05*cos(0.3*t)")
        man.append(f"y({a}) = " + " ".join(terms) + ";")
    dt = rng.choice([0.004, 0.25, 0.5, 0.8, 1.0, 1.3], size=(N, T)) + (rng.random((N, T)) < 0.3) * rng.random((N, T)) * 0.05
    dt[:, 0] = 0.0
    obs = 0.5 + rng.standard_normal((N * T, m))
    obs[rng.random(obs.shape) < 0.15] = np.nan
    person = np.repeat(np.arange(11, 11 + N), T)
    keep = np.ones(N * T, dtype=bool)
    keep[(N - 1) * T + T // 2:] = False                 # last person is shorter
    grouping, gv = None, None
    grp_names = [k for k in params if k.startswith("d")]
    if grp_names and rng.random() < 0.6:
        if rng.random() < 0.5:
            grouping = f"{grp_names[0]} | person;"
        else:
            grouping = f"{grp_names[0]} | g;"
            g = np.zeros(11 + N)
            g[11:] = (np.arange(N) % 2) + 1
            gv = {"g": g}
    kw = dict(dataset={"person": person[keep], "observations": obs[keep], "dt": dt.reshape(-1)[keep]},
              latentEquations="\n".join(lat), manifestEquations="\n".join(man), parameters=params,
              L=_lower(rng, n, rng.random() < 0.5, "l", 0.2, 0.6, True),
              Rchol=_lower(rng, m, rng.random() < 0.5, "r", 0.3, 0.7, True),
              A0=_lower(rng, n, rng.random() < 0.5, "a", 0.5, 1.0, rng.random() < 0.5),
              m0=np.array([[f"m0{i} = {rng.uniform(-1, 1):.6f}"] if rng.random() < 0.5 else [f"{rng.uniform(-1, 1):.6f}"] for i in range(n)], dtype=object),
              grouping=grouping, groupingvariables=gv, alpha=float(rng.choice([0.2, 0.81, 1.0])), beta=2.0, kappa=0.0,
              timeStep=np.array([1.0 / 40]), integrateFunction=stepper, eps=np.array([1e-6]), srUpdate=srUpdate)
    return kw, expected_dep


def _compare(model):
    from oracle.oracle import Oracle
    hs = HostSim(model)
    fit = hs.fit()
    ref = Oracle(model).fit()
    ext = Oracle(model, extended=True)
    assert np.isfinite(ref["m2LL"]).all() and np.isfinite(fit["m2LL"]).all()
    assert np.max(np.abs(fit["m2LL"] - ref["m2LL"]) / np.maximum(np.abs(ref["m2LL"]), 1e-300)) < 1e-9   # 0 vs 0 (nothing observed) -> 0
    scale = max(1.0, np.nanmax(np.abs(ref["latentScores"])))
    assert np.nanmax(np.abs(fit["latentScores"] - ref["latentScores"])) < 1e-9 * scale
    assert np.nanmax(np.abs(fit["predictedManifest"] - ref["predictedManifest"])) < 1e-9 * scale
    g = hs.gradients(eps=1e-6)["grad"]
    gr = ext.gradients(direction="central", nthreads=8)["grad_clean"]
    atol = 64 * np.finfo(float).eps * abs(float(ref["m2LL"].sum())) / 2e-6
    assert np.all(np.abs(g - gr) <= 1e-6 * np.abs(gr) + atol), list(zip(model["pars"]["parameterTable"].theta_labels, g, gr))


@pytest.mark.parametrize("seed", range(24))
def test_random_structure_rk4_sr(seed):
    kw, dep = random_model(1000 + seed)
    model = pd.newPsydiff(**kw)
    # the dependency masks the kernel is specialised with are exactly the states each statement names (or "all")
    assert codegen.drift_dependency(model["_spec"]) == dep
    _compare(model)


@pytest.mark.parametrize("seed", range(4))
def test_random_structure_dopri5(seed):
    kw, _ = random_model(2000 + seed, stepper="runge_kutta_dopri5")
    _compare(pd.newPsydiff(**kw))


@pytest.mark.parametrize("seed", range(4))
def test_random_structure_covariance_variant(seed):
    kw, _ = random_model(3000 + seed, srUpdate=False)
    _compare(pd.newPsydiff(**kw))
