"""Host loops of R/optimWrapper.R / R/GIST.R without a GPU: the BFGS minimiser and the lasso helpers."""
import numpy as np
import pytest

from psydiff_b200 import optim


def test_vmmin_rosenbrock_and_quadratic():
    f = lambda x: (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    g = lambda x: np.array([-2 * (1 - x[0]) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
    out = optim.vmmin([-1.2, 1.0], f, g, maxit=200)
    assert out["convergence"] == 0 and np.allclose(out["par"], [1, 1], atol=1e-3)   # optim(c(-1.2,1), fr, grr, method="BFGS") reaches ~9.6e-18
    A = np.array([[3.0, 1.0], [1.0, 2.0]])
    out = optim.vmmin([5.0, -3.0], lambda x: 0.5 * x @ A @ x, lambda x: A @ x)
    assert np.allclose(out["par"], 0, atol=1e-5)
    with pytest.raises(Exception, match="not finite"):
        optim.vmmin([0.0], lambda x: np.nan, lambda x: np.zeros(1))


def test_reg_value_and_subgradients():
    theta = {"a": 0.5, "b": -2.0, "c": 0.0}
    w = {"a": 1.0, "b": 2.0, "c": 1.0}
    assert optim.getRegValue(10, theta, ["b", "c"], w) == pytest.approx(40.0)
    assert optim.getRegValue(10, theta, ["a"]) == pytest.approx(5.0)
    sub = optim.getSubgradients(theta, {"a": 1.0, "b": 3.0, "c": 4.0}, ["b", "c"], 10, w)
    assert sub["a"] == 1.0 and sub["b"] == pytest.approx(3.0 - 20.0) and sub["c"] == 0.0      # |4| < 10 -> inside the interval
    sub = optim.getSubgradients(theta, {"a": 1.0, "b": 3.0, "c": 14.0}, ["c"], 10, w)
    assert sub["c"] == pytest.approx(4.0)


def test_data_template():
    d = optim.getDataTemplate(3, 2, 4, 0.5)
    assert d["observations"].shape == (12, 2) and list(d["dt"][:5]) == [0, .5, .5, .5, 0] and list(d["person"][:5]) == [1, 1, 1, 1, 2]
    with pytest.raises(Exception, match="length of dts"):
        optim.getDataTemplate(3, 2, 4, [1, 2])


# ---- the optimiser loops end to end on the CPU: evaluations through the host-compiled kernel source (tests/hostsim) -----------
def _hostsim_closures(N=40, T=25):
    """fit / gradient closures with the interfaces optim.GIST(fitTotal=, gradients=, fitTotalsMany=) and
    optim.psydiffOptimBFGS(fn=, gr=) accept — the hooks the sharded multi-GPU path uses — backed by the kernel source compiled
    for the host.  Test double only: the product closures launch the CUDA kernel."""
    import psydiff_b200 as pd
    from psydiff_b200 import synth
    from hostsim.hostsim import HostSim
    model = pd.newPsydiff(**synth.config_c1(N=N, T=T, seed=21))
    hs = HostSim(model)
    pt = model["pars"]["parameterTable"]

    def total(m):
        f = hs.fit(theta=m["pars"]["parameterTable"].theta_values, states=False)["m2LL"]
        return float("nan") if np.isnan(f).any() else float(f.sum())

    def grads(m):
        g = hs.gradients(theta=m["pars"]["parameterTable"].theta_values, eps=1e-6)["grad"]
        return dict(zip(pt.theta_labels, g.tolist()))

    def totals_many(m, sets):
        th = np.array([[s[l] for l in pt.theta_labels] for s in sets])
        f = hs.fit_many(th)
        return np.array([np.nan if np.isnan(row).any() else float(row.sum()) for row in f])   # same summation order as total()

    return pd, model, total, grads, totals_many


def test_gist_lasso_on_the_cpu_through_the_kernel_source():
    pd, model, total, grads, totals_many = _hostsim_closures()
    start = pd.getParameterValues(model)
    start["drift_eta1_eta2"] = 0.15                      # truly 0 in the generating model (R/prepareModel.R:38-39)
    w = {k: 1.0 for k in start}
    kw = dict(lambda_=25.0, adaptiveLassoWeights=w, regularizedParameters=["drift_eta1_eta2"], maxIter_out=30, sig=.4, silent=True,
              fitTotal=total, gradients=grads)
    a = optim.GIST(pd.clonePsydiffModel(model), start, **kw)
    est = pd.getParameterValues(a["model"])
    assert est["drift_eta1_eta2"] == 0.0                 # the lasso zeroes it exactly (proximal step)
    ra = a["regM2LL"][np.isfinite(a["regM2LL"])]
    assert len(ra) >= 2 and np.all(np.diff(ra) <= 1e-9)  # monotone line search (R/GIST.R:121-190)
    # the batched line search (k candidates per launch) walks the same path as the sequential one
    b = optim.GIST(pd.clonePsydiffModel(model), start, batch=4, fitTotalsMany=totals_many, **kw)
    rb = b["regM2LL"][np.isfinite(b["regM2LL"])]
    assert len(ra) == len(rb) and np.allclose(ra, rb, rtol=1e-11)
    assert np.allclose(list(est.values()), list(pd.getParameterValues(b["model"]).values()), rtol=1e-9, atol=1e-12)
    with pytest.raises(Exception, match="Infeasible starting values"):
        optim.GIST(pd.clonePsydiffModel(model), dict(start, drift_eta1=500.0), **kw)


def test_bfgs_on_the_cpu_through_the_kernel_source():
    pd, model, total, grads, _ = _hostsim_closures(N=60, T=30)
    pt = model["pars"]["parameterTable"]

    def fn(pars, parsnames=None, model=None):            # psydiffFitInternal's contract (R/optimWrapper.R:146-157)
        names, vals = optim._named(pars, parsnames)
        pd.setParameterValues(model["pars"]["parameterTable"], vals, names)
        f = total(model)
        return .5 * np.finfo(float).max if np.isnan(f) else f

    def gr(pars, parsnames=None, model=None):            # psydiffGradientsInternal's contract (R/optimWrapper.R:167-190)
        names, vals = optim._named(pars, parsnames)
        pd.setParameterValues(model["pars"]["parameterTable"], vals, names)
        g = grads(model)
        out = np.array([g[n] for n in names])
        out[np.isnan(out)] = 1.0
        return out

    pd.setParameterValues(pt, {"drift_eta1": -0.6, "drift_eta2": -0.2, "mm_Y1": 0.4})
    f0 = total(model)
    out = optim.psydiffOptimBFGS(model, fn=fn, gr=gr, control={"maxit": 40})
    assert out["value"] < f0 - 5
    p = out["par"]
    assert abs(p["drift_eta1"] + 0.3) < 0.2 and abs(p["drift_eta2"] + 0.5) < 0.25 and abs(p["mm_Y1"]) < 0.35
    assert np.max(np.abs(gr(p, None, model))) < 1e-2 * abs(out["value"])
    assert fn(dict(p, drift_eta1=500.0), None, model) == .5 * np.finfo(float).max      # failure convention
