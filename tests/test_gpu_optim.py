"""The callers of the path running on the CUDA kernel: BFGS, GIST, standard errors,
failure conventions, simulateData.  GPU only."""
import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import optim, synth

pytestmark = pytest.mark.gpu


def _model(N=60, T=30, **kw):
    m = pd.newPsydiff(**synth.config_c1(N=N, T=T, seed=21, **kw))
    pd.compileModel(m)
    return m


def test_fit_and_gradient_closures_and_failure_conventions():
    m = _model(N=10, T=10)
    pars = pd.getParameterValues(m)
    names = list(pars)
    f = optim.psydiffFitInternal(list(pars.values()), names, m)
    assert f == pytest.approx(float(pd.fitModel(m)["m2LL"].sum()))
    g = optim.psydiffGradientsInternal(list(pars.values()), names, m)
    assert np.allclose(g, list(pd.getGradients(m).values()))
    bad = dict(pars, drift_eta1=500.0)                                  # explosive drift -> NaN -2LL
    assert optim.psydiffFitInternal(bad, None, m) == .5 * np.finfo(float).max       # R/optimWrapper.R:153-155
    gb = optim.psydiffGradientsInternal(bad, None, m)
    assert np.all(gb[np.isnan(list(pd.getGradients(m).values()))] == 1.0)          # NA -> 1, R/optimWrapper.R:185-188


def test_bfgs_recovers_generating_parameters():
    m = _model(N=200, T=40)
    start = pd.getParameterValues(m)
    pd.setParameterValues(m["pars"]["parameterTable"], {"drift_eta1": -0.6, "drift_eta2": -0.2, "mm_Y1": 0.4})
    f0 = float(pd.fitModel(m)["m2LL"].sum())
    out = optim.psydiffOptimBFGS(m, control={"maxit": 60})
    assert out["value"] < f0 - 10
    p = out["par"]
    assert abs(p["drift_eta1"] + 0.3) < 0.1 and abs(p["drift_eta2"] + 0.5) < 0.12 and abs(p["mm_Y1"]) < 0.2
    g = np.array(list(pd.getGradients(m).values()))
    assert np.max(np.abs(g)) < 1e-2 * abs(out["value"])
    se = optim.getStandardErrors(m)
    assert all(np.isfinite(v) and 0 < v < 1 for v in se.values())


def test_gist_lasso_zeroes_the_absent_cross_drift():
    m = _model(N=200, T=40)
    start = pd.getParameterValues(m)
    start["drift_eta1_eta2"] = 0.15            # truly 0 in the generating model (R/prepareModel.R:38-39)
    w = {k: 1.0 for k in start}
    res = optim.GIST(m, start, lambda_=40.0, adaptiveLassoWeights=w, regularizedParameters=["drift_eta1_eta2"],
                     maxIter_out=30, sig=.4, silent=True)
    est = pd.getParameterValues(res["model"])
    assert est["drift_eta1_eta2"] == 0.0
    r = res["regM2LL"][np.isfinite(res["regM2LL"])]
    assert len(r) >= 2 and np.all(np.diff(r) <= 1e-9) and res["convergence"]      # monotone line search


def test_simulate_data_and_best_of_n():
    m = _model(N=6, T=12)
    sim = optim.simulateData(m, rng=np.random.default_rng(1))
    assert sim["simulatedObservation"].shape == m["data"]["observations"].shape
    # without updates the latent scores are the pure model-implied trajectory, identical for all persons
    ls = sim["latentScores"].reshape(6, 12, 2)
    assert np.allclose(ls, ls[0][None])
    resid = sim["simulatedObservation"] - sim["predictedManifest"]
    assert 0.3 < resid.std() < 1.2                       # Rchol = diag(sqrt(.5))
    best = optim.bestOfN(m, nstart=5, sd=.05, rng=np.random.default_rng(2))
    assert best is not None and set(best) == set(pd.getParameterValues(m))


def test_fit_totals_many_sets_one_launch_and_batched_gist():
    m = _model(N=40, T=20)
    base = pd.getParameterValues(m)
    sets = [dict(base), dict(base, drift_eta1=-0.6, mm_Y1=0.3), dict(base, drift_eta1=500.0)]       # the last one explodes
    tot, bad = pd.fitTotals(m, sets)
    assert pd.getParameterValues(m) == base                               # the model is not modified
    for k in range(2):
        pd.setParameterValues(m["pars"]["parameterTable"], sets[k])
        assert tot[k] == pytest.approx(float(pd.fitModel(m)["m2LL"].sum()), rel=1e-13) and bad[k] == 0
    assert bad[2] == 40
    pd.setParameterValues(m["pars"]["parameterTable"], base)
    # GIST with the inner line search evaluated 4 candidates per launch == the sequential loop
    start = dict(base, drift_eta1_eta2=0.15)
    w = {k: 1.0 for k in start}
    kw = dict(lambda_=8.0, adaptiveLassoWeights=w, regularizedParameters=["drift_eta1_eta2"], maxIter_out=8, sig=.4, silent=True)
    a = optim.GIST(pd.clonePsydiffModel(m), start, **kw)
    b = optim.GIST(pd.clonePsydiffModel(m), start, batch=4, **kw)
    ra, rb = a["regM2LL"][np.isfinite(a["regM2LL"])], b["regM2LL"][np.isfinite(b["regM2LL"])]
    assert len(ra) == len(rb) and np.allclose(ra, rb, rtol=1e-11)
    pa, pb = pd.getParameterValues(a["model"]), pd.getParameterValues(b["model"])
    assert np.allclose(list(pa.values()), list(pb.values()), rtol=1e-9, atol=1e-12)
