"""The callers either side of the hot path: fit / gradient closures, BFGS, GIST, start-value search,
standard errors, data simulation.  Same names and conventions as R/optimWrapper.R, R/GIST.R, R/bestOfN.R,
R/simulateData.R.  These are O(P) host loops; every likelihood / gradient evaluation inside them is one batched launch
of the CUDA kernel through ``fitModel`` / ``getGradients``.
"""
from __future__ import annotations

import sys
import warnings
from typing import Dict, Optional, Sequence

import numpy as np

from ._lib import PsydiffError
from .model import (clonePsydiffModel, fitModel, fitTotals, getGradients, getGradientsParallel, getParameterValues,
                    setParameterValues)

DOUBLE_XMAX = sys.float_info.max


def _named(pars, parsnames):
    if isinstance(pars, dict):
        return list(pars.keys()), np.asarray(list(pars.values()), dtype=float)
    return list(parsnames), np.asarray(pars, dtype=float)


def psydiffFitInternal(pars, parsnames=None, model=None) -> float:
    """R/optimWrapper.R:146-157: theta -> Σ -2LL; a failed fit (error or any NA) maps to .5 * double.xmax."""
    names, vals = _named(pars, parsnames)
    setParameterValues(model["pars"]["parameterTable"], vals, names)
    try:
        fit = fitModel(model, states=False)   # the optimiser only sums -2LL: no rows x (n + m) state matrices per evaluation
    except PsydiffError:
        return .5 * DOUBLE_XMAX
    if np.isnan(fit["m2LL"]).any():
        return .5 * DOUBLE_XMAX
    return float(np.sum(fit["m2LL"]))


def psydiffGradientsInternal(pars, parsnames=None, model=None) -> np.ndarray:
    """R/optimWrapper.R:167-190: theta -> gradient in parsnames order; an error gives all ones, NA entries become 1."""
    names, vals = _named(pars, parsnames)
    setParameterValues(model["pars"]["parameterTable"], vals, names)
    try:
        g = getGradientsParallel(model) if model.get("cl") is not None else getGradients(model)
    except PsydiffError:
        return np.ones(len(names))
    out = np.array([g[n] for n in names], dtype=float)
    out[np.isnan(out)] = 1.0
    return out


def vmmin(b, fn, gr, maxit=100, abstol=-np.inf, reltol=1e-8, trace=0):
    """Variable-metric (BFGS) minimiser with the backtracking line search of R's optim(method = "BFGS") — the optimiser
    psydiffOptimBFGS passes through to (R/optimWrapper.R:23-31).  stepredn = 0.2, acctol = 1e-4, reltest = 10."""
    stepredn, acctol, reltest = 0.2, 1e-4, 10.0
    b = np.array(b, dtype=float)
    n = len(b)
    f = fn(b)
    if not np.isfinite(f):
        raise PsydiffError("initial value in 'vmmin' is not finite")
    Fmin = f
    funcount = gradcount = 1
    g = np.asarray(gr(b), dtype=float)
    it = 1
    ilast = gradcount
    B = np.eye(n)
    count = 0
    while True:
        if ilast == gradcount:
            B = np.eye(n)
        X, c = b.copy(), g.copy()
        t = -B @ g
        gradproj = float(t @ g)
        if gradproj < 0.0:
            steplength, accpoint = 1.0, False
            while True:
                b = X + steplength * t
                count = int(np.sum(reltest + X == reltest + b))
                if count < n:
                    f = fn(b)
                    funcount += 1
                    accpoint = np.isfinite(f) and (f <= Fmin + gradproj * steplength * acctol)
                    if not accpoint:
                        steplength *= stepredn
                if count == n or accpoint:
                    break
            enough = (f > abstol) and abs(f - Fmin) > reltol * (abs(Fmin) + reltol)
            if not enough:
                count = n
                Fmin = f
            if count < n:
                Fmin = f
                g = np.asarray(gr(b), dtype=float)
                gradcount += 1
                it += 1
                t = steplength * t
                c = g - c
                D1 = float(t @ c)
                if D1 > 0:
                    Xc = B @ c
                    D2 = 1.0 + float(Xc @ c) / D1
                    B = B + (D2 * np.outer(t, t) - np.outer(Xc, t) - np.outer(t, Xc)) / D1
                else:
                    ilast = gradcount
            else:
                if ilast < gradcount:
                    count = 0
                    ilast = gradcount
        else:
            count = 0
            if ilast == gradcount:
                count = n
            else:
                ilast = gradcount
        if trace:
            print(f"iter {it:4d} value {Fmin:.6f}")
        if it >= maxit:
            break
        if gradcount - ilast > 2 * n:
            ilast = gradcount
        if count == n and ilast == gradcount:
            break
    return {"par": b, "value": float(Fmin), "counts": {"function": funcount, "gradient": gradcount},
            "convergence": 1 if it >= maxit else 0}


def psydiffOptimBFGS(model, fn=psydiffFitInternal, gr=psydiffGradientsInternal, control: Optional[dict] = None):
    """psydiffOptimBFGS (R/optimWrapper.R:23-31).  Returns optim()'s list: par (named), value, counts, convergence."""
    control = control or {}
    start = getParameterValues(model)
    names = list(start.keys())
    out = vmmin(np.array(list(start.values())), lambda p: fn(p, names, model), lambda p: gr(p, names, model),
                maxit=int(control.get("maxit", 100)), abstol=control.get("abstol", -np.inf),
                reltol=control.get("reltol", 1e-8), trace=int(control.get("trace", 0)))
    out["par"] = dict(zip(names, out["par"]))
    return out


def psydiffOptimNelderMead(model, control: Optional[dict] = None):
    """psydiffOptimNelderMead (R/optimWrapper.R:7-14): pass-through to a library Nelder–Mead (scipy here, stats::optim there)."""
    from scipy.optimize import minimize
    control = control or {}
    start = getParameterValues(model)
    names = list(start.keys())
    res = minimize(lambda p: psydiffFitInternal(p, names, model), np.array(list(start.values())), method="Nelder-Mead",
                   options={"maxiter": int(control.get("maxit", 500)), "xatol": 1e-8, "fatol": control.get("reltol", 1e-8)})
    return {"par": dict(zip(names, res.x)), "value": float(res.fun), "counts": {"function": res.nfev, "gradient": None},
            "convergence": 0 if res.success else 1}


def psydiffOptimx(model, method=("Nelder-Mead", "BFGS"), fn=psydiffFitInternal, gr=None, control: Optional[dict] = None):
    """psydiffOptimx (R/optimWrapper.R:41-48): run several optimisers from the same start; one result row per method."""
    rows = {}
    start = getParameterValues(model)
    for m in method:
        work = clonePsydiffModel(model)
        if m == "Nelder-Mead":
            rows[m] = psydiffOptimNelderMead(work, control)
        elif m == "BFGS":
            rows[m] = psydiffOptimBFGS(work, fn=fn, gr=gr or psydiffGradientsInternal, control=control)
        else:
            raise PsydiffError(f"optimiser {m} is not available in this build (no optimx/nlm/nlminb); use Nelder-Mead or BFGS")
    return {"class": "optimx", "start": start, "rows": rows}


def psydiffDEoptim(model, lower, upper, control: Optional[dict] = None):
    """psydiffDEoptim (R/optimWrapper.R:83-92): pass-through to a library differential-evolution optimiser (scipy here)."""
    from scipy.optimize import differential_evolution
    control = control or {}
    names = list(getParameterValues(model).keys())
    res = differential_evolution(lambda p: psydiffFitInternal(p, names, model), list(zip(lower, upper)),
                                 maxiter=int(control.get("itermax", 200)), seed=control.get("seed"), polish=False)
    return {"optim": {"bestmem": dict(zip(names, res.x)), "bestval": float(res.fun), "nfeval": res.nfev, "iter": res.nit}}


def psydiffSubplex(model, control: Optional[dict] = None):
    """psydiffSubplex (R/optimWrapper.R:102-111).  No subplex implementation ships in this image; the derivative-free
    simplex search it decomposes into (Nelder-Mead) is used on the full parameter vector instead."""
    return psydiffOptimNelderMead(model, control)


def psydiffNLopt(model, eval_f=psydiffFitInternal, control: Optional[dict] = None):
    """psydiffNLopt (R/optimWrapper.R:121-136): the reference asks nloptr for NLOPT_LD_LBFGS; here scipy's L-BFGS-B with
    the batched finite-difference gradient."""
    from scipy.optimize import minimize
    control = control or {}
    start = getParameterValues(model)
    names = list(start.keys())
    res = minimize(lambda p: eval_f(p, names, model), np.array(list(start.values())), jac=lambda p: psydiffGradientsInternal(p, names, model),
                   method="L-BFGS-B", options={"maxfun": int(control.get("maxeval", 500))})
    return {"solution": dict(zip(names, res.x)), "objective": float(res.fun), "iterations": res.nit, "status": int(res.status)}


def extractOptimx(model, opt):
    """extractOptimx (R/optimWrapper.R:57-72): write the best optimiser's parameters into the model (not refitted)."""
    if not (isinstance(opt, dict) and opt.get("class") == "optimx"):
        raise PsydiffError("opt has to be of class optimx")
    best = min(opt["rows"], key=lambda k: opt["rows"][k]["value"])
    setParameterValues(model["pars"]["parameterTable"], opt["rows"][best]["par"])
    return model


def getStandardErrors(model, ndeps: float = 1e-3) -> Dict[str, float]:
    """getStandardErrors (R/optimWrapper.R:199-211): sqrt(diag(2 * solve(H))), H = optimHess of Σ -2LL (central differences
    of central-difference gradients with step ndeps).  Each of the 2P gradient evaluations is one batched sweep."""
    pars = getParameterValues(model)
    names = list(pars.keys())
    x = np.array(list(pars.values()))
    work = clonePsydiffModel(model)
    work["eps"] = np.array([ndeps])
    work["direction"] = "central"
    P = len(x)
    H = np.zeros((P, P))
    for i in range(P):
        xp, xm = x.copy(), x.copy()
        xp[i] += ndeps
        xm[i] -= ndeps
        setParameterValues(work["pars"]["parameterTable"], xp, names)
        g1 = np.array(list(getGradients(work).values()))
        setParameterValues(work["pars"]["parameterTable"], xm, names)
        g2 = np.array(list(getGradients(work).values()))
        H[i] = (g1 - g2) / (2 * ndeps)
    H = 0.5 * (H + H.T)
    se = np.sqrt(np.diag(2 * np.linalg.inv(H)))
    return dict(zip(names, se))


def fitf(pars: Dict[str, float], model) -> float:
    """fitf (R/GIST.R:285-289): Σ -2LL at the named parameter vector."""
    setParameterValues(model["pars"]["parameterTable"], pars)
    return float(np.sum(fitModel(model, states=False)["m2LL"]))


# ---------------------------------------------------------------------------------------------------------- GIST
def getRegValue(lambda_, theta: Dict[str, float], regIndicators: Sequence[str], adaptiveLassoWeights=None) -> float:
    """getRegValue (R/GIST.R:337-363): Σ lambda * w_i * |theta_i| over the regularised parameters."""
    val = 0.0
    for th, v in theta.items():
        if th in regIndicators:
            val += lambda_ * abs(v) * (1.0 if adaptiveLassoWeights is None else adaptiveLassoWeights[th])
    return val


def getSubgradients(theta: Dict[str, float], jacobian: Dict[str, float], regIndicators, lambda_, adaptiveLassoWeights):
    """getSubgradients (R/GIST.R:302-327)."""
    sub = dict(jacobian)
    for lab in regIndicators:
        w = adaptiveLassoWeights[lab]
        if theta[lab] != 0:
            sub[lab] = sub[lab] + w * lambda_ * np.sign(theta[lab])
        else:
            lo, hi = -w * lambda_, w * lambda_
            sub[lab] = 0.0 if (lo < sub[lab] < hi) else np.sign(sub[lab]) * (abs(sub[lab]) - w * lambda_)
    return sub


def GIST(model, startingValues: Dict[str, float], lambda_, adaptiveLassoWeights: Dict[str, float], regularizedParameters,
         eta=1.5, sig=.2, initialStepsize=1.0, stepsizeMin=0.0, stepsizeMax=999999999.0,
         GISTLinesearchCriterion="monotone", GISTNonMonotoneNBack=5, maxIter_out=100, maxIter_in=100, break_outer=1e-8,
         verbose=0, silent=False, rng: Optional[np.random.Generator] = None, fitTotal=None, gradients=None,
         batch: int = 1, fitTotalsMany=None):
    """GIST (R/GIST.R:25-283; Gong et al. 2013): proximal-gradient lasso / adaptive lasso around fitModel + getGradients.
    ``rng`` drives the reference's random step-size reset (`runif(1) > .9`, R/GIST.R:88); ``rng=None`` disables it.
    ``fitTotal(model) -> Σ-2LL or NaN`` and ``gradients(model) -> dict`` replace fitModel / getGradients when persons are
    sharded over GPUs (psydiff_b200.dist.fitTotalSharded / getGradientsSharded); every rank then runs the same loop.
    ``batch`` > 1 evaluates that many consecutive candidates of the inner line search (step sizes s, eta s, eta² s, ...) in
    ONE launch (fitTotals, or ``fitTotalsMany(model, [dict, ...]) -> totals`` for shards) and then walks them in order, so
    the accepted step and every returned value are exactly those of the sequential loop."""
    getGradients_ = gradients or getGradients

    def fit_at(values: Dict[str, float]):
        setParameterValues(model["pars"]["parameterTable"], values)
        try:
            if fitTotal is not None:
                tot = fitTotal(model)
                return None if np.isnan(tot) else float(tot)
            out = fitModel(model, states=False)
        except PsydiffError:
            return None
        return None if np.isnan(out["m2LL"]).any() else float(np.sum(out["m2LL"]))

    k_out, convergence = 1, True
    m2LL_k = fit_at(startingValues)
    if m2LL_k is None:
        raise PsydiffError("Infeasible starting values in GIST.")
    par_km1 = None
    par_k = getParameterValues(model)
    names = list(par_k.keys())
    grad_km1 = None
    grad_k = getGradients_(model)
    reg_k = m2LL_k + getRegValue(lambda_, par_k, regularizedParameters, adaptiveLassoWeights)
    regM2LL = np.full(maxIter_out, np.nan)
    regM2LL[0] = reg_k
    vec = lambda d: np.array([d[n] for n in names], dtype=float)
    while k_out < maxIter_out:
        k = 0
        if par_km1 is None:
            stepsize = initialStepsize
        else:
            x_k, y_k = vec(par_k) - vec(par_km1), vec(grad_k) - vec(grad_km1)
            with np.errstate(divide="ignore", invalid="ignore"):
                stepsize = float(x_k @ y_k) / float(x_k @ x_k)      # Barzilai–Borwein (R/GIST.R:81-84)
            if rng is not None and rng.uniform() > .9:
                stepsize = initialStepsize
            if np.isnan(stepsize) or np.isinf(stepsize):
                if not silent:
                    warnings.warn(f"Outer iteration {k_out}: NA or infinite step size...")
                break
            if stepsize < stepsizeMin or stepsize > stepsizeMax:
                stepsize = initialStepsize
        par_kp1, m2LL_kp1, reg_kp1 = par_k, m2LL_k, reg_k

        def proximal(step):
            u = vec(par_k) - vec(grad_k) / step
            new = {}
            for i, nme in enumerate(names):
                if nme in regularizedParameters:
                    lam_i = lambda_ * adaptiveLassoWeights[nme]
                    new[nme] = float(np.sign(u[i]) * max(0.0, abs(u[i]) - lam_i / step))     # soft threshold
                else:
                    new[nme] = float(u[i])
            return new

        accepted = False
        while k < maxIter_in and not accepted:
            nb = max(1, min(int(batch), maxIter_in - k))
            steps = [stepsize]
            for _ in range(nb - 1):
                steps.append(steps[-1] * eta)       # repeated multiplication, bit for bit the sequential loop's s, s*eta, (s*eta)*eta, ...
            cands = [proximal(st) for st in steps]
            if nb == 1:
                fits = [fit_at(cands[0])]
            else:
                if fitTotalsMany is not None:
                    tots = np.asarray(fitTotalsMany(model, cands), dtype=float)
                else:
                    tot, bad = fitTotals(model, cands)
                    tots = np.where(bad > 0, np.nan, tot)
                fits = [None if np.isnan(v) else float(v) for v in tots]
            for st, new, f in zip(steps, cands, fits):
                stepsize = st
                if f is None:
                    stepsize = st * eta
                    k += 1
                    continue
                par_kp1, m2LL_kp1 = new, f
                reg_kp1 = f + getRegValue(lambda_, new, regularizedParameters, adaptiveLassoWeights)
                d2 = float(np.sum((vec(par_k) - vec(new)) ** 2))
                if GISTLinesearchCriterion == "monotone":
                    ok = reg_kp1 <= reg_k - (sig / 2) * stepsize * d2
                elif GISTLinesearchCriterion == "non-monotone":
                    nback = max(1, k_out - GISTNonMonotoneNBack)
                    ok = reg_kp1 <= np.nanmax(regM2LL[nback - 1:k_out]) - (sig / 2) * stepsize * d2
                else:
                    raise PsydiffError("Unknown GISTLinesearchCriterion. Possible are monotone and non-monotone.")
                if ok:
                    accepted = True
                    break
                stepsize = st * eta
                k += 1
        if k == maxIter_in and not silent:
            warnings.warn("Maximal number of inner iterations used by GIST. Consider increasing the number of inner iterations.")
        setParameterValues(model["pars"]["parameterTable"], par_kp1)
        g = getGradients_(model)
        if np.isnan(list(g.values())).any():
            convergence = False
            if not silent:
                raise PsydiffError("No gradients in GIST")
        par_km1, par_k, grad_km1, grad_k, m2LL_k, reg_k = par_k, par_kp1, grad_k, g, m2LL_kp1, reg_kp1
        k_out += 1
        regM2LL[k_out - 1] = reg_k
        if verbose:
            print(f"## [{k_out:3d}] m2LL: {m2LL_k:.3f} | regM2LL: {reg_k:.3f} | zeroed: "
                  f"{sum(par_k[r] == 0 for r in regularizedParameters):3d} ##")
        with np.errstate(divide="ignore", invalid="ignore"):
            if np.sqrt(np.sum((vec(par_k) - vec(par_km1)) ** 2)) / np.sqrt(np.sum(vec(par_km1) ** 2)) < break_outer:
                break
    if not np.isfinite(reg_k) or not np.isfinite(vec(par_k)).all() or not np.isfinite(vec(grad_k)).all():
        convergence = False
    if k_out == maxIter_out and not silent:
        warnings.warn("Maximal number of outer iterations used by GIST. Consider increasing the number of outer iterations.")
    sub = getSubgradients(par_k, grad_k, regularizedParameters, lambda_, adaptiveLassoWeights)
    return {"type": "psydiff", "model": model, "regM2LL": regM2LL, "convergence": convergence, "subgradients": sub}


# ------------------------------------------------------------------------------------- start values, simulation
def bestOfN(model, minVal=-5, maxVal=5, sd=.5, nstart=100, rng: Optional[np.random.Generator] = None):
    """bestOfN (R/bestOfN.R:1-42): random restarts around the current values; returns the best parameter vector found."""
    rng = rng or np.random.default_rng()
    pars = getParameterValues(model)
    names, x = list(pars.keys()), np.array(list(pars.values()))
    P = len(x)
    minVal, maxVal, sd = (np.full(P, v, dtype=float) if np.isscalar(v) else np.asarray(v, dtype=float) for v in (minVal, maxVal, sd))
    if len(minVal) != P:
        raise PsydiffError("Length of minVal does not correspond to the number of parameters in the model")
    if len(maxVal) != P:
        raise PsydiffError("Length of maxVal does not correspond to the number of parameters in the model")
    if len(sd) != P:
        raise PsydiffError("Length of sd does not correspond to the number of parameters in the model")
    # draw all nstart candidates first (same sampling scheme as the reference), then evaluate them 32 per launch
    cands = []
    for _ in range(nstart):
        cur = x.copy()
        for p in range(P):
            grid = np.linspace(minVal[p], maxVal[p], 1000)
            w = np.exp(-0.5 * ((grid - x[p]) / sd[p]) ** 2)
            cur[p] = rng.choice(grid, p=w / w.sum())
        cands.append(cur)
    best, bestfit = None, np.inf
    for c0 in range(0, nstart, 32):
        chunk = cands[c0:c0 + 32]
        tot, bad = fitTotals(model, np.asarray(chunk))
        for cur, f, b in zip(chunk, tot, bad):
            if b == 0 and np.isfinite(f) and f < bestfit:
                bestfit, best = float(f), dict(zip(names, cur))
    return best


def getDataTemplate(nSubjects, nVariables, nTimpoints, dts):
    """getDataTemplate (R/simulateData.R:9-23)."""
    dts = np.atleast_1d(np.asarray(dts, dtype=float))
    if len(dts) == 1:
        dts = np.tile(np.concatenate([[0.0], np.full(nTimpoints - 1, dts[0])]), nSubjects)
    elif len(dts) == nTimpoints:
        dts = np.tile(dts, nSubjects)
    elif len(dts) != nTimpoints * nSubjects:
        raise PsydiffError("The length of dts has to be 1, nTimepoints, or nTimpoints*nSubjects")
    return {"person": np.repeat(np.arange(1, nSubjects + 1), nTimpoints),
            "observations": np.full((nSubjects * nTimpoints, nVariables), np.nan), "dt": dts}


def simulateData(model, rng: Optional[np.random.Generator] = None):
    """simulateData (R/simulateData.R:70-91): prediction means from fitModel(skipUpdate = TRUE) plus measurement noise.
    The reference reads a non-existent parameterList$Rchol (:78); here R = Rchol Rchol' is rebuilt from RLogChol with the
    current parameter values of the FIRST person (the reference has a single shared parameter list as well)."""
    rng = rng or np.random.default_rng()
    f = fitModel(clonePsydiffModel(model), skipUpdate=True)
    spec, pt = model["_spec"], model["pars"]["parameterTable"]
    c = spec.container("RLogChol")
    Rlc = c.values.copy()
    for i in range(Rlc.shape[0]):
        for j in range(Rlc.shape[1]):
            if c.slots[i, j] >= 0:
                Rlc[i, j] = pt.theta_values[pt.theta_index[0, c.slots[i, j]]]
    Rc = Rlc.copy()
    Rc[np.diag_indices_from(Rc)] = np.exp(np.diag(Rlc))
    noise = rng.standard_normal(f["predictedManifest"].shape) @ Rc.T
    return {"latentScores": f["latentScores"], "predictedManifest": f["predictedManifest"],
            "simulatedObservation": f["predictedManifest"] + noise}
