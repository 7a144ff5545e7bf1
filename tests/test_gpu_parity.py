"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs.  GPU only.

Gates (BASELINE.json north_star): per-person -2LL within 1e-9 relative; finite-difference gradients within 1e-6
relative at eps = 1e-6 (entries with |g| < 1e-6 max|g| compared absolutely); identical non-finite pattern.
"""
import os

import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import synth

pytestmark = pytest.mark.gpu

M2LL_RTOL = 1e-9
GRAD_RTOL = 1e-6


def _oracle(model, extended=False):
    from oracle.oracle import Oracle
    return Oracle(model, extended=extended)


def _check_fit(model, states=True, skipUpdate=False):
    if model.get("_handle") is None:
        pd.compileModel(model)
    fit = pd.fitModel(model, skipUpdate=skipUpdate)
    ref = _oracle(model).fit(skip_update=skipUpdate)
    fin = np.isfinite(ref["m2LL"])
    assert np.array_equal(fin, np.isfinite(fit["m2LL"]))
    if skipUpdate:
        assert np.all(fit["m2LL"] == 0.0)
    else:
        rel = np.abs(fit["m2LL"][fin] - ref["m2LL"][fin]) / np.maximum(np.abs(ref["m2LL"][fin]), 1e-300)   # 0 vs 0 -> 0
        assert rel.max() < M2LL_RTOL, rel.max()
    if states:
        scale = max(1.0, np.nanmax(np.abs(ref["latentScores"])))
        assert np.nanmax(np.abs(fit["latentScores"] - ref["latentScores"])) < 1e-9 * scale
        assert np.nanmax(np.abs(fit["predictedManifest"] - ref["predictedManifest"])) < 1e-9 * scale
    # by-reference mutation of the model object (R/modelTemplate.R:412-414)
    assert model["m2LL"] is fit["m2LL"]
    return fit, ref


def _check_grad(model, direction="central", extended=False, nthreads=8):
    """extended=True compares with the 80-bit build of the oracle: the same reference algorithm with its own rounding
    noise removed.  Needed for nonlinear models, where the double build's FD noise exceeds the 1e-6 gate."""
    model["direction"] = direction
    if model.get("_handle") is None:
        pd.compileModel(model)
    g = pd.getGradients(model)
    ref = _oracle(model, extended).gradients(direction=direction, nthreads=nthreads)
    labels = model["pars"]["parameterTable"].theta_labels
    assert list(g.keys()) == labels
    gv, gr = np.array(list(g.values())), ref["grad_clean"]
    assert np.array_equal(np.isfinite(gv), np.isfinite(gr))
    # 1e-6 relative, plus the rounding-noise floor of a finite difference itself: both sides evaluate Σ-2LL (size |f|)
    # to a few ulp, and the difference is divided by the step actually used
    ftot = float(np.abs(np.nansum(ref["m2LL"])))
    step = float(np.atleast_1d(model["eps"])[0]) * (2.0 if direction == "central" else 1.0)
    atol = 64 * np.finfo(float).eps * ftot / step
    err = np.abs(gv - gr)
    bad = err > GRAD_RTOL * np.abs(gr) + atol
    assert not np.nanmax(bad), [(labels[i], gv[i], gr[i]) for i in np.flatnonzero(bad)]
    return gv, gr


def test_c1_fit_and_states():
    _check_fit(pd.newPsydiff(**synth.config_c1()))


def test_c1_gradient_central():
    _check_grad(pd.newPsydiff(**synth.config_c1()))


@pytest.mark.parametrize("direction", ["left", "right"])
def test_c1_gradient_one_sided(direction):
    _check_grad(pd.newPsydiff(**synth.config_c1(N=12, T=20)), direction)


def test_c1_skip_update():
    _check_fit(pd.newPsydiff(**synth.config_c1(N=8, T=12)), skipUpdate=True)


def test_missing_data_and_ragged_persons():
    kw = synth.config_c1(N=40, T=30, missing=0.25, seed=7)
    obs = kw["dataset"]["observations"]
    obs[5] = np.nan          # a fully missing time point
    obs[31:35] = np.nan
    # ragged: drop the last rows of some persons
    keep = np.ones(len(obs), dtype=bool)
    for p in (3, 17, 22):
        keep[p * 30 + 20:(p + 1) * 30] = False
    for k in ("person", "observations", "dt"):
        kw["dataset"][k] = kw["dataset"][k][keep]
    model = pd.newPsydiff(**kw)
    _check_fit(model)
    _check_grad(model)


def test_irregular_dt_remainder_dropped():
    # dt that are not multiples of h, dt < 1 and dt < h (zero RK4 steps): odeint's integrate_const drops the remainder
    kw = synth.config_c1(N=16, T=25, seed=11)
    rng = np.random.default_rng(5)
    dt = rng.choice([0.01, 0.2, 0.5, 0.77, 1.0, 1.31], size=16 * 25)
    dt[::25] = 0.0
    kw["dataset"]["dt"] = dt
    kw["timeStep"] = np.array([1.0 / 30])
    _check_fit(pd.newPsydiff(**kw))


def test_c4_small_fit_and_gradient():
    # Gradient parity is checked against the 80-bit build of the oracle: at eps = 1e-6 the DOUBLE restatement of the
    # reference formulation is itself only good to ~1e-4 relative on this model (its weighted sigma-point mean,
    # wm0 = -24, cancels ~50 ulp per RHS evaluation; tests/test_oracle.py::test_double_oracle_fd_noise and DESIGN.md).
    model = pd.newPsydiff(**synth.config_c4(N=32, T=16))
    fit, ref = _check_fit(model)
    assert np.isfinite(fit["m2LL"]).all()
    ext = _oracle(model, extended=True).fit(nthreads=8)
    rel_k = np.abs(fit["m2LL"] - ext["m2LL"]) / np.abs(ext["m2LL"])
    rel_o = np.abs(ref["m2LL"] - ext["m2LL"]) / np.abs(ext["m2LL"])
    print("C4 -2LL vs 80-bit oracle: kernel", rel_k.max(), " double oracle", rel_o.max())
    assert rel_k.max() < 1e-10
    gv, gr = _check_grad(model, extended=True)
    gd = _oracle(model).gradients(nthreads=8)["grad_clean"]
    print("C4 gradient vs 80-bit oracle: kernel", np.max(np.abs(gv - gr) / np.abs(gr)), " double oracle", np.max(np.abs(gd - gr) / np.abs(gr)))


def test_eps_fallback_and_nan_pattern():
    # a parameter value that makes one person's filter fail: the NaN pattern must match the oracle's
    kw = synth.config_c1(N=6, T=10)
    model = pd.newPsydiff(**kw)
    pd.compileModel(model)
    pt = model["pars"]["parameterTable"]
    pd.setParameterValues(pt, {"drift_eta1": 400.0})   # explosive drift -> overflow in the integrator
    fit = pd.fitModel(model)
    ref = _oracle(model).fit()
    assert np.array_equal(np.isfinite(fit["m2LL"]), np.isfinite(ref["m2LL"]))
    model["eps"] = np.array([1e-8, 1e-6])
    g = pd.getGradients(model)
    gref = _oracle(model).gradients()
    assert np.array_equal(np.isfinite(list(g.values())), np.isfinite(gref["grad_clean"]))


def test_covariance_variant_srUpdate_false():
    # a24: fitModel covariance variant (R/modelTemplate.R:583-847)
    model = pd.newPsydiff(**synth.config_c1(N=24, T=30, missing=0.2, seed=9, srUpdate=False))
    _check_fit(model)
    _check_grad(model)
    m4 = pd.newPsydiff(**synth.config_c4(N=8, T=10, srUpdate=False))
    fit, ref = _check_fit(m4)
    assert np.isfinite(fit["m2LL"]).all()


def test_dopri5_linear_and_c3():
    # a10: integrate() = controlled dopri5 (R/modelTemplate.R:331-333)
    model = pd.newPsydiff(**synth.config_c1(N=16, T=20, integrateFunction="runge_kutta_dopri5"))
    _check_fit(model)
    c3 = pd.newPsydiff(**synth.config_c3(N=64, T=24))
    pt = c3["pars"]["parameterTable"]
    assert len(pt.theta_labels) == 64 + 7 and "mu_p1" in pt.theta_labels and "mu_p64" in pt.theta_labels
    fit, ref = _check_fit(c3)
    assert np.isfinite(fit["m2LL"]).all()
    _check_grad(c3, extended=True)


def test_stepper_switch_after_compile_swaps_module():
    # the reference reads integrateFunction at every fitModel call (R/modelTemplate.R:174)
    model = pd.newPsydiff(**synth.config_c1(N=8, T=12))
    pd.compileModel(model)
    a = pd.fitModel(model)["m2LL"].copy()
    model["integrateFunction"] = "runge_kutta_dopri5"
    b = pd.fitModel(model)["m2LL"].copy()
    assert not np.array_equal(a, b) and np.allclose(a, b, rtol=1e-5)
    ref = _oracle(model).fit()["m2LL"]
    assert np.max(np.abs(b - ref) / np.abs(ref)) < M2LL_RTOL
    model["integrateFunction"] = "default"                  # rk4 (+ dopri5 only on a non-finite state)
    assert np.array_equal(pd.fitModel(model)["m2LL"], a)


def test_get_gradient_single_parameter_and_parallel_alias():
    model = pd.newPsydiff(**synth.config_c1(N=6, T=10))
    pd.compileParallelGradients(model, nCores=2)
    assert model["cl"] is not None
    g = pd.getGradientsParallel(model)
    vals = pd.getParameterValues(model)
    for par in (0, 5, 10):
        one = pd.getGradient(model, vals, par)
        (lab, v), = one.items()
        assert lab == list(g.keys())[par] and v == g[lab]
    pd.stopParallelGradients(model)


def test_golden_fixtures():
    import json, os
    here = os.path.dirname(os.path.abspath(__file__))
    g1 = json.load(open(os.path.join(here, "golden", "c1_oracle.json")))
    m = pd.newPsydiff(**synth.config_c1())
    pd.compileModel(m)
    f = pd.fitModel(m)
    assert np.max(np.abs(f["m2LL"] - np.array(g1["m2LL"])) / np.abs(g1["m2LL"])) < M2LL_RTOL
    assert np.max(np.abs(f["m2LL"] - np.array(g1["exact_kf_m2LL"])) / np.abs(g1["m2LL"])) < 1e-9   # exact Kalman filter
    assert np.allclose(f["latentScores"][:50], np.array(g1["latentScores_first_person"]), atol=1e-9)
    gv = np.array(list(pd.getGradients(m).values()))
    assert np.allclose(gv, g1["grad_clean"], rtol=1e-6, atol=1e-5)
    g4 = json.load(open(os.path.join(here, "golden", "c4_oracle_ext.json")))
    m4 = pd.newPsydiff(**synth.config_c4(N=16, T=12))
    pd.compileModel(m4)
    f4 = pd.fitModel(m4)["m2LL"]
    assert np.max(np.abs(f4 - np.array(g4["m2LL"])) / np.abs(g4["m2LL"])) < 1e-10
    gv4 = np.array(list(pd.getGradients(m4).values()))
    assert np.max(np.abs(gv4 - np.array(g4["grad_clean"])) / np.abs(g4["grad_clean"])) < GRAD_RTOL


def test_full_size_c2_properties():
    """C2 at BASELINE.json's size (N = 100k, T = 100, 5 % missing): size-independent properties + an oracle subsample."""
    N, T = 100_000, 100
    kw = synth.config_c2(N=N, T=T)
    obs = kw["dataset"]["observations"]
    obs[(N - 1) * T:] = obs[:T]              # the last person is a copy of the first one
    model = pd.newPsydiff(**kw)
    pd.compileModel(model)
    f = pd.fitModel(model)["m2LL"]
    assert np.isfinite(f).all() and f[0] == f[N - 1]          # identical persons -> bit-identical -2LL
    g = pd.getGradients(model)
    # oracle on a fixed 4096-person subsample of the same model, all host threads
    sub = np.arange(0, N, N // 4096)[:4096]
    ref = _oracle(model).fit(persons=sub, states=False, nthreads=os.cpu_count() or 1)["m2LL"]
    assert np.max(np.abs(f[sub] - ref) / np.abs(ref)) < M2LL_RTOL
    # ... and the gradient of that subsample as its own model (Σ over 4096 persons of every parameter's central difference)
    ks = dict(kw)
    rows = (sub[:, None] * T + np.arange(T)[None, :]).reshape(-1)
    ks["dataset"] = {k: v[rows] for k, v in kw["dataset"].items()}
    _check_grad(pd.newPsydiff(**ks), nthreads=os.cpu_count() or 1)
    # linearity over person shards: Σ-2LL and the gradient of the halves add up to the whole
    halves = []
    for lo, hi in ((0, N // 2), (N // 2, N)):
        k2 = dict(kw)
        k2["dataset"] = {k: v[lo * T:hi * T] for k, v in kw["dataset"].items()}
        mh = pd.newPsydiff(**k2)
        pd.compileModel(mh)
        halves.append((pd.fitModel(mh)["m2LL"], np.array(list(pd.getGradients(mh).values()))))
    assert np.array_equal(np.concatenate([halves[0][0], halves[1][0]]), f)
    gv = np.array(list(g.values()))
    assert np.allclose(halves[0][1] + halves[1][1], gv, rtol=1e-7, atol=1e-6 * np.abs(gv).max())


@pytest.mark.parametrize("K", [1, 4, 5])
def test_coupled_oscillator_family_n2_n8_n10(K):
    """n = 2, 8 (C5's model) and 10 latent states through the same kernel template."""
    model = pd.newPsydiff(**synth.config_coupled(K=K, N=6, T=8, n_groups=2) if K > 1 else synth.config_c3(N=6, T=8, integrateFunction="rk4"))
    pd.compileModel(model)
    fit = pd.fitModel(model)
    ext = _oracle(model, extended=True).fit(nthreads=8)
    assert np.isfinite(fit["m2LL"]).all()
    assert np.max(np.abs(fit["m2LL"] - ext["m2LL"]) / np.abs(ext["m2LL"])) < 1e-10
    assert np.nanmax(np.abs(fit["latentScores"] - ext["latentScores"])) < 1e-9
    if K == 4:
        _check_grad(model, extended=True)


def test_eps_fallback_mixed_levels_match_oracle():
    """eps = c(30, 1e-6): the +30 step on the drift diagonal makes the filter explode (non-finite Σ-2LL), so those sides fall
    back to the second eps while the other sides resolve at the first — the denominator is the sum of the eps actually
    used on each side (R/modelTemplate.R:885-898, 905-919, 926).  Exercises the restricted second launch."""
    # A0 fixed: a ±30 step on a log-variance of the INITIAL covariance (1e±13) is decided by cancellation in the first
    # downdate, where the NaN pattern legitimately depends on rounding; the drift / mean / diffusion steps are clean.
    model = pd.newPsydiff(**synth.config_c1(N=10, T=50, eps=np.array([30.0, 1e-6]), A0=np.eye(2)))
    pd.compileModel(model)
    g = pd.getGradients(model)
    ref = _oracle(model).gradients()
    gv, gr = np.array(list(g.values())), ref["grad_clean"]
    assert np.array_equal(np.isfinite(gv), np.isfinite(gr))
    labels = list(g)
    # at least one parameter must actually have used the fallback on one side only
    one = pd.newPsydiff(**synth.config_c1(N=10, T=50, eps=np.array([30.0]), A0=np.eye(2)))
    pd.compileModel(one)
    g1 = np.array(list(pd.getGradients(one).values()))
    assert np.isnan(g1).any() and not np.isnan(g1).all()
    fin = np.isfinite(gr)
    assert np.allclose(gv[fin], gr[fin], rtol=1e-6, atol=1e-6 * np.abs(gr[fin]).max()), list(zip(labels, gv, gr))


def test_sharded_api_on_one_rank_equals_getGradients():
    from psydiff_b200 import dist as pdist
    model = pd.newPsydiff(**synth.config_c1(N=12, T=15))
    pd.compileModel(model)
    a = pd.getGradients(model)
    b, tot = pdist.getGradientsSharded(model, return_total=True)
    assert a == b
    assert tot == pytest.approx(float(pd.fitModel(model)["m2LL"].sum()), rel=1e-13)
    assert pdist.fitTotalSharded(model) == pytest.approx(tot, rel=1e-13)


def test_edge_cases_single_row_all_missing_person_and_abi_errors():
    import ctypes as C
    from psydiff_b200 import _lib
    # one person with a single time point (dt = 0: no integration at all), one person with every observation missing
    kw = synth.config_c1(N=3, T=6, seed=13)
    d = kw["dataset"]
    keep = np.r_[0, np.arange(6, 18)]                      # person 1 keeps only its first row
    d = {k: v[keep] for k, v in d.items()}
    d["observations"][1:7] = np.nan                          # person 2: nothing observed
    kw["dataset"] = d
    model = pd.newPsydiff(**kw)
    fit, ref = _check_fit(model)
    assert fit["m2LL"][1] == 0.0 and np.isfinite(fit["m2LL"]).all()
    # with no update the latent scores of person 2 are the pure prediction
    _check_grad(model)
    # C-ABI argument errors are return codes with a message, never a crash
    L, h = _lib.lib(), model["_handle"].ptr
    theta = np.ascontiguousarray(model["pars"]["parameterTable"].theta_values)
    bad_idx = np.full(model["pars"]["parameterTable"].theta_index.shape, 99, dtype=np.int32)
    assert L.psy_set_slots(h, len(theta), _lib.iptr(bad_idx)) == 1 and b"out of range" in L.psy_last_error()
    good = np.ascontiguousarray(model["pars"]["parameterTable"].theta_index, dtype=np.int32)
    assert L.psy_set_slots(h, len(theta), _lib.iptr(good)) == 0
    obs = np.asfortranarray(d["observations"])
    off = np.array([0, 1, 7, 12], dtype=np.int64)            # does not end at rows
    pid = np.array([1, 2, 3], dtype=np.int32)
    assert L.psy_upload(h, obs.shape[0], 3, obs.ctypes.data_as(_lib.c_double_p), _lib.dptr(d["dt"]), _lib.i64ptr(off), _lib.iptr(pid)) == 1
    g = np.zeros(len(theta))
    eps = np.array([1e-6])
    assert L.psy_gradients(h, _lib.dptr(theta), _lib.dptr(eps), 1, 7, _lib.dptr(g), None) == 1
    assert b"Unknown direction" in L.psy_last_error()
    model["direction"] = "sideways"
    with pytest.raises(pd.PsydiffError, match="Unknown direction"):
        pd.getGradients(model)


@pytest.mark.parametrize("name", ["c1", "c1_missing_cov", "c3", "c4", "c4_T100", "n8"])
def test_reference_run_fixtures(name):
    """The CUDA path against the committed outputs of the REFERENCE'S OWN C++ (tests/golden/*_reference.json: fitModel /
    getGradients of R/modelTemplate.R + inst/include/psydiff*.h, compiled and run by oracle/reference_build.py where
    /root/reference exists).  -2LL and states within 1e-9; FD gradient within 1e-6 plus the double-precision FD noise floor."""
    import json, os
    import reference_fixtures as rf
    ref = rf.load(name)
    model = pd.newPsydiff(**rf.CASES[name]())
    pd.compileModel(model)
    assert model["pars"]["parameterTable"].theta_labels == ref["labels"]
    rf.check_fit(pd.fitModel(model), ref, rtol=M2LL_RTOL)
    if "getGradients" in ref:
        rf.check_gradient(np.array(list(pd.getGradients(model).values())), ref)


@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_subsample_4096_persons_at_benchmark_series_length(name):
    """SURVEY §8(d)'s gate for the configs bench.py times: the CUDA path on a shard of the benchmark workload (T = 100, the
    benchmark's stepper, time step and grouping) against the oracle on a fixed 4096-person subsample with all host threads
    (-2LL within 1e-9 per person, identical non-finite pattern), and the FD gradient of a sub-model of its first persons against the 80-bit
    oracle (1e-6 + FD noise floor; the double oracle's own distance is what DESIGN.md §5 explains).  C2 is covered at its full size
    by test_full_size_c2_properties."""
    nth = os.cpu_count() or 1
    N, T = {"c3": 32768, "c4": 8192, "c5": 4096}[name], 100
    kw = {"c3": synth.config_c3, "c4": synth.config_c4, "c5": synth.config_c5}[name](N=N, T=T)
    model = pd.newPsydiff(**kw)
    pd.compileModel(model)
    f = pd.fitModel(model, states=False)["m2LL"]
    sub = np.arange(0, N, N // 4096)[:4096]
    ref = _oracle(model).fit(persons=sub, states=False, nthreads=nth)["m2LL"]
    assert np.array_equal(np.isfinite(ref), np.isfinite(f[sub]))
    fin = np.isfinite(ref)
    rel = np.abs(f[sub][fin] - ref[fin]) / np.abs(ref[fin])
    print(f"{name}: 4096-person subsample of {N} x T={T}: max rel -2LL error vs oracle {rel.max():.3e}, non-finite persons {int((~fin).sum())}")
    assert rel.max() < M2LL_RTOL and fin.mean() > 0.99
    n_g = {"c3": 256, "c4": 48, "c5": 16}[name]          # the 80-bit oracle runs 2P+1 sweeps per person: sized for about 20 s on 16 threads
    rows = (sub[:n_g, None] * T + np.arange(T)[None, :]).reshape(-1)
    ks = dict(kw)
    ks["dataset"] = {k: v[rows] for k, v in kw["dataset"].items()}
    _check_grad(pd.newPsydiff(**ks), extended=True, nthreads=nth)


def test_lane_pair_kernel_equals_thread_per_instance_kernel_on_ragged_persons():
    """psy_filter_pair.cuh on the device (n = 8 and n = 10 by default, n = 6 on request): persons of different length and
    irregular dt inside one warp, instance lists with and without warp alignment, against the thread-per-instance kernel of the
    same model (-DPSY_PAIR=0) and the oracle."""
    from psydiff_b200 import _lib
    for K, flags in ((4, ""), (5, ""), (3, "-DPSY_PAIR=1")):
        kw = synth.config_coupled(K=K, N=96, T=12, n_groups=3, seed=7 + K)
        keep = np.ones(96 * 12, dtype=bool)
        for p in range(0, 96, 5):
            keep[p * 12 + 1 + (p % 9):(p + 1) * 12] = False          # series of 1..9 rows
        kw["dataset"] = {k: v[keep] for k, v in kw["dataset"].items()}
        pair, one = pd.newPsydiff(**kw), pd.newPsydiff(**kw)
        pd.compileModel(pair, extra_flags=flags)
        pd.compileModel(one, extra_flags="-DPSY_PAIR=0")
        L = _lib.lib()
        assert L.psy_kernel_lanes(pair["_handle"].ptr) == 2 and L.psy_kernel_lanes(one["_handle"].ptr) == 1
        fa, fb = pd.fitModel(pair), pd.fitModel(one)
        assert np.max(np.abs(fa["m2LL"] - fb["m2LL"]) / np.abs(fb["m2LL"])) < 1e-12
        assert np.nanmax(np.abs(fa["latentScores"] - fb["latentScores"])) < 1e-11
        ga, gb = (np.array(list(pd.getGradients(m).values())) for m in (pair, one))
        ftot = float(np.abs(np.sum(fb["m2LL"])))
        assert np.max(np.abs(ga - gb)) < 64 * np.finfo(float).eps * ftot / 2e-6 + 1e-9 * np.max(np.abs(gb))
        # the oracle in 80 bits (at n >= 8 the double build's own rounding noise reaches the 1e-9 gate on short series)
        ext = _oracle(pair, extended=True).fit(nthreads=os.cpu_count() or 1)
        assert np.max(np.abs(fa["m2LL"] - ext["m2LL"]) / np.abs(ext["m2LL"])) < 1e-10
        assert np.nanmax(np.abs(fa["latentScores"] - ext["latentScores"])) < 1e-9


def test_stage_unrolling_and_the_lean_dopri5_controller_on_the_device():
    """The late round-2 kernel defaults against their plain forms, on the device: four RK4 stages per loop body (n <= 4) against the
    rolled stage loop, and the adaptive stepper's step-size controller without libdevice pow / IEEE division against odeint's
    own formulas (-DPSY_DOPRI_LEAN=0) — same filter to rounding, both within the gate against the oracle."""
    for kw in (synth.config_c1(N=64, T=20), synth.config_coupled(K=2, N=32, T=10, n_groups=2)):
        a, b = pd.newPsydiff(**kw), pd.newPsydiff(**kw)
        pd.compileModel(a)
        pd.compileModel(b, extra_flags="-DPSY_STAGE_UNROLL=1")
        fa, fb = pd.fitModel(a)["m2LL"], pd.fitModel(b)["m2LL"]
        assert np.max(np.abs(fa - fb) / np.abs(fb)) < 1e-13
        ga, gb = (np.array(list(pd.getGradients(m).values())) for m in (a, b))
        assert np.max(np.abs(ga - gb)) <= 64 * np.finfo(float).eps * float(np.abs(fb.sum())) / 2e-6 + 1e-9 * np.max(np.abs(gb))
    for kw in (synth.config_c3(N=64, T=24), synth.config_c1(N=32, T=20, integrateFunction="runge_kutta_dopri5")):
        lean, plain = pd.newPsydiff(**kw), pd.newPsydiff(**kw)
        pd.compileModel(lean)
        pd.compileModel(plain, extra_flags="-DPSY_DOPRI_LEAN=0")
        fl, fp = pd.fitModel(lean)["m2LL"], pd.fitModel(plain)["m2LL"]
        assert np.max(np.abs(fl - fp) / np.abs(fp)) < 1e-11
        ref = _oracle(lean).fit()["m2LL"]
        assert np.max(np.abs(fl - ref) / np.abs(ref)) < M2LL_RTOL and np.max(np.abs(fp - ref) / np.abs(ref)) < M2LL_RTOL


def test_several_devices_in_one_handle_equal_one_device():
    """psy_model_set_devices: persons cut into work-balanced blocks over the GPUs of the box, one host thread.  Per-person results
    (-2LL, latent scores, predictions) are independent of the cut — bit-identical; parameter sums differ by summation order only
    (<= 1e-12); the eps[] fallback and one-sided sweeps run on every device; going back to one device reproduces the first result
    bit for bit.  Needs >= 2 GPUs (gpurun --gpus 2)."""
    if pd.deviceCount() < 2:
        pytest.skip("needs at least two GPUs")
    nd = min(pd.deviceCount(), 8)
    # ragged persons with person-specific parameters: the cuts are placed by work, not by count
    kw = synth.config_c3(N=1500, T=30, integrateFunction="rk4")
    keep = np.ones(1500 * 30, dtype=bool)
    for p in range(0, 700, 2):
        keep[p * 30 + 8:(p + 1) * 30] = False                 # the first persons have short series
    kw["dataset"] = {k: v[keep] for k, v in kw["dataset"].items()}
    for direction, eps in (("central", np.array([1e-6])), ("left", np.array([1e-6])), ("central", np.array([30.0, 1e-6]))):
        model = pd.newPsydiff(**dict(kw, direction=direction, eps=eps))
        pd.compileModel(model)
        one = pd.fitModel(model)
        g1 = np.array(list(pd.getGradients(model).values()))
        t1, b1 = pd.fitTotals(model, [pd.getParameterValues(model)] * 2)
        used = pd.setDevices(model, nd)
        assert used == list(range(nd)) or len(used) == nd
        info = pd.shardInfo(model)
        assert sum(s["persons"] for s in info) == 1500 and info[0]["persons"] > info[-1]["persons"]     # short series first: more persons there
        many = pd.fitModel(model)
        for k in ("m2LL", "latentScores", "predictedManifest"):
            assert np.array_equal(one[k], many[k], equal_nan=True), k
        gn = np.array(list(pd.getGradients(model).values()))
        assert np.array_equal(np.isfinite(g1), np.isfinite(gn))
        fin = np.isfinite(g1)
        assert np.max(np.abs(gn[fin] - g1[fin]) / np.maximum(np.abs(g1[fin]), 1e-12)) < 1e-12
        tn, bn = pd.fitTotals(model, [pd.getParameterValues(model)] * 2)
        assert np.allclose(tn, t1, rtol=1e-13) and np.array_equal(bn, b1)
        assert pd.setDevices(model, 1) == used[:1]
        assert np.array_equal(np.array(list(pd.getGradients(model).values())), g1, equal_nan=True)
    # the oracle as the judge of the sharded result
    model = pd.newPsydiff(**synth.config_c4(N=64, T=16))
    pd.compileModel(model)
    pd.setDevices(model, nd)
    _check_fit(model)
    _check_grad(model, extended=True)
