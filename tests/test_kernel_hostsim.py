"""The kernel SOURCE on the CPU: psy_filter.cuh + the generated model TU compiled by g++ (tests/hostsim) and checked against the
oracle with the gates of tests/test_gpu_parity.py.  Runs without a GPU.  This is test infrastructure — it checks the kernel's
arithmetic, indexing and instance-list conventions; device parity proper is tests/test_gpu_parity.py (-m gpu).
"""
import json
import os

import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import synth
from hostsim.hostsim import HostSim

M2LL_RTOL = 1e-9
GRAD_RTOL = 1e-6


def _oracle(model, extended=False):
    from oracle.oracle import Oracle
    return Oracle(model, extended=extended)


def _check_fit(model, skipUpdate=False, flags=""):
    hs = HostSim(model, flags)
    fit = hs.fit(skip_update=skipUpdate)
    ref = _oracle(model).fit(skip_update=skipUpdate)
    fin = np.isfinite(ref["m2LL"])
    assert np.array_equal(fin, np.isfinite(fit["m2LL"]))
    if skipUpdate:
        assert np.all(fit["m2LL"] == 0.0)
    else:
        rel = np.abs(fit["m2LL"][fin] - ref["m2LL"][fin]) / np.maximum(np.abs(ref["m2LL"][fin]), 1e-300)
        assert rel.max() < M2LL_RTOL, rel.max()
    scale = max(1.0, np.nanmax(np.abs(ref["latentScores"])))
    assert np.nanmax(np.abs(fit["latentScores"] - ref["latentScores"])) < 1e-9 * scale
    assert np.nanmax(np.abs(fit["predictedManifest"] - ref["predictedManifest"])) < 1e-9 * scale
    return hs, fit, ref


def _check_grad(model, hs=None, extended=False, flags=""):
    hs = hs or HostSim(model, flags)
    eps = float(np.atleast_1d(model["eps"])[0])
    g = hs.gradients(eps=eps)
    ref = _oracle(model, extended).gradients(direction="central", nthreads=8)
    gv, gr = g["grad"], ref["grad_clean"]
    assert np.array_equal(np.isfinite(gv), np.isfinite(gr))
    ftot = float(np.abs(np.nansum(ref["m2LL"])))
    atol = 64 * np.finfo(float).eps * ftot / (2 * eps)       # rounding floor of the finite difference itself
    bad = np.abs(gv - gr) > GRAD_RTOL * np.abs(gr) + atol
    labels = model["pars"]["parameterTable"].theta_labels
    assert not bad.any(), [(labels[i], gv[i], gr[i]) for i in np.flatnonzero(bad)]
    return gv, gr


def test_module_info_matches_model():
    m = pd.newPsydiff(**synth.config_c4(N=4, T=4))
    hs = HostSim(m)
    assert hs.info[:5] == [0, 1, 6, 6, 30] and hs.info[5] == 0          # rk4, SR, n, m, slots, no shared memory at n = 6
    m8 = pd.newPsydiff(**synth.config_coupled(K=4, N=4, T=4))
    h8 = HostSim(m8)
    # n = 8: the lane-pair kernel; x_n and the running sum of ONE lane (m on both lanes + own columns: 8 + 20 doubles) parked in shared memory
    assert h8.info[2] == 8 and h8.lanes == 2 and h8.block == 64 and h8.info[5] == 2 * (8 + 20) * h8.block * 8
    h8s = HostSim(m8, "-DPSY_PAIR=0")
    assert h8s.lanes == 1 and h8s.info[5] == 2 * (8 + 36) * h8s.block * 8   # thread per instance: RK4 state parked for n >= 7
    m7 = pd.newPsydiff(**synth.config_c4(N=4, T=4))
    assert HostSim(m7, "-DPSY_PAIR=1").lanes == 2 and hs.lanes == 1        # n = 6: pairs only on request


def test_c1_fit_states_gradient_and_golden():
    m = pd.newPsydiff(**synth.config_c1())
    hs, fit, _ = _check_fit(m)
    gv, _ = _check_grad(m, hs)
    here = os.path.dirname(os.path.abspath(__file__))
    g1 = json.load(open(os.path.join(here, "golden", "c1_oracle.json")))
    assert np.max(np.abs(fit["m2LL"] - np.array(g1["m2LL"])) / np.abs(g1["m2LL"])) < M2LL_RTOL
    assert np.max(np.abs(fit["m2LL"] - np.array(g1["exact_kf_m2LL"])) / np.abs(g1["m2LL"])) < 1e-9
    assert np.allclose(gv, g1["grad_clean"], rtol=1e-6, atol=1e-5)


def test_c1_skip_update():
    _check_fit(pd.newPsydiff(**synth.config_c1(N=8, T=12)), skipUpdate=True)


def test_missing_ragged_single_row_and_all_missing_person():
    kw = synth.config_c1(N=40, T=30, missing=0.25, seed=7)
    obs = kw["dataset"]["observations"]
    obs[5] = np.nan
    obs[31:35] = np.nan
    keep = np.ones(len(obs), dtype=bool)
    for p in (3, 17, 22):
        keep[p * 30 + 20:(p + 1) * 30] = False
    keep[30 * 30 + 1:31 * 30] = False                   # person 30: a single row (dt = 0, no integration)
    obs[33 * 30:34 * 30] = np.nan                       # person 33: nothing observed
    for k in ("person", "observations", "dt"):
        kw["dataset"][k] = kw["dataset"][k][keep]
    model = pd.newPsydiff(**kw)
    hs, fit, _ = _check_fit(model)
    assert fit["m2LL"][33] == 0.0
    _check_grad(model, hs)


def test_irregular_dt_remainder_dropped():
    kw = synth.config_c1(N=16, T=25, seed=11)
    rng = np.random.default_rng(5)
    dt = rng.choice([0.01, 0.2, 0.5, 0.77, 1.0, 1.31], size=16 * 25)
    dt[::25] = 0.0
    kw["dataset"]["dt"] = dt
    kw["timeStep"] = np.array([1.0 / 30])
    _check_fit(pd.newPsydiff(**kw))


def test_c4_against_double_and_80bit_oracle_and_golden():
    model = pd.newPsydiff(**synth.config_c4(N=16, T=12))
    hs, fit, ref = _check_fit(model)
    ext = _oracle(model, extended=True).fit(nthreads=8)
    rel_k = np.abs(fit["m2LL"] - ext["m2LL"]) / np.abs(ext["m2LL"])
    assert rel_k.max() < 1e-10
    gv, gr = _check_grad(model, hs, extended=True)
    here = os.path.dirname(os.path.abspath(__file__))
    g4 = json.load(open(os.path.join(here, "golden", "c4_oracle_ext.json")))
    assert np.max(np.abs(fit["m2LL"] - np.array(g4["m2LL"])) / np.abs(g4["m2LL"])) < 1e-10
    # rounding floor of a finite difference of Σ-2LL (same allowance as _check_grad; g++'s FMA contraction pattern differs
    # from nvcc's, so the host build sits elsewhere inside that floor than the device build)
    atol = 64 * np.finfo(float).eps * float(np.abs(np.sum(g4["m2LL"]))) / 2e-6
    assert np.all(np.abs(gv - np.array(g4["grad_clean"])) < GRAD_RTOL * np.abs(g4["grad_clean"]) + atol)


def test_nan_pattern_of_a_failing_person():
    model = pd.newPsydiff(**synth.config_c1(N=6, T=10))
    pd.setParameterValues(model["pars"]["parameterTable"], {"drift_eta1": 400.0})
    fit = HostSim(model).fit()
    ref = _oracle(model).fit()
    assert np.array_equal(np.isfinite(fit["m2LL"]), np.isfinite(ref["m2LL"]))


def test_covariance_variant():
    model = pd.newPsydiff(**synth.config_c1(N=24, T=30, missing=0.2, seed=9, srUpdate=False))
    hs, _, _ = _check_fit(model)
    assert hs.info[1] == 0
    _check_grad(model, hs)
    _check_fit(pd.newPsydiff(**synth.config_c4(N=8, T=10, srUpdate=False)))


def test_dopri5_linear_and_c3():
    _check_fit(pd.newPsydiff(**synth.config_c1(N=16, T=20, integrateFunction="runge_kutta_dopri5")))
    c3 = pd.newPsydiff(**synth.config_c3(N=32, T=24))
    hs, fit, _ = _check_fit(c3)
    assert hs.info[0] == 1 and np.isfinite(fit["m2LL"]).all()
    _check_grad(c3, hs, extended=True)


@pytest.mark.parametrize("K", [1, 4, 5])
def test_coupled_oscillator_family_n2_n8_n10(K):
    model = pd.newPsydiff(**synth.config_coupled(K=K, N=6, T=8, n_groups=2) if K > 1 else synth.config_c3(N=6, T=8, integrateFunction="rk4"))
    hs = HostSim(model)
    fit = hs.fit()
    ext = _oracle(model, extended=True).fit(nthreads=8)
    assert np.isfinite(fit["m2LL"]).all()
    assert np.max(np.abs(fit["m2LL"] - ext["m2LL"]) / np.abs(ext["m2LL"])) < 1e-10
    assert np.nanmax(np.abs(fit["latentScores"] - ext["latentScores"])) < 1e-9
    if K == 4:
        _check_grad(model, hs, extended=True)


@pytest.mark.parametrize("K", [2, 3, 4])
def test_lane_pair_kernel_ragged_unaligned_and_against_the_thread_per_instance_kernel(K):
    """psy_filter_pair.cuh (two lanes integrate one instance, every lane owns one): persons of different length, series of one
    row, irregular dt (pairs of one simulated warp need different step counts: the restore path), instance lists without warp
    alignment, padding slots — against the oracle and against the thread-per-instance kernel, for every parking level."""
    kw = synth.config_coupled(K=K, N=7, T=9, n_groups=2, seed=99 + K)
    obs = kw["dataset"]["observations"]
    keep = np.ones(len(obs), dtype=bool)
    keep[1 * 9 + 4:2 * 9] = False              # person 1: 4 rows
    keep[3 * 9 + 1:4 * 9] = False              # person 3: a single row (no integration at all)
    keep[4 * 9 + 7:5 * 9] = False              # person 4: 7 rows
    obs[5 * 9 + 2] = np.nan                    # a time point without observations
    obs[2 * 9 + 3, 0] = np.nan
    for k in ("person", "observations", "dt"):
        kw["dataset"][k] = kw["dataset"][k][keep]
    model = pd.newPsydiff(**kw)
    hs, fit, _ = _check_fit(model, flags="-DPSY_PAIR=1")
    assert hs.lanes == 2
    one = HostSim(model, "-DPSY_PAIR=0")
    assert one.lanes == 1
    ref = one.gradients(pad=False)
    for flags in ("-DPSY_PAIR=1 -DPSY_PAIR_PARK=0", "-DPSY_PAIR=1 -DPSY_PAIR_PARK=1", "-DPSY_PAIR=1 -DPSY_PAIR_PARK=2"):
        for pad in (False, True):
            g = HostSim(model, flags).gradients(pad=pad)
            assert np.max(np.abs(g["m2LL"] - ref["m2LL"]) / np.abs(ref["m2LL"])) < 1e-12
            fa = g["f"][np.isfinite(g["f"])] if pad else g["f"]
            assert np.max(np.abs(fa - ref["f"]) / np.abs(ref["f"])) < 1e-12
    # parameter sets (psy_fit_many) through the pair kernel
    th = hs.theta()
    sets = np.stack([th, th * 1.01])
    many = hs.fit_many(sets)
    for k in range(2):
        assert np.array_equal(many[k], hs.fit(theta=sets[k], states=False)["m2LL"])


def test_instance_list_padding_and_parameter_sets_do_not_change_results():
    """Warp-alignment padding slots (person = -1) are skipped, and instance results do not depend on where an instance sits:
    padded and unpadded lists give bit-identical numbers; psy_fit_many's set k equals psy_fit at thetas[k]."""
    model = pd.newPsydiff(**synth.config_c4(N=5, T=6))
    hs = HostSim(model)
    a, b = hs.gradients(pad=True), hs.gradients(pad=False)
    assert np.array_equal(a["grad"], b["grad"]) and np.array_equal(a["m2LL"], b["m2LL"])
    assert len(a["f"]) % 32 == 0 and len(b["f"]) == 5 * 61
    th = hs.theta()
    sets = np.stack([th, th * 1.01, th * 0.99])
    many = hs.fit_many(sets)
    for k in range(3):
        assert np.array_equal(many[k], hs.fit(theta=sets[k], states=False)["m2LL"])


@pytest.mark.parametrize("flags", ["-DPSY_RHS_SCALED=0", "-DPSY_RK4_SMEM=1", "-DPSY_RK4_SMEM=0", "-DPSY_IEEE_DIV", "-DPSY_RSQRT_ROT",
                                   "-DPSY_STAGE_UNROLL=2", "-DPSY_STAGE_UNROLL=4", "-DPSY_STEP_UNROLL=2"])
def test_compile_time_variants_agree(flags):
    """Every compile-time variant of the kernel computes the same filter (to rounding)."""
    for kw in (synth.config_c4(N=4, T=8), synth.config_coupled(K=4, N=3, T=6)):
        model = pd.newPsydiff(**kw)
        a = HostSim(model).fit()["m2LL"]
        b = HostSim(model, flags).fit()["m2LL"]
        assert np.max(np.abs(a - b) / np.abs(a)) < 1e-11


def test_stage_unrolling_and_the_lean_dopri5_controller_change_nothing_but_rounding():
    """RK4 stages per loop body: the default is 4 for n <= 4 and 1 above; 1, 2 and 4 give the same filter at n = 2 (bit for bit on
    the host, where nothing is re-associated).  The adaptive stepper with odeint's own pow() and '/' in the step-size controller
    (-DPSY_DOPRI_LEAN=0) against the default branch-free reciprocal: same accepted steps, -2LL equal to rounding."""
    model = pd.newPsydiff(**synth.config_c1(N=6, T=12))
    a = HostSim(model).fit()["m2LL"]
    for flags in ("-DPSY_STAGE_UNROLL=1", "-DPSY_STAGE_UNROLL=2"):
        assert np.array_equal(a, HostSim(model, flags).fit()["m2LL"])
    for kw in (synth.config_c1(N=6, T=12, integrateFunction="runge_kutta_dopri5"), synth.config_c3(N=6, T=12)):
        model = pd.newPsydiff(**kw)
        lean, plain = HostSim(model).fit()["m2LL"], HostSim(model, "-DPSY_DOPRI_LEAN=0").fit()["m2LL"]
        assert np.max(np.abs(lean - plain) / np.abs(plain)) < 1e-12


@pytest.mark.parametrize("name", ["c1", "c1_missing_cov", "c3", "c4", "c4_T100", "n8"])
def test_kernel_source_equals_reference_run_fixtures(name):
    """The kernel source (host-compiled) against the committed outputs of the reference's own C++ (oracle/_ref build)."""
    import reference_fixtures as rf
    ref = rf.load(name)
    model = pd.newPsydiff(**rf.CASES[name]())
    hs = HostSim(model)
    rf.check_fit(hs.fit(), ref)
    if "getGradients" in ref:
        rf.check_gradient(hs.gradients(eps=1e-6)["grad"], ref)


def operator_surface_model(**settings):
    """Every statement form psy_lang.cuh documents: matrix*vector, matrix containers as vectors, %, element-wise /, unary minus,
    scalar scaling and shifting, arma::exp / dot / accu / trans, x(i) and x[i], t (interval-local in the drift, absolute in the
    measurement equation) and person."""
    rng = np.random.default_rng(8)
    N, T = 5, 9
    obs = 0.5 + rng.standard_normal((N * T, 2))
    obs[rng.random(obs.shape) < 0.1] = np.nan
    kw = dict(
        dataset={"person": np.repeat(np.arange(3, 3 + N), T), "observations": obs, "dt": np.tile(np.r_[0.0, np.full(T - 1, 0.6)], N)},
        latentEquations=("dx = DRIFT*x + CINT;\n"
                         "dx(0) = dx(0) + 0.01*arma::dot(x, x);\n"
                         "dx = dx - 0.01*(x % x);\n"
                         "dx = -(-dx);\n"
                         "dx = dx + 0.001*arma::exp(-1.0*(x % x));\n"
                         "dx[1] = dx[1] + 0.01*arma::accu(x) + 0.02*t;\n"
                         "dx = dx/1.25;"),
        manifestEquations=("y = arma::trans(LT)*x + MM;\n"
                           "y = y + 0.01*(x/(x % x + 1.0));\n"
                           "y(0) = y(0) + 0.1*sin(0.3*t) + 0.001*person;"),
        parameters={"DRIFT": np.array([["d11 = -0.5", "0.1"], ["d21 = 0.2", "d22 = -0.6"]], dtype=object),
                    "CINT": np.array([["c1 = 0.3"], ["0.1"]], dtype=object),
                    "LT": np.array([["1", "l12 = 0.4"], ["0", "1"]], dtype=object),
                    "MM": np.array([["0.2"], ["mm2 = -0.1"]], dtype=object)},
        L=np.array([["l1 = 0.5", "0"], ["0.1", "l2 = 0.4"]], dtype=object),
        Rchol=np.array([["r1 = 0.6", "0"], ["0.1", "r2 = 0.5"]], dtype=object),
        A0=np.array([["0.8", "0"], ["0.1", "0.7"]], dtype=object), m0=np.array([["m1 = 1.0"], ["0.5"]], dtype=object),
        integrateFunction="rk4", eps=np.array([1e-6]))
    kw.update(settings)
    return kw


@pytest.mark.parametrize("settings", [{}, {"integrateFunction": "runge_kutta_dopri5"}, {"srUpdate": False}], ids=["rk4", "dopri5", "cov"])
def test_operator_surface_of_the_equation_language(settings):
    model = pd.newPsydiff(**operator_surface_model(**settings))
    hs, _, _ = _check_fit(model)
    if not settings:
        _check_grad(model, hs, extended=True)


def _sanitizer_script(flags):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return f"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth
from hostsim.hostsim import HostSim
flags = {flags!r}
ragged = synth.config_c1(N=5, T=7, missing=0.3)
keep = np.ones(35, dtype=bool); keep[7 + 1:14] = False; keep[30:35] = False          # a single-row person, a shorter last person
ragged["dataset"] = {{k: v[keep] for k, v in ragged["dataset"].items()}}
for kw in (ragged, synth.config_c4(N=3, T=5), synth.config_c3(N=3, T=5), synth.config_c4(N=2, T=4, srUpdate=False),
           synth.config_coupled(K=4, N=2, T=4), synth.config_coupled(K=5, N=2, T=3)):
    hs = HostSim(pd.newPsydiff(**kw), flags)
    assert np.isfinite(hs.fit()["m2LL"]).all() and np.isfinite(hs.gradients()["grad"]).all()
    assert np.isfinite(hs.gradients(pad=False)["grad"]).all()
    assert np.isfinite(hs.fit_many(np.stack([hs.theta(), hs.theta() * 1.01]))).all()
print("SANITIZER-CLEAN")
"""


def test_kernel_source_under_bounds_and_overflow_sanitizers():
    """compute-sanitizer is closed on the GPU pool; the kernel source compiled for the host runs under UBSan
    instead: every fixed-size array index (packed-triangular tri(i, j), slot arrays, the Vec/Mat of the equation language),
    signed overflow, shifts and null dereferences are checked while the filter runs for n = 2, 6, 8, 10, both steppers and both
    update variants, fit and gradient sweeps.  Runs in a subprocess: a finding aborts the process."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", _sanitizer_script("-fsanitize=bounds,signed-integer-overflow,shift,vla-bound,null -fno-sanitize-recover=all -O1")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SANITIZER-CLEAN" in r.stdout and "runtime error" not in r.stderr, r.stderr[-3000:]


def test_kernel_source_under_address_sanitizer():
    """The same runs under AddressSanitizer: every access the kernel makes into the launch buffers (observation rows, dt, step
    counts, row offsets, slot map, theta sets, instance list incl. padding slots, -2LL / state outputs) stays inside them."""
    import subprocess
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-c", _sanitizer_script("-fsanitize=address -fno-omit-frame-pointer -O1")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "SANITIZER-CLEAN" in r.stdout and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]


def vignette_expm_model(use_expm=True, free=True):
    """The vignette's "quite complex models" example (vignettes/psydiff-basics.Rmd:94-98): a typed Armadillo temporary holding the
    matrix exponential of a parameter matrix, here actually used in the drift.  use_expm=False supplies scipy's expm(A) as a
    constant matrix instead; free=False fixes every parameter (a model without free parameters)."""
    import scipy.linalg
    rng = np.random.default_rng(3)
    N, T = 4, 8
    ds = {"person": np.repeat(np.arange(1, N + 1), T), "observations": rng.standard_normal((N * T, 2)),
          "dt": np.tile(np.r_[0.0, np.full(T - 1, 0.5)], N)}
    A = np.array([[-0.4, 0.3], [0.1, -0.7]])
    common = dict(dataset=ds, manifestEquations="y = x;", L=np.diag([0.5, 0.4]), Rchol=np.diag([0.6, 0.5]), A0=np.diag([0.8, 0.7]),
                  m0=np.array([[1.0], [0.5]]), integrateFunction="rk4", eps=np.array([1e-6]), timeStep=np.array([1 / 20]))
    tail = "arma::colvec z = expOfA*x;\ndx[0] = -z[0] + 0.1*x[1];\ndx[1] = -0.5*z[1];"
    if use_expm:
        Apar = np.array([["a11 = -0.4", "0.3"], ["0.1", "a22 = -0.7"]], dtype=object) if free else A
        return dict(latentEquations="arma::mat expOfA = arma::expm(A);\n" + tail, parameters={"A": Apar}, **common)
    return dict(latentEquations=tail, parameters={"expOfA": scipy.linalg.expm(A)}, **common)


def test_vignette_expm_and_typed_temporaries():
    model = pd.newPsydiff(**vignette_expm_model())
    assert "auto expOfA = arma::expm(A);" in model["cppCode"] and "auto z = expOfA*x;" in model["cppCode"]
    hs, fit, _ = _check_fit(model)
    _check_grad(model, hs, extended=True)
    # the matrix exponential itself: the same filter with scipy's expm(A) supplied as a constant
    const = pd.newPsydiff(**vignette_expm_model(use_expm=False))
    assert len(const["pars"]["parameterTable"].theta_labels) == 0            # ... which is also a model without free parameters
    ref = HostSim(const).fit()
    assert np.max(np.abs(fit["m2LL"] - ref["m2LL"]) / np.abs(ref["m2LL"])) < 1e-13
    assert HostSim(const).gradients()["grad"].shape == (0,)
    fixed = pd.newPsydiff(**vignette_expm_model(free=False))
    assert np.array_equal(HostSim(fixed).fit()["m2LL"], fit["m2LL"])         # fixing the parameters at their values changes nothing


def elementwise_functions_model(**settings):
    """arma::sin / cos / tanh / sqrt / abs / square / pow on vectors, arma::norm / max / min / mean / as_scalar / sum."""
    kw = operator_surface_model(**settings)
    kw["latentEquations"] = ("dx = DRIFT*x + CINT;\n"
                             "dx = dx - 0.05*arma::tanh(x) + 0.02*arma::sin(x) - 0.01*arma::cos(2.0*x);\n"
                             "dx = dx - 0.01*arma::pow(arma::abs(x), 1.5) - 0.005*arma::square(x);\n"
                             "dx(0) = dx(0) - 0.01*arma::norm(x) + 0.003*arma::max(x) - 0.002*arma::min(x) + 0.004*arma::mean(x);\n"
                             "dx(1) = dx(1) + 0.01*arma::as_scalar(arma::trans(CINT)*x) + 0.001*arma::sum(x);")
    kw["manifestEquations"] = "y = arma::trans(LT)*x + MM;\ny = y + 0.05*arma::sqrt(arma::square(x) + 1.0);"
    return kw


def test_elementwise_and_reduction_functions_of_the_equation_language():
    model = pd.newPsydiff(**elementwise_functions_model())
    hs, _, _ = _check_fit(model)
    _check_grad(model, hs, extended=True)


def control_flow_model():
    """Plain C++ control flow in the statements: a counted loop over components, if / else on the state, a conditional expression
    on t, element access M(i, j) with a loop index."""
    kw = operator_surface_model()
    kw["latentEquations"] = ("for (int i = 0; i < 2; i++) { dx(i) = DRIFT(i,i)*x(i) + CINT(i); }\n"
                             "if (x(0) > 0) { dx(0) = dx(0) - 0.1*x(0)*x(0); } else { dx(0) = dx(0) + 0.05*x(1); }\n"
                             "double damp = (t > 0.3) ? 0.9 : 1.0;\n"
                             "dx(1) = damp*dx(1) + DRIFT(1,0)*x(0);")
    kw["manifestEquations"] = "y = arma::trans(LT)*x + MM;"
    return kw


def test_control_flow_in_the_statements():
    model = pd.newPsydiff(**control_flow_model())
    assert codegen_dep(model) == [3, 3]                    # not plain element statements: the sparsity analysis gives up
    _check_fit(model)


def codegen_dep(model):
    from psydiff_b200 import codegen
    return codegen.drift_dependency(model["_spec"])


def test_eps_fallback_sweep_on_the_cpu_with_the_library_bookkeeping():
    """eps = c(30, 1e-6) end to end without a GPU: instance results from the host-compiled kernel source, the partial-sum vector
    [Σf0, #nonfinite, D+, D-, nf+, nf-, nbt] formed as the reduction kernels form it, the library's host bookkeeping
    (psy_sweep_resolve / psy_sweep_finish) deciding which sides fall back to the second step — against the oracle's gradient."""
    import ctypes as C
    from psydiff_b200 import _lib
    model = pd.newPsydiff(**synth.config_c1(N=10, T=50, eps=np.array([30.0, 1e-6]), A0=np.eye(2)))
    hs = HostSim(model)
    P = hs.P
    L = _lib.lib()
    sweep = L.psy_sweep_create(P)
    need, done = np.zeros(2 * P + 1, dtype=np.uint8), C.c_int(0)
    used_fallback = False
    for level, eps in enumerate(model["eps"]):
        inst, start, distinct = hs.instance_list(pad=True)
        f = hs._launch(inst, len(inst), hs.theta(), float(eps))
        part = np.zeros(5 * P + 2)
        f0 = np.array([f[int(start[p])] for p in range(hs.N)])
        part[0], part[1] = f0[np.isfinite(f0)].sum() if not np.isfinite(f0).all() else f0.sum(), (~np.isfinite(f0)).sum()
        for p in range(hs.N):
            o = int(start[p])
            f0s = f0[p] if np.isfinite(f0[p]) else 0.0
            for j, t in enumerate(distinct[p]):
                fm, fp = f[o + 1 + 2 * j], f[o + 2 + 2 * j]
                part[2 + t] += fp - f0s                          # D+
                part[2 + P + t] += fm - f0s                      # D-
                part[2 + 2 * P + t] += 0.0 if np.isfinite(fp) else 1.0
                part[2 + 3 * P + t] += 0.0 if np.isfinite(fm) else 1.0
                part[2 + 4 * P + t] += 0.0 if np.isfinite(f0[p]) else 1.0
        if level > 0:
            used_fallback = bool(need[:2 * P].any())
        _lib.check(L.psy_sweep_resolve(sweep, _lib.dptr(part), float(eps), level, len(model["eps"]), 0, _lib.u8ptr(need), C.byref(done)))
        if done.value:
            break
    g, tot = np.zeros(P), C.c_double()
    _lib.check(L.psy_sweep_finish(sweep, 0, _lib.dptr(g), C.byref(tot)))
    L.psy_sweep_destroy(sweep)
    ref = _oracle(model).gradients()
    assert used_fallback                                        # some sides did fail at eps = 30 and were redone at 1e-6
    assert np.array_equal(np.isfinite(g), np.isfinite(ref["grad_clean"]))
    fin = np.isfinite(g)
    assert np.allclose(g[fin], ref["grad_clean"][fin], rtol=1e-6, atol=1e-6 * np.abs(ref["grad_clean"][fin]).max())
    assert tot.value == pytest.approx(float(ref["m2LL"].sum()), rel=1e-12)


def test_flop_model_against_a_dynamic_count_of_the_kernel_source():
    """psydiff_b200/flops.py's closed form (what bench.py's roofline starts from, before the ncu calibration) against the
    operations the kernel SOURCE really executes: the host build with `double` replaced by a counting wrapper
    (tests/hostsim/op_counter.h).  The dynamic count includes what the compiler later removes (drift components a sigma-point
    column cannot move, time bookkeeping), so it sits a few per cent ABOVE the closed form; ncu's executed count sits 9 % below
    it."""
    from psydiff_b200 import flops, codegen, _lib
    L = _lib.lib()
    model = pd.newPsydiff(**synth.config_c4(N=3, T=20))
    hs = HostSim(model, "-DPSY_COUNT_OPS -std=c++20")
    hs.ops()
    f = hs.fit(states=False)["m2LL"]
    ops = hs.ops()
    plain = HostSim(model).fit(states=False)["m2LL"]
    assert np.max(np.abs(f - plain) / np.abs(plain)) < 1e-12          # same arithmetic (the wrapper only keeps a*b+c from contracting)
    h = float(model["timeStep"][0])
    steps = sum(L.psy_rk4_step_count(float(d), h) for d in model["data"]["dt"])
    closed = flops.kernel_flops(6, 6, 39, 0, steps, len(model["data"]["dt"]), l_diag=True, dep=codegen.drift_dependency(model["_spec"]))
    assert ops["div"] == 0                                  # no IEEE division anywhere on the path (reciprocals are seeded + Newton)
    assert 1.0 <= ops["flops"] / closed <= 1.10, (ops, closed)
    assert ops["fma"] * 2 > 0.45 * ops["flops"]             # about half of the flops are written as explicit FMAs


def test_reduction_kernels_k3_on_the_cpu():
    """psy_reduce.cuh (K3) compiled for the host — psy_reduce_chunks with real threads and a barrier for __syncthreads() — against
    numpy: D± = Σ_p (f±_p − f0_p) with a non-finite f0 treated as 0, the non-finite counts per side, the base pseudo-segment,
    multi-chunk parameters (> 4096 persons), small segments (person-specific parameters), and bit-reproducibility."""
    from hostsim import k3sim
    rng = np.random.default_rng(0)
    N, F = 9000, 6
    n_common, n_group = 3, 5                         # 3 parameters shared by everybody (3 chunks each), 5 group labels, N person-specific
    P = n_common + n_group + N
    tidx = np.empty((N, F), dtype=np.int32)
    tidx[:, :3] = np.arange(n_common)
    tidx[:, 3] = n_common + rng.integers(0, n_group, N)
    tidx[:, 4] = tidx[:, 3]                          # a second slot carrying the same label: counted once per person
    tidx[:, 5] = n_common + n_group + np.arange(N)
    ls = k3sim.lists(tidx, P)
    assert len(ls["small"]) == N and len(ls["big"]) == n_common + n_group + 1 and len(ls["chunks"]) == 3 * n_common + n_group + 3
    f = rng.standard_normal(ls["n_inst"]) * 3 + 50
    bad = rng.choice(ls["n_inst"], 40, replace=False)
    f[bad[:20]] = np.nan
    f[bad[20:30]] = np.inf
    f[bad[30:]] = -np.inf
    f[ls["start"][5]], f[ls["start"][4711]] = np.nan, np.inf       # two persons whose BASE instance failed
    part = k3sim.reduce(f, ls, P)
    assert np.array_equal(part, k3sim.reduce(f, ls, P), equal_nan=True)                 # deterministic, bit for bit
    start, distinct = ls["start"], ls["distinct"]
    f0 = f[start]
    want = np.zeros(5 * P + 2)
    want[0], want[1] = f0.sum(), (~np.isfinite(f0)).sum()
    with np.errstate(invalid="ignore"):
        for p in range(N):
            f0s = f0[p] if np.isfinite(f0[p]) else 0.0
            for j, t in enumerate(distinct[p]):
                fm, fp = f[start[p] + 1 + 2 * j], f[start[p] + 2 + 2 * j]
                want[2 + t] += fp - f0s
                want[2 + P + t] += fm - f0s
                want[2 + 2 * P + t] += 0.0 if np.isfinite(fp) else 1.0
                want[2 + 3 * P + t] += 0.0 if np.isfinite(fm) else 1.0
                want[2 + 4 * P + t] += 0.0 if np.isfinite(f0[p]) else 1.0
    cnt = slice(2 + 2 * P, 2 + 5 * P)
    assert np.array_equal(part[cnt], want[cnt]) and part[1] == want[1]                  # counts exactly
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(part))                                       # a non-finite f± makes that D± non-finite
    assert np.allclose(part[fin], want[fin], rtol=1e-12, atol=1e-9)                     # sums to rounding (different tree)
    assert not np.isfinite(part[0]) and part[1] >= 2                                    # Σ f0 carries the failed base instances, counted
    assert part[2 + 4 * P + 0] == part[1]                                               # nbt of a common parameter = all failed persons


def test_builder_kernels_k2_on_the_cpu():
    """psy_k2.cuh (K2) compiled for the host and run in psy_capi.cu's launch order: the device-built instance list, pair lists,
    segment offsets and chunk / small / big lists equal the plain restatement (k3sim.lists) entry for entry — common, group and
    person-specific parameters, a repeated label, more persons than one layout chunk (1024) and than one K3 chunk (4096); the
    selection predicate of restricted sweeps; the per-set reduction of psy_fit_many; the RK4 step counts against the host rule."""
    from hostsim import k2sim, k3sim
    from psydiff_b200 import _lib
    rng = np.random.default_rng(3)
    for N, F in ((5000, 6), (37, 3), (1, 1), (2100, 40)):
        n_common, n_group = min(3, F), 5
        P = n_common + n_group + N
        tidx = np.empty((N, F), dtype=np.int32)
        tidx[:, :n_common] = np.arange(n_common)
        for k in range(n_common, F):
            tidx[:, k] = n_common + rng.integers(0, n_group, N)
        if F > 4:
            tidx[:, 4] = tidx[:, 3]                      # the same label in two slots: one pair per person
            tidx[:, F - 1] = n_common + n_group + np.arange(N)
        if F == 40:
            tidx[:, 5:20] = (n_common + n_group + np.arange(N))[:, None]   # 61-slot blocks vs 19-slot ones: aligned and packed persons mix
            tidx[::3, 5:20] = tidx[::3, 3:4]
        a, b = k2sim.build(tidx, P), k3sim.lists(tidx, P)
        for key in ("n_inst", "start", "pair_minus", "pair_base", "seg_off", "chunks", "tco", "small", "big"):
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), (N, F, key)
        inst = a["inst"]
        for p in (0, N // 2, N - 1):
            o = int(b["start"][p])
            assert tuple(inst[o]) == (p, -1) and tuple(a["base"][p]) == (p, -1)
            for j, t in enumerate(b["distinct"][p]):
                assert tuple(inst[o + 1 + 2 * j]) == (p, 2 * t) and tuple(inst[o + 2 + 2 * j]) == (p, 2 * t + 1)
        real = inst[:, 0] >= 0
        assert real.sum() == N + 2 * sum(len(d) for d in b["distinct"]) and np.all(inst[~real] == -1)       # the rest is padding
        # restricted sweeps: a one-sided first level (base + that side), then an eps[] fallback level for two (theta, side) pairs
        sel = k2sim.select(inst, None, 1, 1, 0)
        assert np.array_equal(sel, np.flatnonzero(real & ((inst[:, 1] < 0) | (inst[:, 1] % 2 == 0))))
        need = np.zeros(2 * P, dtype=np.uint8)
        need[[1, 2 * (n_common + 1)]] = 1
        sel = k2sim.select(inst, need, 0, 1, 1)
        assert np.array_equal(sel, np.flatnonzero(real & (inst[:, 1] >= 0) & np.isin(inst[:, 1], [1, 2 * (n_common + 1)])))
    # psy_fit_many's reduction: per parameter set, Σ of the non-NaN -2LL and the NaN count
    N, K = 9001, 3
    f = rng.standard_normal(K * N) + 40
    f[[5, N + 7, N + 4100]] = np.nan
    f[2 * N + 11] = np.inf
    tot, bad = k2sim.sets(f, N, K)
    fk = f.reshape(K, N)
    assert np.array_equal(bad, [1, 2, 0]) and np.isinf(tot[2])
    assert np.allclose(tot[:2], [np.nansum(fk[0]), np.nansum(fk[1])], rtol=1e-13)
    # rule T1 on the device = the host's psy_rk4_step_count, on intervals around multiples of h and odd values
    L = _lib.lib()
    h = 1.0 / 30
    dt = np.concatenate([np.arange(0, 40) * h, np.arange(1, 40) * h * (1 + 2e-16), np.arange(1, 40) * h * (1 - 2e-16),
                         rng.random(500) * 3, [0.0, -1.0, np.nan, 1e-300, h / 2, 1e4]])
    dev = k2sim.step_counts(dt, h)
    assert np.array_equal(dev, [L.psy_rk4_step_count(float(d), h) for d in dt])
