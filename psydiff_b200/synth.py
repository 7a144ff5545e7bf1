"""Deterministic synthetic workloads for the configs of BASELINE.json: model definition + simulated data.

Each ``config_cX`` returns the keyword arguments of ``newPsydiff`` (same argument names as R/prepareModel.R:399-406).
Data are simulated from the model itself (there is no network for datasets); generators are seeded numpy PCG64.
"""
from __future__ import annotations

import numpy as np


def _ou_discretise(A, Q, dt):
    """Exact discretisation of dx = A x dt + G dW, G G' = Q (Van Loan): returns (Ad, Qd)."""
    from scipy.linalg import expm
    n = A.shape[0]
    Mx = np.zeros((2 * n, 2 * n))
    Mx[:n, :n] = -A
    Mx[:n, n:] = Q
    Mx[n:, n:] = A.T
    E = expm(Mx * dt)
    Ad = E[n:, n:].T
    Qd = Ad @ E[:n, n:]
    return Ad, (Qd + Qd.T) / 2


def config_c1(N=20, T=50, seed=20210129, missing=0.0, **settings):
    """C1/C2: the bivariate linear-drift (OU) example of R/prepareModel.R:36-118: dx = DRIFT*x; y = MANIFESTMEANS + LAMBDA*x;
    11 free parameters (4 drift, 2 manifest means, 3 T0 log-Cholesky, 2 ln-diffusion)."""
    rng = np.random.default_rng(seed)
    DRIFT = np.array([[-.3, 0.0], [.2, -.5]])
    Lc = np.diag([np.sqrt(.5)] * 2)
    Rc = np.diag([np.sqrt(.5)] * 2)
    m0 = np.array([1.0, 2.0])
    Ad, Qd = _ou_discretise(DRIFT, Lc @ Lc.T, 1.0)
    Lq = np.linalg.cholesky(Qd)
    x = m0[None, :] + rng.standard_normal((N, 2))  # A0 = I
    obs = np.empty((N, T, 2))
    for t in range(T):
        if t > 0:
            x = x @ Ad.T + rng.standard_normal((N, 2)) @ Lq.T
        obs[:, t, :] = x + rng.standard_normal((N, 2)) @ Rc.T
    if missing > 0:
        obs[rng.random(obs.shape) < missing] = np.nan
    dt = np.tile(np.concatenate([[0.0], np.ones(T - 1)]), N)
    dataset = {"person": np.repeat(np.arange(1, N + 1), T), "observations": obs.reshape(N * T, 2), "dt": dt}
    parameters = {
        "LAMBDA": np.array([["1", "0"], ["0", "1"]], dtype=object),
        "DRIFT": np.array([["drift_eta1 = -0.3", "drift_eta1_eta2 = 0.0"],
                           ["drift_eta2_eta1 = 0.2", "drift_eta2 = -0.5"]], dtype=object),
        "MANIFESTMEANS": np.array([["mm_Y1 = 0.0"], ["mm_Y2 = 0.0"]], dtype=object),
    }
    kw = dict(
        dataset=dataset,
        latentEquations="dx = DRIFT*x;",
        manifestEquations="y = MANIFESTMEANS + LAMBDA*x;",
        L=np.array([[f"eta1_eta1 = {float(np.sqrt(.5))!r}", "0"], ["0", f"eta2_eta2 = {float(np.sqrt(.5))!r}"]], dtype=object),
        Rchol=Rc,
        A0=np.array([["T0var_eta1 = 1", "0"], ["T0var_eta2_eta1 = 0", "T0var_eta2 = 1"]], dtype=object),
        m0=np.array([["1"], ["2"]], dtype=object),
        parameters=parameters,
        timeStep=np.array([1.0 / 30]), integrateFunction="rk4", eps=np.array([1e-6]), direction="central",
    )
    kw.update(settings)
    return kw


def config_c2(N=100_000, T=100, seed=20210130, **settings):
    """C2: the C1 model at N=100k, T=100, 5 % of the observations missing at random, fixed-step RK4 h=1/30."""
    return config_c1(N=N, T=T, seed=seed, missing=0.05, **settings)


_C4_LATENT = """
dx(0) = x(1);
dx(1) = mu1*(1.0 - x(0)*x(0))*x(1) - om1*om1*x(0) + k12*(x(2) - x(0)) + k13*(x(4) - x(0));
dx(2) = x(3);
dx(3) = mu2*(1.0 - x(2)*x(2))*x(3) - om2*om2*x(2) + k21*(x(0) - x(2)) + k23*(x(4) - x(2));
dx(4) = x(5);
dx(5) = mu3*(1.0 - x(4)*x(4))*x(5) - om3*om3*x(4) + k31*(x(0) - x(4)) + k32*(x(2) - x(4));
"""


def _vdp3_drift(x, mu, om, k):
    q, v = x[:, 0::2], x[:, 1::2]
    dx = np.empty_like(x)
    dx[:, 0::2] = v
    coup = np.empty_like(q)
    coup[:, 0] = k[0] * (q[:, 1] - q[:, 0]) + k[1] * (q[:, 2] - q[:, 0])
    coup[:, 1] = k[2] * (q[:, 0] - q[:, 1]) + k[3] * (q[:, 2] - q[:, 1])
    coup[:, 2] = k[4] * (q[:, 0] - q[:, 2]) + k[5] * (q[:, 1] - q[:, 2])
    dx[:, 1::2] = mu * (1.0 - q * q) * v - (om * om)[None, :] * q + coup
    return dx


def config_c4(N=1_000_000, T=100, seed=20210132, n_groups=8, id_offset=0, simulate_on=None, **settings):
    """C4: three diffusively coupled Van der Pol oscillators (n = m = 6, y = x), irregular time intervals,
    mu1..mu3 specific to 8 groups (P = 3*8 + 27 = 51 parameters, 30 per person -> 61 filter instances per person),
    fixed-step RK4 with the package default h = min(dt>0)/30."""
    rng = np.random.default_rng(seed)
    ids = id_offset + np.arange(1, N + 1)          # person ids of this (shard of the) dataset
    group = ((ids - 1) % n_groups) + 1
    mu_g = np.linspace(0.6, 1.4, n_groups)[:, None] * np.array([1.0, 0.9, 1.1])[None, :]   # [group, oscillator]
    om = np.array([1.0, 1.2, 0.8])
    kk = np.array([0.10, 0.05, 0.08, 0.12, 0.06, 0.09])
    ldiff, rsd = 0.1, 0.3
    m0 = np.array([1.0, 0.0, -0.5, 0.5, 0.3, -0.8])
    dts = rng.choice([0.5, 0.75, 1.0, 1.25, 1.5], size=(N, T)) + rng.random((N, T)) * 0.1
    dts[:, 0] = 0.0
    obs = np.empty((N, T, 6))
    CH = 65536
    sub = 0.05
    if simulate_on is not None:                    # same scheme with torch ops on the named device (bench.py --simulate-on-gpu)
        gain = {(0, 1): kk[0], (0, 2): kk[1], (1, 0): kk[2], (1, 2): kk[3], (2, 0): kk[4], (2, 1): kk[5]}
        obs[:] = _simulate_coupled_torch(simulate_on, seed, m0, mu_g[group - 1], om, gain, dts, ldiff, rsd, sub)
    for c0 in range(0, N if simulate_on is None else 0, CH):
        c1 = min(N, c0 + CH)
        nb = c1 - c0
        mu = mu_g[group[c0:c1] - 1]
        x = m0[None, :] + 0.5 * rng.standard_normal((nb, 6))
        for t in range(T):
            rem = dts[c0:c1, t].copy()
            nsub = int(np.ceil(rem.max() / sub)) if t > 0 else 0
            for _ in range(nsub):
                h = np.minimum(rem, sub)
                rem -= h
                k1 = _vdp3_drift(x, mu, om, kk)
                k2 = _vdp3_drift(x + h[:, None] * k1, mu, om, kk)
                x = x + 0.5 * h[:, None] * (k1 + k2) + ldiff * np.sqrt(h)[:, None] * rng.standard_normal((nb, 6))
            obs[c0:c1, t, :] = x + rsd * rng.standard_normal((nb, 6))
    dataset = {"person": np.repeat(ids, T), "observations": obs.reshape(N * T, 6), "dt": dts.reshape(-1)}
    parameters = {"mu1": 1.0, "mu2": 0.9, "mu3": 1.1, "om1": 1.0, "om2": 1.2, "om3": 0.8,
                  "k12": 0.10, "k13": 0.05, "k21": 0.08, "k23": 0.12, "k31": 0.06, "k32": 0.09}
    Lm = np.full((6, 6), "0", dtype=object)
    Rm = np.full((6, 6), "0", dtype=object)
    for i in range(6):
        Lm[i, i] = f"l{i + 1} = {ldiff}"
        Rm[i, i] = f"r{i + 1} = {rsd}"
    kw = dict(
        dataset=dataset, latentEquations=_C4_LATENT, manifestEquations="y = x;",
        L=Lm, Rchol=Rm, A0=np.eye(6) * 0.5,
        m0=np.array([[f"m0_{i + 1} = {m0[i]}"] for i in range(6)], dtype=object),
        parameters=parameters,
        grouping="mu1 | grp; mu2 | grp; mu3 | grp;", groupingvariables={"grp": (((np.arange(1, id_offset + N + 1) - 1) % n_groups) + 1).astype(float)},  # indexed by id
        integrateFunction="rk4", eps=np.array([1e-6]), direction="central",
    )
    kw.update(settings)
    return kw


def config_c3(N=250_000, T=100, seed=20210131, id_offset=0, **settings):
    """C3: Van der Pol oscillator dx0 = x1; dx1 = mu_p (1 − x0²) x1 − om² x0 with a person-specific mu ("mu | person;",
    mu_p ~ U(0.5, 1.5)), common om, L = diag(.1), Rchol = diag(.3), dt = 0.25, adaptive Dormand–Prince
    (integrateFunction = "runge_kutta_dopri5", initial step dt/30).  P = N + 7 parameters, 8 per person -> 17 instances."""
    rng = np.random.default_rng(seed)
    ids = id_offset + np.arange(1, N + 1)
    mu = rng.uniform(0.5, 1.5, size=N)
    om, ldiff, rsd, dt = 1.0, 0.1, 0.3, 0.25
    x = np.array([1.0, 0.0])[None, :] + 0.3 * rng.standard_normal((N, 2))
    obs = np.empty((N, T, 2))
    nsub, h = 5, dt / 5

    def f(x):
        return np.stack([x[:, 1], mu * (1.0 - x[:, 0] ** 2) * x[:, 1] - om * om * x[:, 0]], axis=1)
    for t in range(T):
        if t > 0:
            for _ in range(nsub):
                k1 = f(x)
                k2 = f(x + h * k1)
                x = x + 0.5 * h * (k1 + k2) + ldiff * np.sqrt(h) * rng.standard_normal((N, 2))
        obs[:, t, :] = x + rsd * rng.standard_normal((N, 2))
    dts = np.tile(np.concatenate([[0.0], np.full(T - 1, dt)]), N)
    dataset = {"person": np.repeat(ids, T), "observations": obs.reshape(N * T, 2), "dt": dts}
    kw = dict(
        dataset=dataset,
        latentEquations="dx(0) = x(1);\ndx(1) = mu*(1.0 - x(0)*x(0))*x(1) - om*om*x(0);",
        manifestEquations="y(0) = x(0);\ny(1) = x(1);",
        L=np.array([["l1 = 0.1", "0"], ["0", "l2 = 0.1"]], dtype=object),
        Rchol=np.array([["r1 = 0.3", "0"], ["0", "r2 = 0.3"]], dtype=object),
        A0=np.eye(2) * 0.3, m0=np.array([["m0_1 = 1.0"], ["m0_2 = 0.0"]], dtype=object),
        parameters={"mu": 1.0, "om": 1.0}, grouping="mu | person;",
        timeStep=np.array([dt / 30]), integrateFunction="runge_kutta_dopri5", eps=np.array([1e-6]), direction="central",
    )
    kw.update(settings)
    return kw


_ONE_STREAM = 65536          # up to here a dataset is simulated from one random stream (bit-stable small cases)
_BLOCK = 8192


def _sim_threads():
    """Threads for the data simulation of one rank: the host's cores divided by the ranks of this node."""
    import os
    return max(1, min(8, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))


def _simulate_coupled_torch(device, seed, m0, mu, om, gain, dts, ldiff, rsd, sub):
    """Heun / Euler-Maruyama simulation of config_coupled's data with torch ops on `device` (synthetic data only: the random
    stream is torch's, so the values differ from the numpy path's; minutes of numpy become a second on the GPU)."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    N, T = dts.shape
    n = m0.shape[0]
    K = n // 2
    f64 = dict(dtype=torch.float64, device=dev)
    mu_t, om2, dts_t = torch.as_tensor(mu, **f64), torch.as_tensor(om * om, **f64)[None, :], torch.as_tensor(dts, **f64)
    G = torch.zeros(K, K, **f64)                   # acc_i += sum_j g_ij (q_j - q_i)  ==  q @ G^T - rowsum(G) * q
    for (i, j), g in gain.items():
        G[i, j] = g
    Gs = G.sum(1)[None, :]

    def drift(x):
        q, v = x[:, 0::2], x[:, 1::2]
        dx = torch.empty_like(x)
        dx[:, 0::2] = v
        dx[:, 1::2] = mu_t * (1.0 - q * q) * v - om2 * q + q @ G.T - Gs * q
        return dx
    randn = lambda: torch.randn(N, n, generator=gen, **f64)
    x = torch.as_tensor(m0, **f64)[None, :] + 0.5 * randn()
    obs = torch.empty(N, T, n, **f64)
    for t in range(T):
        rem = dts_t[:, t].clone()
        for _ in range(int(np.ceil(float(rem.max()) / sub)) if t > 0 else 0):
            h = torch.clamp(rem, max=sub)
            rem -= h
            k1 = drift(x)
            k2 = drift(x + h[:, None] * k1)
            x = x + 0.5 * h[:, None] * (k1 + k2) + ldiff * torch.sqrt(h)[:, None] * randn()
        obs[:, t, :] = x + rsd * randn()
    return obs.cpu().numpy()


def config_coupled(K=4, N=1000, T=100, seed=20210133, n_groups=8, id_offset=0, simulate_on=None, **settings):
    """C5 family: K diffusively coupled Van der Pol oscillators on a ring with all-to-all coupling gains
    (n = m = 2K; K = 4 is the 8-latent model of config 5, its 12 coupling gains are the lasso targets of the GIST fit).
    mu_i group-specific as in C4, irregular dt, rk4 with the package default step.  `simulate_on` = a torch device string
    simulates the observations there instead of in numpy (examples/c5_gist_fit.py at 125 000 persons per GPU)."""
    rng = np.random.default_rng(seed)
    n = 2 * K
    ids = id_offset + np.arange(1, N + 1)
    group = ((ids - 1) % n_groups) + 1
    mu_g = np.linspace(0.6, 1.4, n_groups)[:, None] * np.linspace(0.9, 1.1, K)[None, :]
    om = np.linspace(0.8, 1.2, K)
    pairs = [(i, j) for i in range(K) for j in range(K) if i != j]
    gain = {pr: (0.08 if (pr[1] - pr[0]) % K == 1 else 0.0) for pr in pairs}    # only the ring neighbours are truly coupled
    ldiff, rsd = 0.1, 0.3
    m0 = np.array([(1.0 if i % 2 == 0 else 0.0) * (1 - 0.3 * (i // 2)) for i in range(n)])
    dts = rng.choice([0.5, 0.75, 1.0, 1.25, 1.5], size=(N, T)) + rng.random((N, T)) * 0.1
    dts[:, 0] = 0.0

    def drift(x, mu):
        q, v = x[:, 0::2], x[:, 1::2]
        dx = np.empty_like(x)
        dx[:, 0::2] = v
        acc = mu * (1.0 - q * q) * v - (om * om)[None, :] * q
        for (i, j), g in gain.items():
            if g:
                acc[:, i] += g * (q[:, j] - q[:, i])
        dx[:, 1::2] = acc
        return dx
    obs = np.empty((N, T, n))
    sub = 0.05

    def simulate(c0, c1, rng):
        x = m0[None, :] + 0.5 * rng.standard_normal((c1 - c0, n))
        mu = mu_g[group[c0:c1] - 1]
        for t in range(T):
            rem = dts[c0:c1, t].copy()
            for _ in range(int(np.ceil(rem.max() / sub)) if t > 0 else 0):
                h = np.minimum(rem, sub)
                rem -= h
                k1 = drift(x, mu)
                k2 = drift(x + h[:, None] * k1, mu)
                x = x + 0.5 * h[:, None] * (k1 + k2) + ldiff * np.sqrt(h)[:, None] * rng.standard_normal((c1 - c0, n))
            obs[c0:c1, t, :] = x + rsd * rng.standard_normal((c1 - c0, n))
    if simulate_on is not None:                    # C5 at its stated size: the same scheme in torch on the named device
        obs[:] = _simulate_coupled_torch(simulate_on, seed, m0, mu_g[group - 1], om, gain, dts, ldiff, rsd, sub)
    elif N <= _ONE_STREAM:
        simulate(0, N, rng)                        # one random stream: what every fixture and the bench's C5 shard were made with
    else:                                          # C5 at its stated size: cache-sized person blocks, one stream each, on a few threads
        from concurrent.futures import ThreadPoolExecutor
        blocks = [(c0, min(N, c0 + _BLOCK)) for c0 in range(0, N, _BLOCK)]
        with ThreadPoolExecutor(max_workers=_sim_threads()) as pool:
            list(pool.map(lambda b: simulate(b[0], b[1], np.random.default_rng([seed, b[0]])), blocks))
    lines = []
    parameters = {}
    for i in range(K):
        parameters[f"mu{i + 1}"] = float(np.linspace(0.9, 1.1, K)[i])
        parameters[f"om{i + 1}"] = float(om[i])
    for (i, j) in pairs:
        parameters[f"k{i + 1}{j + 1}"] = 0.05
    for i in range(K):
        lines.append(f"dx({2 * i}) = x({2 * i + 1});")
        coup = " + ".join(f"k{i + 1}{j + 1}*(x({2 * j}) - x({2 * i}))" for j in range(K) if j != i)
        lines.append(f"dx({2 * i + 1}) = mu{i + 1}*(1.0 - x({2 * i})*x({2 * i}))*x({2 * i + 1}) - om{i + 1}*om{i + 1}*x({2 * i}) + {coup};")
    Lm = np.full((n, n), "0", dtype=object)
    Rm = np.full((n, n), "0", dtype=object)
    for i in range(n):
        Lm[i, i] = f"l{i + 1} = {ldiff}"
        Rm[i, i] = f"r{i + 1} = {rsd}"
    kw = dict(
        dataset={"person": np.repeat(ids, T), "observations": obs.reshape(N * T, n), "dt": dts.reshape(-1)},
        latentEquations="\n".join(lines), manifestEquations="y = x;", L=Lm, Rchol=Rm, A0=np.eye(n) * 0.5,
        m0=np.array([[f"m0_{i + 1} = {m0[i]}"] for i in range(n)], dtype=object), parameters=parameters,
        grouping=" ".join(f"mu{i + 1} | grp;" for i in range(K)),
        groupingvariables={"grp": (((np.arange(1, id_offset + N + 1) - 1) % n_groups) + 1).astype(float)},
        integrateFunction="rk4", eps=np.array([1e-6]), direction="central",
    )
    kw.update(settings)
    return kw


def config_c5(N=1_000_000, T=100, **settings):
    """C5: four coupled oscillators (n = m = 8); the 12 coupling gains are lasso-penalised in the GIST fit."""
    return config_coupled(K=4, N=N, T=T, **settings)
