"""A second, independent restatement of the reference's SR-UKF path in numpy — TEST INFRASTRUCTURE ONLY.

Purpose: cross-check oracle/psy_oracle.cpp with an implementation that shares no code with it and that gets its dense
linear algebra from a real LAPACK (numpy.linalg.qr / inv / cholesky), the way the reference gets it from Armadillo.
In particular it confirms SURVEY §8c T3 (the sign convention of LAPACK's QR does not leak through cholupdate).
Written function by function after inst/include/psydiffInternals.h and R/modelTemplate.R; pure Python loops, so only
for tiny cases.  Drift / measurement equations are passed as Python callables f(x, pars, person, t).
"""
from __future__ import annotations

import numpy as np

DBL_EPS = np.finfo(float).eps


def getSigmaPoints(m, A, c):                       # psydiffInternals.h:4-16 (covarianceIsRoot = TRUE)
    n = len(m)
    X = np.zeros((n, 2 * n + 1))
    X[:, 1:n + 1] = np.sqrt(c) * A
    X[:, n + 1:] = -1 * np.sqrt(c) * A
    return X + m[:, None]


def getMeanWeights(n, alpha, beta, kappa):         # :18-24
    lam = alpha ** 2 * (n + kappa) - n
    w = np.full(2 * n + 1, 1 / (2 * (n + lam)))
    w[0] = lam / (n + lam)
    return w


def getCovWeights(n, alpha, beta, kappa):          # :26-33
    lam = alpha ** 2 * (n + kappa) - n
    w = np.full(2 * n + 1, 1 / (2 * (n + lam)))
    w[0] = lam / (n + lam) + (1 - alpha ** 2 + beta)
    return w


def getWMatrix(wm, wc):                            # :35-44
    D = np.eye(len(wm)) - wm[:, None]
    return D @ np.diag(wc) @ D.T


def getAFromSigmaPoints(X, wm, c):                 # :51-60
    n = X.shape[0]
    return (X[:, 1:n + 1] - (X @ wm)[:, None]) / np.sqrt(c)


def getPhi(M):                                     # :69-75
    P = np.tril(M)
    P[np.diag_indices_from(P)] = .5 * np.diag(M)
    return P


def cholupdate(L, x, v, direction):                # :133-165
    L, x = L.copy(), x.copy()
    d = {"update": 1, "downdate": -1}[direction]
    v = abs(v) ** .25
    x = v * x
    n = len(x)
    for k in range(n):
        r = np.sqrt(L[k, k] ** 2 + d * x[k] ** 2)
        c, s = r / L[k, k], x[k] / L[k, k]
        L[k, k] = r
        if k < n - 1:
            L[k + 1:, k] = (L[k + 1:, k] + d * s * x[k + 1:]) / c
            x[k + 1:] = c * x[k + 1:] - s * L[k + 1:, k]
    return L


def predictSquareRootOfS(Y, mu, Rchol, wc0, wc1):  # :174-198, qr_C :167-171 through LAPACK
    T = np.hstack([np.sqrt(wc1) * (Y - mu[:, None]), Rchol])
    R = np.linalg.qr(T.T, mode="r")
    srS = R.T.copy()
    srS = cholupdate(srS, Y[:, 0] - mu, abs(wc0), "downdate" if wc0 < 0 else "update")
    return np.tril(srS)


def m2ll_chol(nobs, raw, mean, srS):               # :117-131
    ci = np.linalg.inv(np.tril(srS))
    r = raw - mean
    return nobs * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(srS))) + float(r @ ci.T @ ci @ r)


def rhs(X, t, drift, pars, person, wm, wc, c, L):  # odeintModel::operator() + getMMatrix, R/modelTemplate.R:85-163
    n, s = X.shape
    A = getAFromSigmaPoints(X, wm, c)
    Ainv = np.linalg.inv(np.tril(A))
    F = np.column_stack([drift(X[:, i], pars, person, t) for i in range(s)])
    mx, mf = X @ wm, F @ wm
    s1 = sum(wc[i] * np.outer(X[:, i] - mx, F[:, i] - mf) for i in range(s))
    s2 = sum(wc[i] * np.outer(F[:, i] - mf, X[:, i] - mx) for i in range(s))
    M = Ainv @ (s1 + s2 + L @ np.eye(n) @ L.T) @ Ainv.T
    APhi = A @ getPhi(M)
    aug = np.zeros_like(X)
    aug[:, 1:n + 1] = APhi
    aug[:, n + 1:] = -1 * APhi
    return mf[:, None] + np.sqrt(c) * aug


def integrate_const_rk4(X, t1, h, f):              # Boost.odeint integrate_const<runge_kutta4>, rule T1
    step, time = 0, 0.0
    while (time + h) - t1 <= DBL_EPS:
        k1 = f(X, time)
        k2 = f(X + h / 2 * k1, time + h / 2)
        k3 = f(X + h / 2 * k2, time + h / 2)
        k4 = f(X + h * k3, time + h)
        X = X + h / 6 * k1 + h / 3 * k2 + h / 3 * k3 + h / 6 * k4
        step += 1
        time = 0.0 + step * h
    return X


def fit_person(obs, dts, drift, meas, pars, person, m0, A0, Lmat, Rchol, alpha=.2, beta=2.0, kappa=0.0, h=1 / 30, stepper="rk4"):
    """Body of fitModel's person loop, SR variant (R/modelTemplate.R:244-411).  Returns (-2LL, latentScores, predictedManifest)."""
    n = len(m0)
    if n + kappa == 0:
        kappa -= 1
    c = alpha ** 2 * (n + kappa)
    wm, wc = getMeanWeights(n, alpha, beta, kappa), getCovWeights(n, alpha, beta, kappa)
    W = getWMatrix(wm, wc)
    m, A = m0.astype(float).copy(), A0.astype(float).copy()
    m2ll, tsum = 0.0, 0.0
    lat, pred = [], []
    for tp in range(len(dts)):
        tsum += dts[tp]
        ob = obs[tp]
        nm = np.flatnonzero(np.isfinite(ob))
        m_ = m.copy()
        X = getSigmaPoints(m_, A, c)
        if abs(dts[tp]) > 0:
            g = lambda Z, t: rhs(Z, t, drift, pars, person, wm, wc, c, Lmat)
            X = integrate_const_rk4(X, dts[tp], h, g) if stepper == "rk4" else integrate_dopri5(X, dts[tp], h, g)
            m_ = X[:, 0].copy()
            A = np.tril(getAFromSigmaPoints(X, wm, c))
        Y = np.column_stack([meas(X[:, i], pars, person, tsum) for i in range(X.shape[1])])
        mu = Y @ wm
        pred.append(mu)
        if len(nm) > 0:
            srS = predictSquareRootOfS(Y[nm], mu[nm], Rchol[np.ix_(nm, nm)], wc[0], wc[1])
            C = X @ W @ Y[nm].T
            si = np.linalg.inv(np.tril(srS))
            K = C @ si.T @ si
            m = m_ + K @ (ob - mu)[nm]
            U = K @ srS
            for co in range(U.shape[1]):
                A = np.tril(cholupdate(A, U[:, co], 1, "downdate"))
            m2ll += m2ll_chol(len(nm), ob[nm], mu[nm], srS)
        else:
            m = m_
        lat.append(m.copy())
    return m2ll, np.array(lat), np.array(pred)


# ---- Boost.odeint integrate() = integrate_adaptive(controlled_runge_kutta<runge_kutta_dopri5>, 1e-6, 1e-6) — rule T2 --------
_DP_C = [0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_A = [[], [1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9], [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
         [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656], [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
_DP_B5 = np.array([35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0])
_DP_B4 = np.array([5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40])


def integrate_dopri5(X, t1, dt, f, eps_abs=1e-6, eps_rel=1e-6):
    """Adaptive Dormand–Prince 5(4) with odeint's default controller; state X may be any array (max-norm over all entries)."""
    t = 0.0
    k1 = f(X, t)
    while t1 - t > DBL_EPS:
        if (t + dt) - t1 > DBL_EPS:
            dt = t1 - t
        for _ in range(500):
            ks = [k1]
            for s in range(1, 7):
                Xs = X + dt * sum(a * k for a, k in zip(_DP_A[s], ks))
                ks.append(f(Xs, t + _DP_C[s] * dt))
            Xn = X + dt * sum(b * k for b, k in zip(_DP_B5[:6], ks[:6]))     # = the stage-7 point (FSAL)
            xerr = dt * sum((b5 - b4) * k for b5, b4, k in zip(_DP_B5, _DP_B4, ks))
            err = np.max(np.abs(xerr) / (eps_abs + eps_rel * (np.abs(X) + abs(dt) * np.abs(k1))))
            if err > 1.0:
                dt *= max(0.9 * err ** (-1 / 3), 0.2)
                continue
            t += dt
            if err < 0.5:
                dt *= 0.9 * max(5.0 ** -5, err) ** (-1 / 5)
            X, k1 = Xn, ks[6]
            break
        else:
            raise RuntimeError("step size underflow")
    return X


def fit_person_cov(obs, dts, drift, meas, pars, person, m0, A0, Lmat, Rchol, alpha=.2, beta=2.0, kappa=0.0, h=1 / 30):
    """Covariance variant of fitModel's person loop (srUpdate = FALSE, R/modelTemplate.R:660-838): chol(P) per time point,
    S = Y W Y' + R symmetrised, K = C inv(S), P = P_ - K S K' symmetrised, det-based -2LL (psydiffInternals.h:103-115)."""
    n = len(m0)
    if n + kappa == 0:
        kappa -= 1
    c = alpha ** 2 * (n + kappa)
    wm, wc = getMeanWeights(n, alpha, beta, kappa), getCovWeights(n, alpha, beta, kappa)
    W = getWMatrix(wm, wc)
    m = m0.astype(float).copy()
    P = A0 @ A0.T
    R = Rchol @ Rchol.T
    m2ll, tsum, lat = 0.0, 0.0, []
    for tp in range(len(dts)):
        tsum += dts[tp]
        ob = obs[tp]
        nm = np.flatnonzero(np.isfinite(ob))
        m_, P_ = m.copy(), P.copy()
        A = np.linalg.cholesky(P_)
        X = getSigmaPoints(m_, A, c)
        if abs(dts[tp]) > 0:
            X = integrate_const_rk4(X, dts[tp], h, lambda Z, t: rhs(Z, t, drift, pars, person, wm, wc, c, Lmat))
            m_ = X[:, 0].copy()
            Af = getAFromSigmaPoints(X, wm, c)
            P_ = Af @ Af.T
        Y = np.column_stack([meas(X[:, i], pars, person, tsum) for i in range(X.shape[1])])
        mu = Y @ wm
        if len(nm) > 0:
            S = Y[nm] @ W @ Y[nm].T + R[np.ix_(nm, nm)]
            S = (S + S.T) / 2
            C = X @ W @ Y[nm].T
            K = C @ np.linalg.inv(S)
            r = (ob - mu)[nm]
            m = m_ + K @ r
            P = P_ - K @ S @ K.T
            P = (P + P.T) / 2
            m2ll += len(nm) * np.log(2 * np.pi) + np.log(np.linalg.det(S)) + float(r @ np.linalg.inv(S) @ r)
        else:
            m, P = m_, (P_ + P_.T) / 2
        lat.append(m.copy())
    return m2ll, np.array(lat)
