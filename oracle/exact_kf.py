"""Independent anchor for the oracle: the exact discrete-time Kalman filter of a LINEAR continuous-time model
(matrix-exponential / Van Loan discretisation).  TEST INFRASTRUCTURE ONLY.

For linear drift and linear measurement the SR-UKF with RK4 step h -> 0 converges to this filter, so the
oracle's -2LL must agree with it to O(h^4).  Nothing here shares code with the oracle or with the CUDA path.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm


def van_loan(A, Q, dt):
    n = A.shape[0]
    M = np.zeros((2 * n, 2 * n))
    M[:n, :n] = -A
    M[:n, n:] = Q
    M[n:, n:] = A.T
    E = expm(M * dt)
    Ad = E[n:, n:].T
    Qd = Ad @ E[:n, n:]
    return Ad, (Qd + Qd.T) / 2


def kalman_m2ll(A, b, G, Lam, d, Rchol, m0, P0, obs, dts):
    """-2 log-likelihood of one person.  dx = (A x + b) dt + G dW ; y = d + Lam x + e, e ~ N(0, Rchol Rchol').
    obs: T x m with NaN = missing; dts: T (first usually 0)."""
    n = A.shape[0]
    Q = G @ G.T
    R = Rchol @ Rchol.T
    m, P = m0.astype(float).copy(), P0.astype(float).copy()
    out = 0.0
    cache = {}
    for t in range(obs.shape[0]):
        dt = dts[t]
        if dt != 0:
            if dt not in cache:
                Ad, Qd = van_loan(A, Q, dt)
                # integral of expm(A s) b ds over [0, dt] via the augmented matrix
                Ab = np.zeros((n + 1, n + 1))
                Ab[:n, :n] = A
                Ab[:n, n] = b
                bd = expm(Ab * dt)[:n, n]
                cache[dt] = (Ad, Qd, bd)
            Ad, Qd, bd = cache[dt]
            m = Ad @ m + bd
            P = Ad @ P @ Ad.T + Qd
        y = obs[t]
        ok = np.isfinite(y)
        if ok.any():
            H = Lam[ok]
            S = H @ P @ H.T + R[np.ix_(ok, ok)]
            r = y[ok] - (d[ok] + H @ m)
            K = P @ H.T @ np.linalg.inv(S)
            out += ok.sum() * np.log(2 * np.pi) + np.linalg.slogdet(S)[1] + r @ np.linalg.solve(S, r)
            m = m + K @ r
            P = P - K @ S @ K.T
            P = (P + P.T) / 2
    return out
