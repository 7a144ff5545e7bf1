"""Flop accounting for the filter kernel (FMA = 2 flops; a reciprocal, sqrt, exp or log counts 1).

``kernel_flops``  — the arithmetic the kernel's reduced (m, A) formulation executes per RK4 step / per update, counted by
                    walking the same loop nests as psy_filter.cuh (checked against ncu's DADD/DMUL/DFMA instruction
                    counters, profiles/).  This is the figure bench.py's roofline uses: it is what the FP64 pipe must do.
``survey_flops``  — SURVEY.md §8(d)'s closed form for the reference's full-sigma-matrix formulation, reported beside it.
"""
from __future__ import annotations


def rhs_flops(n: int, Ff: int, l_diag: bool = True, dep=None) -> int:
    """One evaluation of Filter::rhs_scaled (column-scaled form: unit-diagonal FMA chains).  ``dep`` = the Jacobian
    sparsity masks of codegen.drift_dependency (None = dense): unmoved (component, column) pairs cost nothing."""
    if dep is None:
        dep = [(1 << n) - 1] * n
    moved = [[(dep[i] >> j) != 0 for j in range(n)] for i in range(n)]   # moved[i][j]: column j can change component i
    tn = n * (n + 1) // 2
    off = n * (n - 1) // 2
    f = n                        # reciprocals of the diagonal
    f += off                     # A~ = A D^-1
    for j in range(n):           # W = A~^-1 (unit lower): seeded FMA chains
        for i in range(j + 1, n):
            f += 2 * (i - j - 1)
    f += (2 * n + 1) * Ff        # drift at every sigma point
    f += 4 * tn                  # sigma points m ± sqrt(c) A_j (two FMAs per element)
    f += 4 * sum(moved[i][j] for i in range(n) for j in range(n))              # d = fp - fm; dm += wm1 (fp + fm)
    f += 2 * sum(moved[k][j] for j in range(n) for i in range(n) for k in range(i))   # G~[:, j] = W d_j
    f += n                       # dm seed: (wm0 + 2 wm1 #unmoved) f0
    f += n                       # kap * A_jj
    if l_diag:
        f += off                 # Bq = W diag(q)
        for i in range(n):
            for j in range(i + 1):
                f += 2 * j
    else:
        for i in range(n):       # Z = W Q, then lower(Z W')
            for j in range(n):
                f += 2 * i
        for i in range(n):
            for j in range(i + 1):
                f += 2 * j
    f += 3 * n + 4 * off         # fold kap (G~ D + D G~') into Phi, halve the diagonal
    for i in range(n):           # dA = (A~ Phi) D^-1
        for j in range(i + 1):
            f += 2 * (i - j)     # the D^-1 column scaling is folded into the RK4 coefficients (step_flops)
    return f


def step_flops(n: int, Ff: int, l_diag: bool = True, dep=None) -> int:
    ns = n + n * (n + 1) // 2
    return 4 * rhs_flops(n, Ff, l_diag, dep) + 4 * (4 * ns + 2 * n)


def update_flops(n: int, m: int, Fh: int) -> int:
    s = 2 * n + 1
    tn = n * (n + 1) // 2
    f = s * Fh + 3 * tn                      # measurement images of the sigma points
    f += m * (2 * n + 3)                     # mu
    rows = s + m
    per_row = 2 * m                          # deviation rows
    for k in range(m):
        per_row += 3 + 2 + 2 + 6 * (m - 1 - k)   # sqrt, recip, c, s, rotate the rest of the row
    f += rows * per_row
    for k in range(m):                       # column-0 up/down-date
        f += 8 + 5 * (m - 1 - k)
    f += m                                   # reciprocals of diag(srS)
    for a in range(m):                       # w = srS^-1 r, dist, log
        f += 1 + 2 * a + 1 + 2 + 1
    for i in range(n):                       # C row and Z row
        f += m * (3 * (i + 1) + 1)
        for a in range(m):
            f += 2 * a + 1
    f += 2 * n * m                           # m += Z w
    for _ in range(m):                       # m rank-1 downdates of A
        for k in range(n):
            f += 8 + 5 * (n - 1 - k)
    return f + 4


def kernel_flops(n, m, Ff, Fh, total_rk4_steps, timepoints, l_diag=True, dep=None):
    """Executed flops of one filter instance that takes ``total_rk4_steps`` RK4 steps over ``timepoints`` time points."""
    return total_rk4_steps * step_flops(n, Ff, l_diag, dep) + timepoints * update_flops(n, m, Fh)


def survey_flops(n, m, Ff, Fh, total_rk4_steps, timepoints):
    """F_rhs = s F_f + 7ns + 2n^2 s + 6n^2 + 3n^3, F_step = 4 F_rhs + 14ns, plus F_upd per time point."""
    s = 2 * n + 1
    o = m
    rhs = s * Ff + 7 * n * s + 2 * n * n * s + 6 * n * n + 3 * n ** 3
    step = 4 * rhs + 14 * n * s
    upd = (s * Fh + 2 * m * s + 2 * (s + o) * o * o + 4 * o * o + (2 * n * o * s + 2 * n * s) + o ** 3 // 3 + 2 * n * o * o
           + n * o * o + 3 * n * n * o + o * o + 2 * n * n)
    return total_rk4_steps * step + timepoints * upd


def rk4_step_count(t1: float, h: float) -> int:
    """Steps odeint's integrate_const takes over [0, t1] with step h (rule T1): the remainder is dropped."""
    eps = 2.220446049250313e-16
    if not (h > 0) or t1 != t1:
        return 0
    step, time = 0, 0.0
    while (time + h) - t1 <= eps:
        step += 1
        time = 0.0 + step * h
    return step
