"""The filter primitives the reference exports to R one by one (R/RcppExports.R:12-257, src/exposeInlineFunctions.cpp),
with the same names and argument lists, over the C ABI of libpsydiff_b200.so (host utilities of psy_primitives.cpp —
they are not on the fit / gradient path, which is the CUDA kernel).  Matrices are numpy arrays (any layout; converted to
column-major for the call, like arma::mat)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PsydiffError


def _f(a):
    return np.asfortranarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))


def _v(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


_d = _lib.dptr


def getSigmaPoints(m, A, covarianceIsRoot, c):
    """R/RcppExports.R:49 — inst/include/psydiffInternals.h:4-16."""
    m, A = _v(m), _f(A)
    n = len(m)
    X = np.zeros((n, 2 * n + 1), order="F")
    if _lib.lib().psy_getSigmaPoints(n, _d(m), _d(A), 1 if covarianceIsRoot else 0, float(c), _d(X)) != 0:
        raise PsydiffError("chol(): decomposition failed")
    return X


def getMeanWeights(n, alpha, beta, kappa):
    w = np.zeros(2 * int(n) + 1)
    _lib.lib().psy_getMeanWeights(int(n), float(alpha), float(beta), float(kappa), _d(w))
    return w


def getCovWeights(n, alpha, beta, kappa):
    w = np.zeros(2 * int(n) + 1)
    _lib.lib().psy_getCovWeights(int(n), float(alpha), float(beta), float(kappa), _d(w))
    return w


def getWMatrix(meanWeights, covWeights):
    wm, wc = _v(meanWeights), _v(covWeights)
    W = np.zeros((len(wm), len(wm)), order="F")
    _lib.lib().psy_getWMatrix(len(wm), _d(wm), _d(wc), _d(W))
    return W


def computeMeanFromSigmaPoints(sigmaPoints, meanWeights):
    X, wm = _f(sigmaPoints), _v(meanWeights)
    out = np.zeros(X.shape[0])
    _lib.lib().psy_computeMeanFromSigmaPoints(X.shape[0], X.shape[1], _d(X), _d(wm), _d(out))
    return out


def getAFromSigmaPoints(sigmaPoints, meanWeights, c):
    X, wm = _f(sigmaPoints), _v(meanWeights)
    A = np.zeros((X.shape[0], X.shape[0]), order="F")
    _lib.lib().psy_getAFromSigmaPoints(X.shape[0], _d(X), _d(wm), float(c), _d(A))
    return A


def computePFromSigmaPoints(sigmaPoints, meanWeights, c):
    X, wm = _f(sigmaPoints), _v(meanWeights)
    P = np.zeros((X.shape[0], X.shape[0]), order="F")
    _lib.lib().psy_computePFromSigmaPoints(X.shape[0], _d(X), _d(wm), float(c), _d(P))
    return P


def getPhi(squareMatrix):
    M = _f(squareMatrix)
    out = np.zeros_like(M, order="F")
    _lib.lib().psy_getPhi(M.shape[0], _d(M), _d(out))
    return out


def predictMu(Y_, meanWeights):
    Y, wm = _f(Y_), _v(meanWeights)
    mu = np.zeros(Y.shape[0])
    _lib.lib().psy_predictMu(Y.shape[0], Y.shape[1], _d(Y), _d(wm), _d(mu))
    return mu


def predictS(Y_, W, R):
    Y, W, R = _f(Y_), _f(W), _f(R)
    S = np.zeros((Y.shape[0], Y.shape[0]), order="F")
    _lib.lib().psy_predictS(Y.shape[0], Y.shape[1], _d(Y), _d(W), _d(R), _d(S))
    return S


def predictC(sigmaPoints, Y_, W):
    X, Y, W = _f(sigmaPoints), _f(Y_), _f(W)
    Cm = np.zeros((X.shape[0], Y.shape[0]), order="F")
    _lib.lib().psy_predictC(X.shape[0], Y.shape[0], X.shape[1], _d(X), _d(Y), _d(W), _d(Cm))
    return Cm


def computeK(C_, S):
    Cm, S = _f(C_), _f(S)
    K = np.zeros_like(Cm, order="F")
    if _lib.lib().psy_computeK(Cm.shape[0], Cm.shape[1], _d(Cm), _d(S), _d(K)) != 0:
        raise PsydiffError("inv(): matrix is singular")
    return K


def updateM(m_, K, residual):
    m_, K, r = _v(m_), _f(K), _v(residual)
    out = np.zeros(len(m_))
    _lib.lib().psy_updateM(K.shape[0], K.shape[1], _d(m_), _d(K), _d(r), _d(out))
    return out


def updateP(P_, K, S):
    P_, K, S = _f(P_), _f(K), _f(S)
    out = np.zeros_like(P_, order="F")
    _lib.lib().psy_updateP(K.shape[0], K.shape[1], _d(P_), _d(K), _d(S), _d(out))
    return out


def computeIndividualM2LL(nObservedVariables, rawData, expectedMeans, expectedCovariance):
    """R/RcppExports.R:214 — inst/include/psydiffInternals.h:103-115."""
    raw, mean, cov = _v(rawData), _v(expectedMeans), _f(expectedCovariance)
    out = C.c_double()
    _lib.check(_lib.lib().psy_computeIndividualM2LL(int(nObservedVariables), _d(raw), _d(mean), _d(cov), C.byref(out)))
    return out.value


def cholupdate(L, x, v, direction):
    """R/RcppExports.R:227 — inst/include/psydiffInternals.h:133-165 (v enters as v^(1/4); unknown direction is an error)."""
    Lc, x = _f(L).copy(order="F"), _v(x)
    if _lib.lib().psy_cholupdate(Lc.shape[0], _d(Lc), _d(x), float(v), str(direction).encode()) != 0:
        raise PsydiffError("Unknown direction in cholupdate. Possible are downdate and update.")
    return Lc


def qr_(X):
    X = _f(X)
    R = np.zeros((X.shape[1], X.shape[1]), order="F")
    _lib.lib().psy_qr(X.shape[0], X.shape[1], _d(X), _d(R))
    return R


def logChol2Chol(logChol):
    lc = _f(logChol)
    out = np.zeros_like(lc, order="F")
    _lib.lib().psy_logChol2Chol(lc.shape[0], _d(lc), _d(out))
    return out


def setParameterList(parameterTable, parameterList, person):
    """setParameterList (inst/include/psydiffBasic.h:63-121): copy one person's CHANGED rows of the parameter table into
    the parameter list at (target, row, col); mutates ``parameterList``.  (The GPU path never needs this — every filter
    instance gathers its own values through the slot map — it is kept because the reference exports it.)"""
    hits = np.flatnonzero(np.asarray(parameterTable.persons) == person)
    if len(hits) == 0:
        raise PsydiffError(f"Person {person} not found in data set.")
    p = int(hits[0])
    for k, row in enumerate(parameterTable.slot_rows):
        if not parameterTable.changed[p, k]:
            continue
        value = float(parameterTable.theta_values[parameterTable.theta_index[p, k]])
        target = row["target"]
        if target not in parameterList:
            import warnings
            warnings.warn(f"Target {target} not found")
            continue
        if row["targetClass"] == "matrix":
            parameterList[target][row["row"], row["col"]] = value
        elif row["targetClass"] == "double":
            parameterList[target] = value
        else:
            raise PsydiffError("Parameter class not found")
    return parameterList
