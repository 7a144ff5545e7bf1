"""psydiff_b200 — B200-native SR-UKF likelihood / finite-difference gradient path of jhorzek/psydiff.

Host-side mirror of the reference's R API (same names) over the C ABI of ``libpsydiff_b200.so``.
"""
from ._lib import PsydiffError  # noqa: F401
from .model import (  # noqa: F401
    newPsydiff, compileModel, compileModelSource, uploadData, fitModel, fitTotals, getGradients, getGradient, getGradientsParallel,
    compileParallelGradients, stopParallelGradients, setDevices, usedDevices, deviceCount, shardInfo, getParameterValues, setParameterValues, clonePsydiffModel,
    sdeModelMatrix, sepNameFromValue, chol2LogChol, extractParameters, makeGroups, checkEquation, fromMxMatrix,
)
from .codegen import modelTemplate, prepareEquations  # noqa: F401,E402
from .optim import (  # noqa: F401,E402
    psydiffFitInternal, psydiffGradientsInternal, psydiffOptimBFGS, psydiffOptimNelderMead, psydiffOptimx, extractOptimx, psydiffDEoptim, psydiffSubplex, psydiffNLopt,
    getStandardErrors, GIST, getRegValue, getSubgradients, bestOfN, getDataTemplate, simulateData, fitf,
)
from .primitives import (  # noqa: F401,E402
    getSigmaPoints, getMeanWeights, getCovWeights, getWMatrix, computeMeanFromSigmaPoints, getAFromSigmaPoints,
    computePFromSigmaPoints, getPhi, predictMu, predictS, predictC, computeK, updateM, updateP, computeIndividualM2LL,
    cholupdate, qr_, logChol2Chol, setParameterList,
)
