"""The committed outputs of the reference's own C++ (tests/golden/*_reference.json, written by tests/golden/make_reference_golden.py
from the oracle/_ref build) and the model each one belongs to.  Used by the oracle, host-compiled-kernel and GPU parity tests."""
import json
import os

import numpy as np

from psydiff_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "c1": lambda: synth.config_c1(),
    "c1_missing_cov": lambda: synth.config_c1(N=24, T=30, missing=0.2, seed=9, srUpdate=False),
    "c3": lambda: synth.config_c3(N=16, T=12),
    "c4": lambda: synth.config_c4(N=16, T=12),
    "c4_T100": lambda: synth.config_c4(N=8, T=100),
    "n8": lambda: synth.config_coupled(K=4, N=6, T=8, n_groups=2),
}


def load(name):
    d = json.load(open(os.path.join(HERE, "golden", f"{name}_reference.json")))
    for k in ("m2LL", "latentScores_first_person", "predictedManifest_first_person", "getGradients", "grad_80bit"):
        if k in d:
            d[k] = np.array(d[k], dtype=float)
    return d


def check_fit(fit, ref, rtol=1e-9):
    """Per-person -2LL within rtol relative; first person's states within rtol of the state scale."""
    assert np.array_equal(np.isfinite(fit["m2LL"]), np.isfinite(ref["m2LL"]))
    rel = np.abs(fit["m2LL"] - ref["m2LL"]) / np.maximum(np.abs(ref["m2LL"]), 1e-300)
    assert rel.max() < rtol, rel.max()
    rows = len(ref["latentScores_first_person"])
    scale = max(1.0, np.nanmax(np.abs(ref["latentScores_first_person"])))
    assert np.nanmax(np.abs(fit["latentScores"][:rows] - ref["latentScores_first_person"])) < rtol * scale
    assert np.nanmax(np.abs(fit["predictedManifest"][:rows] - ref["predictedManifest_first_person"])) < rtol * scale
    return rel.max()


def check_gradient(g, ref, eps=1e-6, exact=None):
    """FD gradient against the reference's getGradients: 1e-6 relative plus the rounding floor of a double-precision finite
    difference of Σ-2LL (both sides evaluate Σ-2LL to a few ulp and divide by 2 eps).  The fixture's `grad_80bit` (the same finite
    difference by the oracle's 80-bit build) widens the allowance by the reference's OWN distance from the exact value of its
    formulation: on the nonlinear C4 model and under adaptive stepping (C3) the reference's double-precision gradient is
    ~1e-4 relative away from it."""
    gr = ref["getGradients"]
    exact = ref.get("grad_80bit") if exact is None else exact
    atol = 64 * np.finfo(float).eps * abs(float(np.sum(ref["m2LL"]))) / (2 * eps)
    allow = 1e-6 * np.abs(gr) + atol
    if exact is not None:
        allow = allow + 3.0 * np.max(np.abs(gr - np.asarray(exact)))     # noise level of the reference's own double-precision FD
    bad = np.abs(np.asarray(g) - gr) > allow
    assert not bad.any(), [(ref["labels"][i], float(np.asarray(g)[i]), float(gr[i])) for i in np.flatnonzero(bad)]
