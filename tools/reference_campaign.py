"""One-off verification campaign (CPU only, needs /root/reference): seeded random model structures (tests/test_kernel_fuzz.py's
generator) through the reference's own C++ (oracle/_ref), the oracle and the host-compiled kernel source.

    python tools/reference_campaign.py fit 40        # -2LL / states / NaN pattern, 40 structures
    python tools/reference_campaign.py grad 16       # FD gradients against the 80-bit oracle, 16 structures

Results of the run recorded in DESIGN.md §5 (seeds 20000.. / 30000..): fit — reference vs oracle <= 6.3e-13, reference vs kernel
<= 1.0e-11, states <= 3.7e-11, identical NaN patterns; grad — kernel inside the 1e-6 + FD-noise gate on all 16 (max |err| 1.7e-7),
the reference's own gradient outside it on one (3 entries).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import psydiff_b200 as pd  # noqa: E402
from oracle import reference_build as rb  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from hostsim.hostsim import HostSim  # noqa: E402
from test_kernel_fuzz import random_model  # noqa: E402


def fit(n):
    worst = dict(ref_oracle=0.0, ref_kernel=0.0, states=0.0)
    for s in range(n):
        kw, _ = random_model(20000 + s, stepper=("rk4", "runge_kutta_dopri5")[s % 2], srUpdate=(s % 4) < 3)
        m = pd.newPsydiff(**dict(kw, breakEarly=False))
        R, O, K = rb.Reference(m).fit(), Oracle(m).fit(), HostSim(m).fit()
        fin = np.isfinite(R["m2LL"])
        assert np.array_equal(fin, np.isfinite(O["m2LL"])) and np.array_equal(fin, np.isfinite(K["m2LL"])), (s, "NaN pattern")
        d = np.maximum(np.abs(R["m2LL"][fin]), 1e-300)
        worst["ref_oracle"] = max(worst["ref_oracle"], float(np.max(np.abs(R["m2LL"] - O["m2LL"])[fin] / d)))
        worst["ref_kernel"] = max(worst["ref_kernel"], float(np.max(np.abs(R["m2LL"] - K["m2LL"])[fin] / d)))
        worst["states"] = max(worst["states"], float(np.nanmax(np.abs(R["latentScores"] - K["latentScores"]))))
    print("fit campaign, worst over", n, "structures:", worst)


def grad(n):
    for s in range(n):
        kw, _ = random_model(30000 + s, stepper="rk4", srUpdate=(s % 3 != 0))
        m = pd.newPsydiff(**dict(kw, breakEarly=False))
        labels = m["pars"]["parameterTable"].theta_labels
        g = rb.Reference(m).gradients()
        k = HostSim(m).gradients(eps=1e-6)
        e80 = Oracle(m, extended=True).gradients()["grad_clean"]
        gv = np.array([g[l] for l in labels])
        atol = 64 * np.finfo(float).eps * abs(float(np.nansum(k["m2LL"]))) / 2e-6
        gate = 1e-6 * np.abs(e80) + atol
        print(f"{s:3d} P={len(gv):3d} max|ref-80bit|={np.abs(gv - e80).max():.2e} max|kernel-80bit|={np.abs(k['grad'] - e80).max():.2e} "
              f"outside gate: kernel {int((np.abs(k['grad'] - e80) > gate).sum())}, reference {int((np.abs(gv - e80) > gate).sum())}", flush=True)


if __name__ == "__main__":
    {"fit": fit, "grad": grad}[sys.argv[1]](int(sys.argv[2]) if len(sys.argv) > 2 else 8)
