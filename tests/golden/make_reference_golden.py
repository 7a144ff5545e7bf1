"""Writes tests/golden/*_reference.json: outputs of the REFERENCE'S OWN C++ for this path, run here through
oracle/reference_build.py (the fitModel / getGradients text of /root/reference/R/modelTemplate.R and the headers of
/root/reference/inst/include, compiled against the stand-in Rcpp / Armadillo / Boost.odeint headers of oracle/refbuild).
The GPU box has no /root/reference, so the GPU tests (and the host-compiled-kernel tests) read these committed fixtures.
Run here:  python tests/golden/make_reference_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import psydiff_b200 as pd  # noqa: E402
from psydiff_b200 import synth  # noqa: E402
from oracle import reference_build as rb  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
HOW = ("reference C++ (R/modelTemplate.R template + inst/include/psydiff*.h) compiled against oracle/refbuild stand-ins for "
       "Rcpp / Armadillo / Boost.odeint; g++ -O2 -ffp-contract=off")


def case(name, config, kw, grad=True, first_rows=None):
    m = pd.newPsydiff(**kw)
    R = rb.Reference(m)
    fit = R.fit()
    rows = first_rows or int(m["_row_off"][1])
    out = {"config": config, "how": HOW, "labels": m["pars"]["parameterTable"].theta_labels, "m2LL": fit["m2LL"].tolist(),
           "latentScores_first_person": fit["latentScores"][:rows].tolist(),
           "predictedManifest_first_person": fit["predictedManifest"][:rows].tolist()}
    if grad:
        g = R.gradients()
        assert list(g) == out["labels"]
        out["getGradients"] = list(g.values())
        # the same finite difference evaluated by the oracle's 80-bit build: how far the reference's own double-precision
        # gradient is from the exact value of its formulation (the allowance the parity checks add to the 1e-6 gate)
        from oracle.oracle import Oracle
        out["grad_80bit"] = Oracle(m, extended=True).gradients(nthreads=8)["grad_clean"].tolist()
    json.dump(out, open(os.path.join(HERE, f"{name}_reference.json"), "w"), indent=1)
    print(name, "sum(-2LL) =", float(np.sum(fit["m2LL"])))


def main():
    case("c1", "C1: 2x2 OU, N=20, T=50, rk4 h=1/30, eps=1e-6 central, seed 20210129 (synth.config_c1())", synth.config_c1())
    case("c1_missing_cov", "C1 model, covariance variant (srUpdate=FALSE), N=24, T=30, 20 % missing, seed 9 "
         "(synth.config_c1(N=24, T=30, missing=0.2, seed=9, srUpdate=False))", synth.config_c1(N=24, T=30, missing=0.2, seed=9, srUpdate=False))
    case("c3", "C3: Van der Pol, mu | person, dopri5, N=16, T=12, eps=1e-6 central (synth.config_c3(N=16, T=12))", synth.config_c3(N=16, T=12))
    case("c4", "C4: 3 coupled Van der Pol (n=m=6), irregular dt, 8 groups, N=16, T=12, rk4 h=min(dt)/30, eps=1e-6 central, seed 20210132 "
         "(synth.config_c4(N=16, T=12)); note: the reference's double-precision FD gradient of this model carries ~1e-4 relative "
         "rounding noise (see c4_oracle_ext.json for the 80-bit value)", synth.config_c4(N=16, T=12))
    case("c4_T100", "C4 at the benchmark's series length: N=8, T=100 (about 62 RK4 steps per interval, 99 intervals), fit only "
         "(synth.config_c4(N=8, T=100))", synth.config_c4(N=8, T=100), grad=False)
    case("n8", "C5's model: 4 coupled Van der Pol (n=m=8), N=6, T=8, 2 groups (synth.config_coupled(K=4, N=6, T=8, n_groups=2))",
         synth.config_coupled(K=4, N=6, T=8, n_groups=2), grad=False)


if __name__ == "__main__":
    main()
