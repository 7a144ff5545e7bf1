"""Writes tests/golden/*.json.  The reference (R + Rcpp + Armadillo + Boost) cannot be run in this image and ships no
golden vectors, so these fixtures are ORACLE outputs (oracle/psy_oracle.cpp) on the seeded synthetic configs,
stored next to the independent exact-Kalman-filter value where one exists.  They pin the oracle against regressions and
give the GPU tests a fixture that does not need the oracle at run time.   Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import psydiff_b200 as pd  # noqa: E402
from psydiff_b200 import synth  # noqa: E402
from oracle import exact_kf  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    kw = synth.config_c1()
    m = pd.newPsydiff(**kw)
    o = Oracle(m)
    fit, g = o.fit(), o.gradients()
    A = np.array([[-.3, 0], [.2, -.5]]); G = np.diag([np.sqrt(.5)] * 2); R = np.diag([np.sqrt(.5)] * 2)
    obs = kw["dataset"]["observations"].reshape(20, 50, 2); dts = kw["dataset"]["dt"].reshape(20, 50)
    ex = [exact_kf.kalman_m2ll(A, np.zeros(2), G, np.eye(2), np.zeros(2), R, np.array([1., 2.]), np.eye(2), obs[p], dts[p]) for p in range(20)]
    json.dump({"config": "C1: 2x2 OU, N=20, T=50, rk4 h=1/30, eps=1e-6 central, seed 20210129",
               "labels": m["pars"]["parameterTable"].theta_labels, "m2LL": fit["m2LL"].tolist(),
               "grad_clean": g["grad_clean"].tolist(), "grad_ref": g["grad_ref"].tolist(), "exact_kf_m2LL": ex,
               "latentScores_first_person": fit["latentScores"][:50].tolist()},
              open(os.path.join(HERE, "c1_oracle.json"), "w"), indent=1)
    m4 = pd.newPsydiff(**synth.config_c4(N=16, T=12))
    e4 = Oracle(m4, extended=True)
    json.dump({"config": "C4: 3 coupled Van der Pol, N=16, T=12, rk4 h=min(dt)/30, eps=1e-6 central, seed 20210132; 80-bit oracle build",
               "labels": m4["pars"]["parameterTable"].theta_labels, "m2LL": e4.fit()["m2LL"].tolist(),
               "grad_clean": e4.gradients(nthreads=8)["grad_clean"].tolist()},
              open(os.path.join(HERE, "c4_oracle_ext.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
