"""Tiny run of every kernel path for compute-sanitizer (memcheck): SR/cov, rk4/dopri5, padded and unpadded person blocks,
one-sided sweep (sub-list launch), eps fallback, small- and chunked-segment reductions, state outputs."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth
for kw in (synth.config_c1(N=7, T=9, missing=0.2), synth.config_c1(N=5, T=6, srUpdate=False), synth.config_c3(N=9, T=5),
           synth.config_c4(N=3, T=4), synth.config_c1(N=70, T=3, eps=np.array([30.0, 1e-6]), A0=np.eye(2))):
    m = pd.newPsydiff(**kw)
    pd.compileModel(m)
    f = pd.fitModel(m)
    g = pd.getGradients(m)
    m["direction"] = "left"
    g2 = pd.getGradients(m)
    print(m["nlatent"], float(np.nansum(f["m2LL"])), len(g), flush=True)
print("done")
