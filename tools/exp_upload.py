"""GPU experiment: where the end-to-end step (uploadData + one gradient evaluation) spends its host time."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib, model as M
L = _lib.lib()
for N in [int(a) for a in sys.argv[1:]] or [8192]:
    m = pd.newPsydiff(**synth.config_c4(N=N, T=100, simulate_on="cuda:0"))
    m["data"]["observations"] = np.asfortranarray(m["data"]["observations"])
    pd.compileModel(m)
    h = m["_handle"].ptr
    pd.getGradients(m)
    for rep in range(6):
        t0 = time.perf_counter()
        ro, pu = M._person_blocks(m["data"]["person"])
        t1 = time.perf_counter()
        pd.uploadData(m)
        t2 = time.perf_counter()
        g = pd.getGradients(m)
        t3 = time.perf_counter()
        print(f"N={N} rep {rep}: person_blocks {1e3*(t1-t0):7.1f} ms  uploadData {1e3*(t2-t1):7.1f} ms  getGradients {1e3*(t3-t2):8.1f} ms  kernel {L.psy_last_kernel_ms(h):8.1f} ms", flush=True)
