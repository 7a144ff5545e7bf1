import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib, flops
L = _lib.lib()
for K, N in ((4, 2048), (1, 65536)):
    kw = synth.config_coupled(K=K, N=N, T=100) if K > 1 else synth.config_c2(N=N, T=100)
    m = pd.newPsydiff(**kw)
    pd.compileModel(m)
    h = m["_handle"].ptr
    g = pd.getGradients(m); t = time.time(); g = pd.getGradients(m); dt_ = time.time() - t
    ms = L.psy_last_kernel_ms(h); ni = L.psy_n_instances(h)
    print(f"K={K} n={m['nlatent']} N={N} instances={ni} kernel_ms={ms:.1f} wall_s={dt_:.3f} person-timesteps/s={N*100/dt_:.0f}", flush=True)
