import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
L = _lib.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
t = time.time(); kw = synth.config_c3(N=N, T=100); print("synth s", time.time() - t, flush=True)
m = pd.newPsydiff(**kw)
t = time.time(); pd.compileModel(m); print("compile+upload+slots s", time.time() - t, flush=True)
h = m["_handle"].ptr
for i in range(2):
    t = time.time(); g = pd.getGradients(m); dt_ = time.time() - t
    print(f"C3 N={N} P={len(g)} instances={L.psy_n_instances(h)} kernel_ms={L.psy_last_kernel_ms(h):.1f} wall_s={dt_:.3f} person-timesteps/s={N*100/dt_:.0f}", flush=True)
t = time.time(); f = pd.fitModel(m); print("fit wall s", time.time() - t, "nonfinite", int((~np.isfinite(f['m2LL'])).sum()))
