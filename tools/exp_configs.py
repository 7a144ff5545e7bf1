"""Throughput of every BASELINE.json config's model on one GPU (gradient sweep = -2LL + full central FD gradient)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
L = _lib.lib()
cases = [("C1", synth.config_c1, dict(N=20, T=50)), ("C2", synth.config_c2, dict(N=100_000, T=100)),
         ("C3", synth.config_c3, dict(N=250_000, T=100)), ("C4", synth.config_c4, dict(N=16_384, T=100)),
         ("C5-model", synth.config_c5, dict(N=4096, T=100))]
for name, fn, kw in cases:
    m = pd.newPsydiff(**fn(**kw))
    pd.compileModel(m)
    h = m["_handle"].ptr
    pd.getGradients(m)
    t = time.time(); g = pd.getGradients(m); dt_ = time.time() - t
    r, l, mt = C.c_int(), C.c_int(), C.c_int(); L.psy_kernel_info(h, C.byref(r), C.byref(l), C.byref(mt))
    N, T = kw["N"], kw["T"]
    print(f"{name:9s} n={m['nlatent']} N={N} T={T} P={len(g)} instances={L.psy_n_instances(h)} regs={r.value} local={l.value} "
          f"kernel_ms={L.psy_last_kernel_ms(h):.2f} wall_ms={dt_*1e3:.2f} person-timesteps/s={N*T/dt_:.4g}", flush=True)
