import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
import ctypes as C
t=time.time()
kw = synth.config_c4(N=20000, T=100)
print("synth s", time.time()-t, flush=True)
m = pd.newPsydiff(**kw)
t=time.time(); pd.compileModel(m); print("compile+upload s", time.time()-t, flush=True)
L=_lib.lib(); h=m["_handle"].ptr
pk=C.c_double(); _lib.check(L.psy_measure_fp64_peak(C.byref(pk))); print("fp64 peak TF", pk.value)
r=C.c_int(); l=C.c_int(); mt=C.c_int(); L.psy_kernel_info(h, C.byref(r), C.byref(l), C.byref(mt)); print("regs", r.value, "local", l.value)
t=time.time(); f=pd.fitModel(m); print("fit s", time.time()-t, "kernel ms", L.psy_last_kernel_ms(h), "sum", f["m2LL"].sum(), "nonfinite", (~np.isfinite(f["m2LL"])).sum(), flush=True)
for i in range(2):
    t=time.time(); g=pd.getGradients(m); dt_=time.time()-t
    ms=L.psy_last_kernel_ms(h); ni=L.psy_last_kernel_instances(h)
    print("grad s", dt_, "kernel ms", ms, "instances", ni, "person-timesteps/s", 20000*100/dt_, flush=True)
print(list(g.items())[:6])
