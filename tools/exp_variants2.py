"""GPU experiment: kernel time for compile variants on C4 (n=6) and the C5 model (n=8)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
L = _lib.lib()
variants = sys.argv[1:] or ["", "-DPSY_G_SOLVE=1"]
for name, kw, N in (("C4", synth.config_c4(N=8192, T=100), 8192), ("n8", synth.config_coupled(K=4, N=2048, T=100), 2048)):
    ref = None
    for v in variants:
        m = pd.newPsydiff(**kw)
        pd.compileModel(m, extra_flags=v)
        h = m["_handle"].ptr
        r, l, mt = C.c_int(), C.c_int(), C.c_int(); L.psy_kernel_info(h, C.byref(r), C.byref(l), C.byref(mt))
        g = pd.getGradients(m); g = pd.getGradients(m)
        ms = L.psy_last_kernel_ms(h)
        gv = np.array(list(g.values()))
        if ref is None: ref = gv
        print(f"{name} {v!r:40s} regs={r.value:3d} local={l.value:5d} kernel_ms={ms:9.2f} pts/s={N*100/ms*1e3:10.0f} maxreldiff={np.max(np.abs(gv-ref)/np.abs(ref)):.2e}", flush=True)
