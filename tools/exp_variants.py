"""GPU experiment: kernel time of the C4 gradient sweep for several compile variants (launch bounds / block size)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kw = synth.config_c4(N=N, T=100)
L = _lib.lib()
variants = sys.argv[2:] or ["", "-DPSY_MIN_BLOCKS=3", "-DPSY_MIN_BLOCKS=4", "-DPSY_BLOCK=64 -DPSY_MIN_BLOCKS=4", "-DPSY_BLOCK=64 -DPSY_MIN_BLOCKS=6", "-DPSY_BLOCK=64 -DPSY_MIN_BLOCKS=8", "-DPSY_BLOCK=256 -DPSY_MIN_BLOCKS=1"]
ref = None
for v in variants:
    m = pd.newPsydiff(**kw)
    t = time.time(); pd.compileModel(m, extra_flags=v); tc = time.time() - t
    h = m["_handle"].ptr
    r, l, mt = C.c_int(), C.c_int(), C.c_int(); L.psy_kernel_info(h, C.byref(r), C.byref(l), C.byref(mt))
    g = pd.getGradients(m); g = pd.getGradients(m)
    ms = L.psy_last_kernel_ms(h)
    gv = np.array(list(g.values()))
    if ref is None: ref = gv
    print(f"{v!r:45s} regs={r.value:3d} local={l.value:5d} block={mt.value:3d} kernel_ms={ms:9.2f} pts/s={N*100/ms*1e3:10.0f} maxreldiff={np.max(np.abs(gv-ref)/np.abs(ref)):.2e} compile_s={tc:.1f}", flush=True)
