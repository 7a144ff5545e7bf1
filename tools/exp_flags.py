"""GPU experiment: kernel time of compile-flag variants on C4 (8192 persons) and the n = 8 model (2048 persons); min of 3 calls."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
L = _lib.lib()
variants = sys.argv[1:] or ["", "-DPSY_IEEE_DIV"]
for name, kw, N in (("C4", synth.config_c4(N=8192, T=100), 8192), ("n8", synth.config_coupled(K=4, N=2048, T=100), 2048)):
    ref = None
    for v in variants + variants[:1]:                      # the first variant again at the end: drift check
        m = pd.newPsydiff(**kw)
        pd.compileModel(m, extra_flags=v)
        h = m["_handle"].ptr
        ms = []
        for _ in range(3):
            g = pd.getGradients(m)
            ms.append(L.psy_last_kernel_ms(h))
        gv = np.array(list(g.values()))
        ref = gv if ref is None else ref
        print(f"{name} {v!r:22s} kernel_ms min={min(ms):9.2f} all={[round(x, 2) for x in ms]} maxreldiff={np.max(np.abs(gv - ref) / np.abs(ref)):.2e}", flush=True)
