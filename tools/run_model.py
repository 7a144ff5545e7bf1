"""One model on the current GPU: gradient sweeps (-2LL + full central FD gradient) timed by the kernel's own CUDA events.

    python tools/run_model.py CONFIG N [--flags "..."] [--reps 3] [--T 100] [--fit]

CONFIG = c1 | c2 | c3 | c4 | c5 (the n = 8 model of config 5) | k<K> (K coupled Van der Pol oscillators, n = 2K).
Used for compile-flag / carveout experiments and as the command under ncu (one launch per rep).  With `--compile-only`
the cubin is built here (no GPU needed) so it travels with the gpurun snapshot.
"""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib


def config(name, N, T):
    if name == "c1":
        return synth.config_c1(N=N, T=T)
    if name == "c2":
        return synth.config_c2(N=N, T=T)
    if name == "c3":
        return synth.config_c3(N=N, T=T)
    if name == "c4":
        return synth.config_c4(N=N, T=T)
    if name == "c5":
        return synth.config_c5(N=N, T=T)
    if name.startswith("k"):
        return synth.config_coupled(K=int(name[1:]), N=N, T=T)
    raise SystemExit(f"unknown config {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("N", type=int)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--flags", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--fit", action="store_true", help="time fitModel (base instances only) instead of the gradient sweep")
    ap.add_argument("--compile-only", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("--meta", default="", help="write {config, persons, T, instances, kernel hash} as JSON (tools/calibrate.py reads it)")
    ap.add_argument("--devices", type=int, default=1, help="spread the model over this many GPUs inside the library")
    a = ap.parse_args()
    m = pd.newPsydiff(**config(a.config, 8 if a.compile_only else a.N, 4 if a.compile_only else a.T))
    if a.compile_only:
        t = time.time()
        p = pd.compileModelSource(m, a.flags)
        log = p.replace(".cubin", ".log")
        info = ""
        if os.path.exists(log):
            for ln in open(log):
                if "registers" in ln or "stack frame" in ln:
                    info += " " + ln.strip()
        print(f"{a.config} {a.flags!r}: {os.path.basename(p)} ({time.time() - t:.1f}s){info}")
        return
    L = _lib.lib()
    pd.compileModel(m, extra_flags=a.flags)
    h = m["_handle"].ptr
    if a.devices > 1:
        print("devices:", pd.setDevices(m, a.devices))
    r, l, mt = C.c_int(), C.c_int(), C.c_int()
    L.psy_kernel_info(h, C.byref(r), C.byref(l), C.byref(mt))
    ms = []
    res = None
    for _ in range(a.reps):
        if a.fit:
            res = pd.fitModel(m, states=False)["m2LL"]
        else:
            res = np.array(list(pd.getGradients(m).values()))
        ms.append(L.psy_last_kernel_ms(h))
    chk = float(np.nansum(res)) if a.fit else float(np.linalg.norm(res))
    if a.meta:
        import json
        json.dump({"config": a.config.upper(), "persons": a.N, "T": a.T, "instances": int(L.psy_n_instances(h)),   # real instances (no padding slots)
                   "kernel_hash": os.path.basename(m["_handle"].cubin).replace(".cubin", ""), "registers": r.value, "local_bytes": l.value,
                   "block": mt.value, "flags": a.flags, "kernel_ms": min(ms)}, open(a.meta, "w"))
    print(f"{a.tag or a.config} N={a.N} flags={a.flags!r} carveout={os.environ.get('PSY_SMEM_CARVEOUT', '-')} regs={r.value} "
          f"local={l.value} block={mt.value} instances={L.psy_last_kernel_instances(h)} kernel_ms min={min(ms):.2f} "
          f"all={[round(x, 2) for x in ms]} check={chk:.12g} pt/s={a.N * a.T / (min(ms) * 1e-3):.4g}", flush=True)


if __name__ == "__main__":
    main()
