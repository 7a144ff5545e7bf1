"""Cross-check of a GPU bench line on the CPU: Σ-2LL of the benchmark shard that `python bench.py --persons-per-gpu N` filtered on
the B200 (its `check.m2ll_total`), all N persons x T = 100, against
  * the SAME kernel source compiled for the host (tests/hostsim) — default: arithmetic of the device build vs the host build, or
  * the ORACLE (`--oracle`: oracle/psy_oracle.cpp, the CPU restatement of the reference's algorithm) — parity at benchmark scale.

    python tools/bench_total_crosscheck.py profiles/r01f_bench_8192persons_1gpu.json [--oracle] [--config C2|C3|C5]

`--gradient` (C4 headline shard, with --oracle): the FULL central-FD gradient of the shard from the oracle — 2P_i + 1 = 61 filter sweeps per
person on the host, 46 min for 8192 persons on 8 cores in double, 2 h 45 min with `--extended` (the oracle's 80-bit build) — against the
line's `check.grad_l2` and `check.grad_head`.
`--config` checks one of the side legs of the line (`configs.<name>.check.m2ll_total`: C2 100 000 persons, C3 250 000 persons under
the adaptive stepper, C5's model 16 384 persons) instead of the C4 headline shard.

Recorded result (r01f, 8192 persons x 100 time points, about 62 RK4 steps per interval): GPU 4773326.782180222, host-compiled kernel
source 4773326.78218022 — relative difference 3.9e-16 (30 s on 8 host cores); r01e, the default 125 000-person shard:
GPU 72781608.99735554, host 72781608.99735555 — 2.1e-16 (9 min).
"""
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import psydiff_b200 as pd  # noqa: E402
from psydiff_b200 import synth  # noqa: E402


_KW = None          # the shard's newPsydiff arguments, made once in the parent (the workers are forked from it)
_ORACLE = False
_GRADIENT = False
_EXTENDED = False


def shard_kwargs(name, N, T):
    """newPsydiff arguments of bench.py's rank-0 shard of workload `name` (bench.make_model)."""
    if name == "C4":
        return synth.config_c4(N=N, T=T, seed=20210132, id_offset=0)
    if name == "C2":
        return synth.config_c2(N=N, T=T, seed=20210130)
    if name == "C3":
        return synth.config_c3(N=N, T=T, seed=20210131, id_offset=0)
    if name == "C5":
        return synth.config_c5(N=N, T=T, seed=20210133, id_offset=0)
    raise SystemExit(f"unknown config {name}")


def work(args):
    lo, hi, N, T = args
    kw = dict(_KW)
    ds = kw["dataset"]
    if "timeStep" not in kw:
        mn = np.min(ds["dt"][ds["dt"] > 0])
        kw["timeStep"] = np.linspace(mn / 30, mn / 10, 5)                # the step of the WHOLE shard (R/prepareModel.R:435-438)
    kw["dataset"] = {k: v[lo * T:hi * T] for k, v in ds.items()}
    model = pd.newPsydiff(**kw)
    if _ORACLE and _GRADIENT:
        from oracle.oracle import Oracle
        g = Oracle(model, extended=_EXTENDED).gradients()
        return lo, g["m2LL"], dict(zip(model["pars"]["parameterTable"].theta_labels, g["grad_clean"]))
    if _ORACLE:
        from oracle.oracle import Oracle
        return lo, Oracle(model).fit()["m2LL"]
    from hostsim.hostsim import HostSim
    return lo, HostSim(model).fit(states=False)["m2LL"]


def main():
    global _KW, _ORACLE, _GRADIENT, _EXTENDED
    _ORACLE = "--oracle" in sys.argv[2:]
    _GRADIENT = "--gradient" in sys.argv[2:]
    _EXTENDED = "--extended" in sys.argv[2:]
    line = json.load(open(sys.argv[1]))
    name = sys.argv[sys.argv.index("--config") + 1].upper() if "--config" in sys.argv else "C4"
    T = int(line["config"]["timepoints"])
    if name == "C4":
        N, gpu = int(line["config"]["persons_per_gpu"]), float(line["check"]["m2ll_total"])
    else:
        N, gpu = int(line["configs"][name]["persons"]), float(line["configs"][name]["check"]["m2ll_total"])
    t0 = time.time()
    _KW = shard_kwargs(name, N, T)                                      # bench.py's rank-0 shard
    if _ORACLE:
        from oracle import oracle as orc
        orc.build(extended=_EXTENDED)
    step = max(64, N // 64)
    with Pool(min(16, os.cpu_count() or 1)) as p:
        res = sorted(p.map(work, [(i, min(i + step, N), N, T) for i in range(0, N, step)]))
    per = np.concatenate([r[1] for r in res])
    tot = float(np.sum(per))
    if _GRADIENT:
        grad = {}
        for r in res:
            for k, v in r[2].items():
                grad[k] = grad.get(k, 0.0) + float(v)
        gl2 = float(np.linalg.norm(list(grad.values())))
        chk = line["check"]
        print(f"{name}: persons {N} x T {T}: FD gradient over the whole shard, oracle{' (80-bit build)' if _EXTENDED else ''} vs GPU")
        print(f"  |g|_2  GPU {chk['grad_l2']!r}  oracle {gl2!r}  relative difference {abs(gl2 - chk['grad_l2']) / gl2:.2e}")
        for k, v in chk.get("grad_head", {}).items():
            print(f"  {k:8s} GPU {v!r}  oracle {grad[k]!r}  relative difference {abs(grad[k] - v) / abs(grad[k]):.2e}")
        print(f"  gate of the parity tests for one entry: 1e-6 |g| + 64 eps |sum f| / (2 * 1e-6) = 1e-6 |g| + {64 * np.finfo(float).eps * abs(tot) / 2e-6:.3e}")
    what = "oracle (CPU restatement of the reference)" if _ORACLE else "host-compiled kernel source"
    print(f"{name}: persons {N} x T {T}: GPU {gpu!r}  {what} {tot!r}  relative difference {abs(tot - gpu) / abs(gpu):.2e}  "
          f"(non-finite persons: {int((~np.isfinite(per)).sum())}; {time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
