"""Cross-check of a GPU bench line on the CPU: Σ-2LL of the benchmark shard that `python bench.py --persons-per-gpu N` filtered on
the B200 (its `check.m2ll_total`) against the SAME kernel source compiled for the host (tests/hostsim), all N persons x T = 100.

    python tools/bench_total_crosscheck.py profiles/r01f_bench_8192persons_1gpu.json

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


def work(args):
    lo, hi, N, T = args
    from hostsim.hostsim import HostSim
    kw = synth.config_c4(N=N, T=T, seed=20210132, id_offset=0)          # bench.py's rank-0 shard
    ds = kw["dataset"]
    mn = np.min(ds["dt"][ds["dt"] > 0])
    kw["timeStep"] = np.linspace(mn / 30, mn / 10, 5)                    # the step of the WHOLE shard (R/prepareModel.R:435-438)
    kw["dataset"] = {k: v[lo * T:hi * T] for k, v in ds.items()}
    return lo, HostSim(pd.newPsydiff(**kw)).fit(states=False)["m2LL"]


def main():
    line = json.load(open(sys.argv[1]))
    N, T = int(line["config"]["persons_per_gpu"]), int(line["config"]["timepoints"])
    gpu = float(line["check"]["m2ll_total"])
    t0 = time.time()
    step = max(64, N // 16)
    with Pool(min(8, os.cpu_count() or 1)) as p:
        res = sorted(p.map(work, [(i, min(i + step, N), N, T) for i in range(0, N, step)]))
    tot = float(np.sum(np.concatenate([r[1] for r in res])))
    print(f"persons {N} x T {T}: GPU {gpu!r}  host-compiled kernel source {tot!r}  relative difference {abs(tot - gpu) / abs(gpu):.2e}  ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
