#!/usr/bin/env python
"""Config 5 of BASELINE.json end to end: GIST lasso fit (R/GIST.R) of the 8-latent coupled-oscillator model with the 12
coupling gains penalised, followed by a BFGS polish of the unpenalised likelihood (R/optimWrapper.R:23-31), persons
sharded over all GPUs of the box.

    python examples/c5_gist_fit.py --persons-per-gpu 2048                               # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        examples/c5_gist_fit.py --persons-per-gpu 125000                                # C5: N = 1M over 8 x B200
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--persons-per-gpu", type=int, default=2048)
    ap.add_argument("--timepoints", type=int, default=100)
    ap.add_argument("--lambda-per-person", dest="lam", type=float, default=200.0,
                    help="lasso weight per person (lambda = this x N).  -2LL is a SUM over persons, and under the UKF approximation the "
                         "score of an absent gain at 0 has a systematic part of about -150 per person (pseudo-true value about 0.01, "
                         "measured with the CPU oracle), so a lambda that zeroes the absent gains grows with N")
    ap.add_argument("--stepsize-per-person", type=float, default=1e4,
                    help="GIST's initialStepsize = this x N (the curvature of the summed -2LL grows with N; from the package default 1 the "
                         "x1.5 search needs about 55 candidates per outer iteration at N = 1M)")
    ap.add_argument("--outer", type=int, default=15)
    ap.add_argument("--bfgs-iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8, help="line-search candidates evaluated per launch")
    ap.add_argument("--numpy-data", action="store_true", help="simulate the data with numpy on the host (minutes at 125 000 persons) instead of torch on the GPU")
    ap.add_argument("--progress", default="", help="rank 0 appends one JSON line per gradient evaluation (= per GIST outer iteration) here")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import psydiff_b200 as pd
    from psydiff_b200 import dist as pdist, optim, synth
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    npg = a.persons_per_gpu
    t_setup = time.time()
    model = pd.newPsydiff(**synth.config_c5(N=npg, T=a.timepoints, seed=20210133 + rank, id_offset=rank * npg,
                                            simulate_on=None if a.numpy_data else f"cuda:{local}"))
    t_data = time.time() - t_setup
    if world > 1:
        pdist.align_parameters(model)
    pd.compileModel(model)
    start = pd.getParameterValues(model)
    gains = [k for k in start if k.startswith("k") and k[1:].isdigit()]
    t_start = time.time()
    calls = {"grad": 0, "fit": 0}

    def progress(kind, seconds, **extra):
        if rank == 0 and a.progress:
            with open(a.progress, "a") as f:
                f.write(json.dumps(dict(kind=kind, seconds=round(seconds, 3), since_start=round(time.time() - t_start, 3), **extra)) + "\n")

    last = {}

    def gradients_logged(m, **kw):
        """One sharded FD-gradient evaluation; the most recent one is kept, so that BFGS starting where GIST stopped does not
        repeat GIST's last sweep (39 s at 125 000 persons per GPU)."""
        t = time.time()
        cur = pd.getParameterValues(m)
        key = (tuple(cur.items()), tuple(sorted(kw.items())))
        if last.get("key") == key:
            return dict(last["g"])
        g = pdist.getGradientsSharded(m, **kw)
        last.update(key=key, g=dict(g))
        calls["grad"] += 1
        progress("gradient", time.time() - t, call=calls["grad"], zeroed=[k for k in gains if cur[k] == 0.0])
        return g

    def fits_logged(m, cands, **kw):
        t = time.time()
        out = pdist.fitTotalsSharded(m, cands, **kw)
        calls["fit"] += len(cands)
        progress("line_search_batch", time.time() - t, candidates=len(cands), totals=[float(x) for x in np.asarray(out)])
        return out

    progress("setup", time.time() - t_setup, persons=npg * world, gpus=world, parameters=len(start), data_seconds=round(t_data, 2))
    t0 = time.time()
    res = optim.GIST(model, start, lambda_=a.lam * npg * world, adaptiveLassoWeights={k: 1.0 for k in start},
                     regularizedParameters=gains, maxIter_out=a.outer, sig=.4, silent=True,
                     initialStepsize=a.stepsize_per_person * npg * world, stepsizeMax=1e18,
                     fitTotal=pdist.fitTotalSharded, gradients=gradients_logged,
                     batch=a.batch, fitTotalsMany=fits_logged)
    t_gist = time.time() - t0
    est = pd.getParameterValues(model)
    names = [k for k in est if not (k in gains and est[k] == 0.0)]      # post-lasso refit: the gains GIST zeroed stay at 0
    progress("gist_done", t_gist, regM2LL=[float(x) for x in res["regM2LL"][np.isfinite(res["regM2LL"])]],
             gains={k: est[k] for k in gains})
    t0 = time.time()

    def bfgs_gradient(p):               # psydiffGradientsInternal (R/optimWrapper.R:167-190) through the logged, memoised sweep
        pd.setParameterValues(model["pars"]["parameterTable"], {k: float(v) for k, v in zip(names, p)})
        g = gradients_logged(model)
        gv = np.array([g[k] for k in names], dtype=float)
        gv[np.isnan(gv)] = 1.0
        return gv

    def bfgs_fit(p):
        t = time.time()
        v = pdist.psydiffFitInternalSharded(p, names, model)
        calls["fit"] += 1
        progress("bfgs_fit", time.time() - t, m2LL=float(v))
        return v
    out = optim.vmmin(np.array([est[k] for k in names]), bfgs_fit, bfgs_gradient, maxit=a.bfgs_iters)
    t_bfgs = time.time() - t0
    if rank == 0:
        r = res["regM2LL"][np.isfinite(res["regM2LL"])]
        print(json.dumps({
            "persons": npg * world, "gpus": world, "lambda": a.lam * npg * world, "gist_outer_iterations": int(len(r) - 1), "gist_seconds": t_gist,
            "seconds_per_outer_iteration": t_gist / max(len(r) - 1, 1), "regM2LL_first": float(r[0]), "regM2LL_last": float(r[-1]),
            "zeroed_gains": [k for k in gains if est[k] == 0.0], "kept_gains": {k: round(est[k], 4) for k in gains if est[k] != 0.0},
            "bfgs_iterations": a.bfgs_iters, "bfgs_seconds": t_bfgs, "m2LL_after_bfgs": out["value"],
            "gradient_evaluations": calls["grad"], "line_search_fits": calls["fit"],
            "gains_after_bfgs": {k: round(float(v), 4) for k, v in zip(names, out["par"]) if k in gains},
            "generating_gains": "ring neighbours k12 k23 k34 k41 = 0.08, the other eight 0"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
