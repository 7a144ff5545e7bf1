"""The N > 1 path on CPU: two gloo ranks each hold a contiguous person shard, build their partial-sum vector
[Σf0, #nonfinite, D+[P], D-[P], nf+[P], nf-[P], nbt[P]], all-reduce it and run the library's host-side sweep bookkeeping
(psy_sweep_*).  The result must equal the unsharded gradient.  Per-instance -2LL values come from the CPU oracle here
(no GPU in this test); on the GPU box the same protocol runs in psydiff_b200.dist.getGradientsSharded with NCCL."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import psydiff_b200 as pd
from psydiff_b200 import _lib, dist as pdist, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_kw(kw, lo, hi, T):
    ds = kw["dataset"]
    sel = slice(lo * T, hi * T)
    out = dict(kw)
    out["dataset"] = {"person": ds["person"][sel], "observations": ds["observations"][sel], "dt": ds["dt"][sel]}
    return out


def _worker(rank, world, port, N, T, q):
    import torch.distributed as dist
    import torch
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = synth.config_c1(N=N, T=T, missing=0.1, seed=5)
    kw["grouping"] = "mm_Y1 | g;"
    kw["groupingvariables"] = {"g": np.array([2, 2, 2, 1, 1, 3, 3, 3][:N], dtype=float)}
    lo, hi = pdist.shard_bounds(N, world)[rank]
    model = pd.newPsydiff(**_shard_kw(kw, lo, hi, T))
    pdist.align_parameters(model)                    # gloo all_gather_object of the shards' label lists
    pt = model["pars"]["parameterTable"]
    P, eps = len(pt.theta_labels), float(model["eps"][0])
    orc = Oracle(model)
    theta = np.ascontiguousarray(pt.theta_values)
    f0 = orc.fit(theta, states=False)["m2LL"]
    part = np.zeros(5 * P + 2)
    part[0], part[1] = f0.sum(), (~np.isfinite(f0)).sum()
    for t in range(P):
        touched = np.flatnonzero((pt.theta_index == t).any(axis=1))
        if len(touched) == 0:
            continue
        for side, sign in ((1, +1.0), (0, -1.0)):
            th = theta.copy()
            th[t] += sign * eps
            f = orc.fit(th, states=False)["m2LL"]
            part[2 + (0 if side else P) + t] = (f[touched] - f0[touched]).sum()
    tpart = torch.from_numpy(part)
    dist.all_reduce(tpart)
    L = _lib.lib()
    s = L.psy_sweep_create(P)
    need, done = np.zeros(2 * P + 1, dtype=np.uint8), C.c_int(0)
    red = np.ascontiguousarray(tpart.numpy())
    _lib.check(L.psy_sweep_resolve(s, _lib.dptr(red), eps, 0, 1, 0, _lib.u8ptr(need), C.byref(done)))
    g, tot = np.zeros(P), C.c_double()
    _lib.check(L.psy_sweep_finish(s, 0, _lib.dptr(g), C.byref(tot)))
    L.psy_sweep_destroy(s)
    q.put((rank, list(pt.theta_labels), g.tolist(), tot.value))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_gradient_equals_unsharded():
    import torch.multiprocessing as mp
    from oracle.oracle import Oracle
    from psydiff_b200 import build
    build.build_library()
    N, T, world = 8, 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]          # identical on every rank
    kw = synth.config_c1(N=N, T=T, missing=0.1, seed=5)
    kw["grouping"] = "mm_Y1 | g;"
    kw["groupingvariables"] = {"g": np.array([2, 2, 2, 1, 1, 3, 3, 3], dtype=float)}
    full = pd.newPsydiff(**kw)
    assert full["pars"]["parameterTable"].theta_labels == res[0][1]   # merged labels = unsharded first-appearance order
    ref = Oracle(full).gradients()
    assert np.allclose(res[0][2], ref["grad_clean"], rtol=1e-6, atol=1e-5)
    assert res[0][3] == pytest.approx(ref["m2LL"].sum(), rel=1e-12)


def test_merge_labels_first_appearance():
    labels, values, pos = pdist.merge_labels([["a", "b_p2"], ["a", "b_p1", "c"]], [[1.0, 2.0], [1.0, 3.0, 4.0]])
    assert labels == ["a", "b_p2", "b_p1", "c"] and list(values) == [1.0, 2.0, 3.0, 4.0] and pos["c"] == 3
    assert pdist.shard_bounds(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_work_balanced_shard_bounds():
    """contiguous person blocks balanced by work Σ_t K_t (2 P_i + 1), not by count."""
    rng = np.random.default_rng(0)
    # ragged persons with irregular dt: work differs by person
    kw = synth.config_c1(N=40, T=30, seed=3)
    keep = np.ones(40 * 30, dtype=bool)
    for p in range(0, 40, 3):
        keep[p * 30 + 8:(p + 1) * 30] = False              # every third person has 8 instead of 30 time points
    d = {k: v[keep] for k, v in kw["dataset"].items()}
    d["dt"] = np.where(d["dt"] > 0, rng.choice([0.5, 1.0, 2.0], size=len(d["dt"])), 0.0)
    kw["dataset"] = d
    model = pd.newPsydiff(**kw)
    w = pdist.person_work(model)
    assert w.shape == (40,) and np.all(w > 0)
    # a person's work = (2 P + 1) x Σ_t (K_t + 1)
    L = pd._lib.lib()
    h = float(model["timeStep"][0])
    off = model["_row_off"]
    want0 = (2 * 11 + 1) * sum(L.psy_rk4_step_count(float(x), h) + 1 for x in d["dt"][off[0]:off[1]])
    assert w[0] == want0
    for world in (2, 3, 8):
        b = pdist.shard_bounds(40, world, work=w)
        assert b[0][0] == 0 and b[-1][1] == 40 and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        shares = np.array([w[lo:hi].sum() for lo, hi in b]) / w.sum()
        by_count = np.array([w[lo:hi].sum() for lo, hi in pdist.shard_bounds(40, world)]) / w.sum()
        assert np.max(np.abs(shares - 1 / world)) <= np.max(w) / w.sum() + 1e-12          # within one person of the ideal cut
        assert np.max(np.abs(shares - 1 / world)) <= np.max(np.abs(by_count - 1 / world)) + 1e-12
    assert pdist.shard_bounds(8, 4, work=np.ones(8)) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert pdist.shard_bounds(5, 2, work=np.zeros(5)) == pdist.shard_bounds(5, 2)
    with pytest.raises(ValueError):
        pdist.shard_bounds(5, 2, work=np.ones(4))
