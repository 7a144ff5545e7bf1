"""Multi-GPU evaluation: persons are sharded over ranks (one process per GPU), each rank filters its own persons and
the per-parameter partial sums are combined with ONE all-reduce per eps level (NCCL over NVLink on the device buffer).

This replaces the reference's only parallel path — getGradientsParallel's PSOCK workers, which ship the whole model
including the data to every worker on every call (R/prepareModel.R:608-616).  Persons are independent units
(R/modelTemplate.R:244), so there is no data-path collective; the all-reduce carries 5P+2 doubles.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from . import _lib
from .model import _direction, _push_settings


class _DevVec:
    """Zero-copy view of a device buffer owned by libpsydiff_b200 (consumed by torch.as_tensor)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def shard_bounds(n_persons: int, world: int, work=None):
    """Contiguous person blocks per rank.  Without ``work``: balanced by count (equal T and P_i make that balanced by work as
    well — the benchmark configs).  With ``work`` (one non-negative number per person, e.g. ``person_work(model)``): the cut
    points are placed where the running sum of work crosses r/world of the total, so ragged series, irregular dt and
    person-specific parameters (different instance counts) still give every GPU the same share of filter steps."""
    if work is None:
        b = [(n_persons * r) // world for r in range(world + 1)]
        return [(b[r], b[r + 1]) for r in range(world)]
    w = np.asarray(work, dtype=np.float64)
    if w.shape != (n_persons,) or np.any(w < 0) or not np.all(np.isfinite(w)):
        raise ValueError("work must hold one finite non-negative number per person")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    if cum[-1] <= 0:
        return shard_bounds(n_persons, world)
    b = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target, side="left"))          # first cut with cum >= target ...
        if k > 0 and target - cum[k - 1] < cum[min(k, n_persons)] - target:
            k -= 1                                                   # ... or the one before, whichever is closer
        b.append(min(max(k, b[-1]), n_persons))
    b.append(n_persons)
    return [(b[r], b[r + 1]) for r in range(world)]


def person_work(psydiffModel) -> np.ndarray:
    """Filter work of each person in one gradient evaluation: (2 P_i + 1) instances x Σ_t (K_t RK4 steps + 1 measurement update),
    K_t by odeint's step rule (psy_rk4_step_count); P_i = number of distinct θ entries the person carries.  For dopri5 the step
    count is data dependent; the rk4 count at the initial step size is used as the proxy."""
    L = _lib.lib()
    pt = psydiffModel["pars"]["parameterTable"]
    dt = np.asarray(psydiffModel["data"]["dt"], dtype=np.float64)
    h = float(np.atleast_1d(psydiffModel["timeStep"])[0])
    udt, inv = np.unique(dt, return_inverse=True)
    ksteps = np.array([L.psy_rk4_step_count(float(d), h) for d in udt], dtype=np.float64)[inv] + 1.0
    off = np.asarray(psydiffModel["_row_off"], dtype=np.int64)
    steps = np.add.reduceat(ksteps, off[:-1])
    tidx = np.asarray(pt.theta_index)
    distinct = np.array([len(np.unique(r)) for r in tidx]) if tidx.size else np.zeros(len(off) - 1)
    return steps * (2.0 * distinct + 1.0)


def allreduce_partial(partial_ptr: int, n: int, group=None) -> np.ndarray:
    """Sum the device-resident partial vector across ranks and return the reduced vector on the host."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(_DevVec(partial_ptr, n), device="cuda")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def getGradientsSharded(psydiffModel, group=None, return_total: bool = False):
    """getGradients over person shards: every rank passes its own shard's model (same parameters, same labels).
    Returns the full gradient (identical on every rank)."""
    L = _lib.lib()
    h = _push_settings(psydiffModel)
    pt = psydiffModel["pars"]["parameterTable"]
    theta = np.ascontiguousarray(pt.theta_values, dtype=np.float64)
    eps = np.ascontiguousarray(psydiffModel["eps"], dtype=np.float64)
    direction = _direction(psydiffModel)
    P = len(theta)
    n = int(L.psy_partial_len(h.ptr))
    need = np.zeros(2 * P + 1, dtype=np.uint8)
    done = C.c_int(0)
    for level in range(len(eps)):
        dptr = C.c_void_p()
        _lib.check(L.psy_grad_level(h.ptr, _lib.dptr(theta), float(eps[level]), direction,
                                    None if level == 0 else _lib.u8ptr(need), C.byref(dptr)))
        part = np.ascontiguousarray(allreduce_partial(dptr.value, n, group))
        _lib.check(L.psy_grad_resolve(h.ptr, _lib.dptr(part), float(eps[level]), level, len(eps), direction,
                                      _lib.u8ptr(need), C.byref(done)))
        if done.value:
            break
    g = np.empty(P)
    tot = C.c_double(0.0)
    _lib.check(L.psy_grad_finish(h.ptr, direction, _lib.dptr(g), C.byref(tot)))
    from .model import NamedVector
    out = NamedVector(pt.theta_labels, g, getattr(pt, "_label_pos", None))
    return (out, tot.value) if return_total else out


def merge_labels(label_lists, value_lists):
    """Global theta = shard label lists concatenated in rank order, first appearance wins.  Because shards are contiguous
    person blocks in dataset order, this IS the first-appearance order of the unsharded parameterTable."""
    labels, values, pos = [], [], {}
    for ll, vv in zip(label_lists, value_lists):
        for l, v in zip(ll, vv):
            if l not in pos:
                pos[l] = len(labels)
                labels.append(l)
                values.append(float(v))
    return labels, np.asarray(values, dtype=float), pos


def align_parameters(psydiffModel, label_lists=None, value_lists=None, group=None):
    """Give every rank's shard model the same theta vector (labels, order, length): gathers the shards' label lists,
    merges them and re-indexes this shard's slot map.  Parameters no person of this shard carries simply have no
    instances here.  Must be called before compileModel (the slot map is uploaded there)."""
    pt = psydiffModel["pars"]["parameterTable"]
    if label_lists is None:
        import torch.distributed as dist
        world = dist.get_world_size(group)
        gathered = [None] * world
        dist.all_gather_object(gathered, (list(pt.theta_labels), [float(v) for v in pt.theta_values]), group=group)
        label_lists = [g[0] for g in gathered]
        value_lists = [g[1] for g in gathered]
    labels, values, pos = merge_labels(label_lists, value_lists)
    remap = np.asarray([pos[l] for l in pt.theta_labels], dtype=np.int32)
    pt.theta_index = np.ascontiguousarray(remap[pt.theta_index], dtype=np.int32)
    pt.theta_labels = labels
    pt.theta_values = values
    pt._label_pos = pos
    return psydiffModel


def fitTotalSharded(psydiffModel, group=None) -> float:
    """Σ -2LL over ALL persons of all shards (fitModel on every rank + one all-reduce of [Σ, #nonfinite]); NaN if any
    person of any shard failed — the value psydiffFitInternal tests with anyNA (R/optimWrapper.R:153)."""
    import torch
    import torch.distributed as dist
    from .model import fitModel
    f = fitModel(psydiffModel, states=False)["m2LL"]
    t = torch.tensor([float(np.nansum(f)), float(np.isnan(f).sum())], dtype=torch.float64,
                     device="cuda" if torch.cuda.is_available() else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    tot, bad = t.tolist()
    return float("nan") if bad > 0 else tot


def psydiffFitInternalSharded(pars, parsnames=None, model=None, group=None) -> float:
    """psydiffFitInternal (R/optimWrapper.R:146-157) over person shards: identical value on every rank."""
    import sys
    from .model import setParameterValues
    from .optim import _named
    names, vals = _named(pars, parsnames)
    setParameterValues(model["pars"]["parameterTable"], vals, names)
    f = fitTotalSharded(model, group)
    return .5 * sys.float_info.max if np.isnan(f) else f


def psydiffGradientsInternalSharded(pars, parsnames=None, model=None, group=None) -> np.ndarray:
    """psydiffGradientsInternal (R/optimWrapper.R:167-190) over person shards."""
    from .model import setParameterValues
    from .optim import _named
    names, vals = _named(pars, parsnames)
    setParameterValues(model["pars"]["parameterTable"], vals, names)
    g = getGradientsSharded(model, group)
    out = np.array([g[n] for n in names], dtype=float)
    out[np.isnan(out)] = 1.0
    return out


def fitTotalsSharded(psydiffModel, parameterSets, group=None) -> np.ndarray:
    """fitTotals over person shards: K totals (NaN where any person of any shard failed), one all-reduce of 2K doubles."""
    import torch
    import torch.distributed as dist
    from .model import fitTotals
    tot, bad = fitTotals(psydiffModel, parameterSets)
    t = torch.tensor(np.concatenate([tot, bad]), dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    v = t.cpu().numpy()
    K = len(tot)
    out = v[:K].copy()
    out[v[K:] > 0] = np.nan
    return out
