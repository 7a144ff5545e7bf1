"""Host-side mirror of psydiff's model object and R-facing API for the SR-UKF likelihood / gradient path.

Same names, argument meaning and error behaviour as the reference (R/prepareModel.R, R/modelTemplate.R,
src/exposeInlineFunctions.cpp); R lists become dicts, named numeric vectors become ordered dicts, matrices are
numpy arrays.  All numerics happen in the CUDA kernel behind the C ABI (``_lib``) — this module only parses,
builds the integer slot map that replaces the string-keyed parameterTable, and moves buffers.
"""
from __future__ import annotations

import copy
import ctypes as C
import warnings
from collections.abc import Mapping
from typing import Dict, List, Sequence

import numpy as np

from . import _lib, codegen
from ._lib import PsydiffError

_PERSON_WORDS = ("person", "persons", "individual", "individuals")  # R/prepareModel.R:717
_STEPPERS = {"rk4": 0, "runge_kutta_dopri5": 1}


def _rnum(v) -> str:
    """paste0()-style rendering of a number inside a label (R/prepareModel.R:719, 728)."""
    fv = float(v)
    return str(int(fv)) if fv == int(fv) else repr(fv)


def checkEquation(equation: str):
    """R/prepareModel.R:618-622."""
    if ";" not in equation:
        raise PsydiffError(f"Expected each command in ...\n{equation}\n to end with ; ")


def _parse_element(el):
    """One matrix element: a number, a numeric string, or "label = value" (R/prepareModel.R:641-654)."""
    if isinstance(el, (int, float, np.integer, np.floating)):
        return "", float(el)
    s = str(el)
    if "=" in s:
        lab, val = s.split("=", 1)
        return lab.replace(" ", ""), float(val)
    return "", float(s)


def sepNameFromValue(mat):
    """R/prepareModel.R:800-818: split a (character or numeric) matrix into labels and values."""
    arr = np.asarray(mat, dtype=object)
    if arr.ndim == 0:
        arr = arr.reshape(1, 1)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    labels = np.full(arr.shape, "", dtype=object)
    values = np.zeros(arr.shape, dtype=float)
    for i in range(arr.shape[0]):
        for j in range(arr.shape[1]):
            labels[i, j], values[i, j] = _parse_element(arr[i, j])
    return {"labels": labels, "values": values}


def sdeModelMatrix(values, labels):
    """R/prepareModel.R:783-798: combine numeric values and labels into a "label = value" character matrix."""
    values = np.asarray(values, dtype=float)
    labels = np.asarray(labels, dtype=object)
    if values.shape != labels.shape:
        raise PsydiffError("values and labels have different dimensions")
    out = np.empty(values.shape, dtype=object)
    for i in range(values.shape[0]):
        for j in range(values.shape[1]):
            lab = labels[i, j]
            empty = lab is None or lab == "" or (isinstance(lab, float) and np.isnan(lab))
            out[i, j] = repr(float(values[i, j])) if empty else f"{lab} = {float(values[i, j])!r}"
    return out


def fromMxMatrix(mxMatrixObject):
    """fromMxMatrix (R/prepareModel.R:772-781): an OpenMx-style matrix object (``values`` and ``labels`` of equal shape, given
    as attributes or dict entries; missing labels None/NaN) -> the "label = value" character matrix newPsydiff accepts."""
    get = (lambda k: mxMatrixObject[k]) if isinstance(mxMatrixObject, dict) else (lambda k: getattr(mxMatrixObject, k))
    values, labels = np.asarray(get("values"), dtype=float), np.asarray(get("labels"), dtype=object)
    out = np.empty(values.shape, dtype=object)
    for idx in np.ndindex(values.shape):
        lab = labels[idx]
        na = lab is None or (isinstance(lab, float) and np.isnan(lab))
        out[idx] = repr(float(values[idx])) if na else f"{lab} = {float(values[idx])!r}"
    return out


def chol2LogChol(matValues, matLabels, sanityChecks=True, matrixName=""):
    """R/prepareModel.R:820-842: log-Cholesky parameterisation; diagonal labels get the prefix "ln"."""
    vals = np.array(matValues, dtype=float)
    labs = np.array(matLabels, dtype=object)
    d = np.diag_indices(vals.shape[0])
    if sanityChecks and np.any(np.abs(vals[d]) < .01):
        warnings.warn(f"Setting the diagonal elements of {matrixName} to very small values will often result in errors. "
                      "The small values will be replaced by .01. Set sanityChecks = FALSE to prevent this.")
        dv = vals[d]
        dv[np.abs(dv) < .01] = .01
        vals[d] = dv
    with np.errstate(divide="ignore", invalid="ignore"):
        vals[d] = np.log(vals[d])
    out = np.empty(vals.shape, dtype=object)
    for i in range(vals.shape[0]):
        for j in range(vals.shape[1]):
            lab = labs[i, j]
            if i == j and lab != "":
                lab = "ln" + lab
            out[i, j] = repr(float(vals[i, j])) if lab == "" else f"{lab} = {float(vals[i, j])!r}"
    return out


class Container:
    """One entry of model$pars$parameterList: a scalar ("double") or a matrix, with the free-slot id of every element."""

    def __init__(self, name, kind, values, slots):
        self.name, self.kind, self.values, self.slots = name, kind, values, slots  # values/slots: 2-D arrays

    @property
    def shape(self):
        return self.values.shape


def extractParameters(elementNames: Sequence[str], parameters: Dict[str, object]):
    """R/prepareModel.R:624-702.  Returns (containers, slot table rows, startValues)."""
    containers: List[Container] = []
    rows = []  # dicts: label, value, target, targetClass, row, col
    start = {}
    for name in elementNames:
        cur = parameters[name]
        if isinstance(cur, (int, float, np.integer, np.floating)):
            slot = len(rows)
            rows.append(dict(label=name, value=float(cur), target=name, targetClass="double", row=0, col=0))
            containers.append(Container(name, "double", np.array([[float(cur)]]), np.array([[slot]])))
            start[name] = float(cur)
            continue
        arr = np.asarray(cur, dtype=object)
        if arr.ndim == 1:
            # quirk Q5: the reference tags these "colvec", which setParameterList_C rejects
            # (inst/include/psydiffBasic.h:102-111); vectors are treated as n x 1 matrices here.
            arr = arr.reshape(-1, 1)
        if arr.ndim != 2:
            raise PsydiffError(f"Unknown data type for parameter {name}. Allowed are numeric, matrix, and vector.")
        vals = np.zeros(arr.shape, dtype=float)
        slots = np.full(arr.shape, -1, dtype=np.int64)
        for ro in range(arr.shape[0]):
            for co in range(arr.shape[1]):
                lab, v = _parse_element(arr[ro, co])
                vals[ro, co] = v
                if lab != "":
                    slots[ro, co] = len(rows)
                    rows.append(dict(label=lab, value=v, target=name, targetClass="matrix", row=ro, col=co))
        containers.append(Container(name, "matrix", vals, slots))
        start[name] = vals.copy()
    return containers, rows, start


class ParameterTable:
    """model$pars$parameterTable (R/prepareModel.R:471-476) without materialising N x F rows of strings.

    The reference replicates the F slot rows once per person and relabels grouped parameters; everything the hot
    path needs from that table is ``theta_index[N, F]`` (slot -> position of its label among the unique labels).
    The R-visible columns are available as properties for small models.
    """

    def __init__(self, slot_rows, persons, theta_labels, theta_values, theta_index):
        self.slot_rows = slot_rows
        self.persons = np.asarray(persons)
        self.theta_labels: List[str] = list(theta_labels)
        self.theta_values = np.asarray(theta_values, dtype=float).copy()
        self.theta_index = np.ascontiguousarray(theta_index, dtype=np.int32)  # N x F
        self.changed = np.ones(self.theta_index.shape, dtype=bool)
        self._label_pos = {l: i for i, l in enumerate(self.theta_labels)}

    # R-visible columns (person-major, slot-minor), materialised on demand
    @property
    def label(self):
        tl = np.asarray(self.theta_labels, dtype=object)
        return tl[self.theta_index.reshape(-1)]

    @property
    def value(self):
        return self.theta_values[self.theta_index.reshape(-1)]

    @property
    def person(self):
        return np.repeat(self.persons, self.theta_index.shape[1])

    @property
    def target(self):
        return np.tile(np.asarray([r["target"] for r in self.slot_rows], dtype=object), len(self.persons))

    def __len__(self):
        return self.theta_index.size

    def copy(self):
        c = ParameterTable(self.slot_rows, self.persons, self.theta_labels, self.theta_values, self.theta_index)
        c.changed = self.changed.copy()
        return c


def makeGroups(grouping, groupingvariables, slot_rows, persons_unique):
    """R/prepareModel.R:704-734 → per-person labels → (theta_labels, theta_values, theta_index[N,F]).

    "par | person;" gives label par_p<id>; "par | g;" gives par_p<groupingvariables$g[id]> (R indexes the grouping
    vector by the person id, 1-based).  theta is ordered by first appearance in the person-major table.
    """
    N, F = len(persons_unique), len(slot_rows)
    slot_labels = [r["label"] for r in slot_rows]
    group_of: Dict[str, str] = {}
    if grouping is not None:
        checkEquation(grouping)
        g = grouping.replace("\n", "").replace(" ", "")
        for grp in [s for s in g.split(";") if s != ""]:
            parts = grp.split("|")
            if len(parts) < 2:
                raise PsydiffError(f"Cannot parse grouping statement {grp}")
            if parts[1] not in _PERSON_WORDS:
                if groupingvariables is None or parts[1] not in groupingvariables:
                    raise PsydiffError(f"Grouping variable {parts[1]} not found in groupingvariables")
            group_of[parts[0]] = parts[1]
    # label ids per (person, slot)
    labels: List[str] = []
    pos: Dict[str, int] = {}
    lab_idx = np.empty((N, F), dtype=np.int64)
    for k, lab in enumerate(slot_labels):
        if lab not in group_of:
            if lab not in pos:
                pos[lab] = len(labels)
                labels.append(lab)
            lab_idx[:, k] = pos[lab]
            continue
        gv = group_of[lab]
        if gv in _PERSON_WORDS:
            keys = np.asarray(persons_unique, dtype=float)
        else:
            gvar = groupingvariables[gv]
            if isinstance(gvar, dict):
                keys = np.asarray([gvar[p] for p in persons_unique], dtype=float)
            else:
                gvar = np.asarray(gvar, dtype=float)
                keys = gvar[np.asarray(persons_unique, dtype=np.int64) - 1]  # R: groupingIndicator[p]
        uk, inv = np.unique(keys, return_inverse=True)
        ids = np.empty(len(uk), dtype=np.int64)
        for u, kv in enumerate(uk):
            l2 = f"{lab}_p{_rnum(kv)}"
            if l2 not in pos:
                pos[l2] = len(labels)
                labels.append(l2)
            ids[u] = pos[l2]
        lab_idx[:, k] = ids[inv]
    # order theta by first appearance (person-major, slot-minor)
    flat = lab_idx.reshape(-1)
    first = np.full(len(labels), flat.size, dtype=np.int64)
    np.minimum.at(first, flat, np.arange(flat.size, dtype=np.int64))
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(labels), dtype=np.int64)
    rank[order] = np.arange(len(labels))
    theta_index = rank[lab_idx].astype(np.int32)
    theta_labels = [labels[i] for i in order]
    slot_vals = np.asarray([r["value"] for r in slot_rows], dtype=float)
    theta_values = np.zeros(len(labels))
    # value of a label = value of its first row (getParameterValues_C, inst/include/psydiffBasic.h:47-54)
    fpos = first[order]
    theta_values[:] = slot_vals[fpos % max(F, 1)] if F > 0 else 0.0
    return theta_labels, theta_values, theta_index


class PsydiffModel(dict):
    """The psydiffModel list of R/prepareModel.R:518-540 (a dict with the same field names)."""

    def __deepcopy__(self, memo):
        new = PsydiffModel()
        for k, v in self.items():
            if k in ("_handle",):
                new[k] = v  # device handle and resident dataset are shared by clones (uploadData refuses to replace it)
                if v is not None:
                    v.shared = True
            elif k == "data":
                new[k] = v  # the dataset is never mutated
            else:
                new[k] = copy.deepcopy(v, memo)
        return new


def newPsydiff(dataset, latentEquations, manifestEquations, L, Rchol, A0, m0, grouping=None, parameters=None,
               groupingvariables=None, additional=None, srUpdate=True, alpha=.2, beta=2.0, kappa=0.0, timeStep=None,
               integrateFunction="default", breakEarly=True, verbose=0, eps=(1e-8, 1e-6, 1e-5, 1e-4),
               direction="central", sanityChecks=True) -> PsydiffModel:
    """newPsydiff (R/prepareModel.R:399-545): validate, build the log-Cholesky containers and the parameter table,
    splice the user's equations into the kernel template.  ``model["cppCode"]`` holds the CUDA translation unit."""
    A0s, Ls, Rs = sepNameFromValue(A0), sepNameFromValue(L), sepNameFromValue(Rchol)
    nlatent, nmanifest = Ls["values"].shape[1], Rs["values"].shape[1]
    checkEquation(latentEquations)
    checkEquation(manifestEquations)
    if not isinstance(dataset, dict):
        raise PsydiffError("dataset must be a list")
    for nm in dataset:
        if nm not in ("person", "observations", "dt"):
            raise PsydiffError("Unknown field in dataset. Dataset must have the fields 'person',\n           'observations', 'dt'")
    obs = np.asarray(dataset["observations"], dtype=float)
    if obs.ndim != 2:
        raise PsydiffError("Field observations in dataset has to be of type matrix")
    if obs.shape[1] != nmanifest:
        raise PsydiffError("Dataset and R differ in their number of manifest variables.")
    if parameters is None or not isinstance(parameters, dict):
        raise PsydiffError("parameters must be of class data.frame")
    if not (additional is None or isinstance(additional, dict)):
        raise PsydiffError("additional must be of class data.frame or NULL")
    dt = np.asarray(dataset["dt"], dtype=float).reshape(-1)
    persons = np.asarray(dataset["person"]).reshape(-1)
    if len(dt) != obs.shape[0] or len(persons) != obs.shape[0]:
        raise PsydiffError("person, observations and dt must have the same number of rows")
    if timeStep is None:  # R/prepareModel.R:435-438
        mn = np.min(dt[dt > 0])
        timeStep = np.linspace(mn / 30, mn / 10, 5)
    timeStep = np.atleast_1d(np.asarray(timeStep, dtype=float))

    A0LogChol = chol2LogChol(A0s["values"], A0s["labels"], sanityChecks, "A0")
    LLogChol = chol2LogChol(Ls["values"], Ls["labels"], sanityChecks, "L")
    RLogChol = chol2LogChol(Rs["values"], Rs["labels"], sanityChecks, "Rchol")
    parameters = dict(parameters)
    parameters["A0LogChol"] = A0LogChol
    parameters["LLogChol"] = LLogChol
    parameters["RLogChol"] = RLogChol
    parameters["m0"] = m0
    containers, slot_rows, startValues = extractParameters(list(parameters.keys()), parameters)

    # persons: contiguous row blocks (the reference writes to min(rowInd)+timePoint, R/modelTemplate.R:361)
    row_off, persons_unique = _person_blocks(persons)
    theta_labels, theta_values, theta_index = makeGroups(grouping, groupingvariables, slot_rows, persons_unique)
    ptab = ParameterTable(slot_rows, persons_unique, theta_labels, theta_values, theta_index)

    spec = codegen.ModelSpec(nlatent=nlatent, nmanifest=nmanifest, containers=containers, n_slots=len(slot_rows),
                             latentEquations=latentEquations, manifestEquations=manifestEquations, srUpdate=bool(srUpdate),
                             stepper=_STEPPERS.get(integrateFunction, 0))
    cuda_src = codegen.emit_cuda(spec)

    model = PsydiffModel(
        pars={"parameterList": startValues, "parameterTable": ptab},
        nlatent=nlatent, nmanifest=nmanifest,
        data={"person": persons, "observations": obs, "dt": dt},
        additional=additional,
        m2LL=np.zeros(len(persons_unique)),
        predictedManifest=np.full((obs.shape[0], nmanifest), -99999.99),
        latentScores=np.full((obs.shape[0], nlatent), -99999.99),
        alpha=float(alpha), beta=float(beta), kappa=float(kappa), timeStep=timeStep,
        integrateFunction=integrateFunction, breakEarly=bool(breakEarly), verbose=verbose,
        eps=np.atleast_1d(np.asarray(eps, dtype=float)), direction=direction, srUpdate=bool(srUpdate),
        cppCode=cuda_src,
    )
    model["_spec"] = spec
    model["_row_off"] = row_off
    model["_persons_unique"] = persons_unique
    model["_handle"] = None
    return model


class _Handle:
    """Owns the psy_model* and the HBM-resident dataset; shared by clones of the model object."""

    def __init__(self, ptr, cubin):
        self.ptr, self.cubin = ptr, cubin

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().psy_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _variant(model):
    """(stepper, srUpdate) the kernel must be generated for; "default" integrates with rk4 (R/modelTemplate.R:334-344:
    dopri5 is only tried on an already non-finite state, which stays non-finite)."""
    return (1 if model["integrateFunction"] == "runge_kutta_dopri5" else 0, bool(model.get("srUpdate", True)))


def _sync_source(model):
    """Regenerate model["cppCode"] if integrateFunction / srUpdate were changed after newPsydiff."""
    spec = model["_spec"]
    st, sr = _variant(model)
    if (spec.stepper, spec.srUpdate) != (st, sr):
        spec.stepper, spec.srUpdate = st, sr
        model["cppCode"] = codegen.emit_cuda(spec)


def compileModelSource(psydiffModel, extra_flags: str = "") -> str:
    """Compile the model's CUDA TU to a cubin for sm_100a (no GPU needed).  Returns the cubin path."""
    L = _lib.lib()
    _sync_source(psydiffModel)
    buf = C.create_string_buffer(4096)
    _lib.check(L.psy_model_compile(psydiffModel["cppCode"].encode(), _lib.CSRC_DIR.encode(), _lib.CACHE_DIR.encode(),
                                   extra_flags.encode(), buf, 4096))
    return buf.value.decode()


def compileModel(psydiffModel, extra_flags: str = ""):
    """compileModel (R/prepareModel.R:553-562): compile the model for sm_100a, load it and make the dataset resident
    on the current device.  Afterwards fitModel / getGradients / getGradient work on this model object."""
    L = _lib.lib()
    cubin = compileModelSource(psydiffModel, extra_flags)
    spec = psydiffModel["_spec"]
    ptr = L.psy_model_create(spec.nlatent, spec.nmanifest, spec.n_slots, cubin.encode())
    if not ptr:
        raise PsydiffError(L.psy_last_error().decode())
    h = _Handle(ptr, cubin)
    h.variant = _variant(psydiffModel)
    h.extra_flags = extra_flags
    psydiffModel["_handle"] = h
    uploadData(psydiffModel)
    ptab = psydiffModel["pars"]["parameterTable"]
    tidx = np.ascontiguousarray(ptab.theta_index, dtype=np.int32)
    _lib.check(L.psy_set_slots(ptr, len(ptab.theta_labels), _lib.iptr(tidx)))
    return psydiffModel


def _person_blocks(persons):
    """Contiguous row blocks of the persons of a dataset (the reference writes to min(rowInd)+timePoint, R/modelTemplate.R:361)."""
    persons = np.asarray(persons).reshape(-1)
    change = np.flatnonzero(persons[1:] != persons[:-1]) + 1
    row_off = np.concatenate([[0], change, [len(persons)]]).astype(np.int64)
    persons_unique = persons[row_off[:-1]]
    if len(np.unique(persons_unique)) != len(persons_unique):
        raise PsydiffError("rows of one person must be contiguous in the dataset")
    return row_off, persons_unique


def uploadData(psydiffModel):
    """Copy model$data (host, R layout: observations rows x m column-major, dt, person blocks) into HBM.  compileModel
    calls this once; call it again after replacing model["data"] with a dataset of the SAME persons in the same order (e.g. a
    new simulation, other series lengths) — the slot map and the compiled kernel are kept.  The person blocks are recomputed
    from the new data; a dataset with other persons is refused (the slot map belongs to the persons of newPsydiff), and so is
    an upload through a handle that clones of the model share (it would swap the dataset under the other model objects)."""
    h = psydiffModel.get("_handle")
    if h is None or not h.ptr:
        raise PsydiffError("model is not compiled: call compileModel(model) first")
    data = psydiffModel["data"]
    row_off, persons_unique = _person_blocks(data["person"])
    old = np.asarray(psydiffModel["_persons_unique"])
    if len(persons_unique) != len(old) or np.any(np.asarray(persons_unique) != old):
        raise PsydiffError("uploadData: the dataset holds other persons (or another person order) than the model was built for; "
                           "build a new model with newPsydiff")
    if getattr(h, "uploaded", False) and getattr(h, "shared", False):
        raise PsydiffError("uploadData: this device handle is shared with clones of the model; compile the clone itself "
                           "(compileModel) before giving it other data")
    obs_cm = data["observations"]
    if not (isinstance(obs_cm, np.ndarray) and obs_cm.dtype == np.float64 and obs_cm.flags.f_contiguous):
        obs_cm = np.asfortranarray(obs_cm, dtype=np.float64)
    dt = np.ascontiguousarray(data["dt"], dtype=np.float64)
    if obs_cm.shape[0] != len(dt) or obs_cm.shape[0] != row_off[-1]:
        raise PsydiffError("person, observations and dt must have the same number of rows")
    psydiffModel["_row_off"] = row_off
    pid = np.ascontiguousarray(np.asarray(persons_unique, dtype=np.float64).astype(np.int32))
    _lib.check(_lib.lib().psy_upload(h.ptr, obs_cm.shape[0], len(pid), obs_cm.ctypes.data_as(_lib.c_double_p), _lib.dptr(dt),
                                     _lib.i64ptr(row_off.astype(np.int64)), _lib.iptr(pid)))
    h.uploaded = True
    return psydiffModel


def deviceCount() -> int:
    """Number of visible CUDA devices (0 without a driver)."""
    return int(_lib.lib().psy_device_count())


def setDevices(psydiffModel, devices=None):
    """Give the compiled model the GPUs it runs on: ``None`` = all visible, an int n = the current device plus the next n-1,
    or a list of device ids.  Persons are cut into contiguous blocks balanced by filter work, one block per device; fitModel,
    getGradients and the optimisers then use all of them from this one process (include/psydiff_b200.h: psy_model_set_devices).
    Returns the device ids in use."""
    h = psydiffModel.get("_handle")
    if h is None or not h.ptr:
        raise PsydiffError("model is not compiled: call compileModel(model) first")
    L = _lib.lib()
    if devices is None:
        _lib.check(L.psy_model_set_devices(h.ptr, 0, None))
    elif isinstance(devices, (int, np.integer)):
        _lib.check(L.psy_model_set_devices(h.ptr, int(devices), None))
    else:
        ids = np.ascontiguousarray(list(devices), dtype=np.int32)
        _lib.check(L.psy_model_set_devices(h.ptr, len(ids), _lib.iptr(ids)))
    return usedDevices(psydiffModel)


def usedDevices(psydiffModel):
    h = psydiffModel.get("_handle")
    if h is None or not h.ptr:
        return []
    ids = np.zeros(64, dtype=np.int32)
    n = int(_lib.lib().psy_model_devices(h.ptr, _lib.iptr(ids), 64))
    return [int(x) for x in ids[:n]]


def shardInfo(psydiffModel):
    """Per device: id, first person, person count, filter instances of a full gradient sweep, kernel ms of the last launch."""
    h = psydiffModel["_handle"]
    out = []
    for k in range(len(usedDevices(psydiffModel))):
        dev, p0, n = C.c_int(), C.c_int(), C.c_int()
        ni, ms = C.c_int64(), C.c_double()
        _lib.check(_lib.lib().psy_shard_info(h.ptr, k, C.byref(dev), C.byref(p0), C.byref(n), C.byref(ni), C.byref(ms)))
        out.append({"device": dev.value, "first_person": p0.value, "persons": n.value, "instances": ni.value, "last_kernel_ms": ms.value})
    return out


def _push_settings(model):
    h = model.get("_handle")
    if h is None or not h.ptr:
        raise PsydiffError("model is not compiled: call compileModel(model) first")
    if getattr(h, "variant", None) != _variant(model):
        # the reference reads integrateFunction at every fitModel call (R/modelTemplate.R:174); here the stepper and the
        # update variant are baked into the kernel, so a change swaps in the matching module (cached after the first time)
        cubin = compileModelSource(model, getattr(h, "extra_flags", ""))
        _lib.check(_lib.lib().psy_model_load_module(h.ptr, cubin.encode()))
        h.variant, h.cubin = _variant(model), cubin
    ts = np.ascontiguousarray(model["timeStep"], dtype=np.float64)
    stepper = _STEPPERS.get(model["integrateFunction"], 2)  # anything else is the "default" branch (R/modelTemplate.R:334)
    _lib.check(_lib.lib().psy_model_settings(h.ptr, float(model["alpha"]), float(model["beta"]), float(model["kappa"]),
                                             stepper, _lib.dptr(ts), len(ts), 1 if model.get("srUpdate", True) else 0))
    return h


def getParameterValues(psydiffModel) -> Dict[str, float]:
    """getParameterValues (inst/include/psydiffBasic.h:36-58): unique label -> value (first-appearance order, quirk Q6)."""
    pt = psydiffModel["pars"]["parameterTable"]
    return {l: float(v) for l, v in zip(pt.theta_labels, pt.theta_values)}


def setParameterValues(parameterTable: ParameterTable, parameterValues, parameterLabels=None):
    """setParameterValues (inst/include/psydiffBasic.h:4-33): write values by label, flag changed rows; mutates the table."""
    if parameterLabels is None and isinstance(parameterValues, dict):
        parameterLabels = list(parameterValues.keys())
        parameterValues = list(parameterValues.values())
    parameterValues = np.atleast_1d(np.asarray(parameterValues, dtype=float))
    if len(parameterValues) != len(parameterLabels):
        raise PsydiffError("Length of parameterValues and parameterLabels does not match")
    for v, lab in zip(parameterValues, parameterLabels):
        i = parameterTable._label_pos.get(lab)
        if i is None:
            warnings.warn("The following parameter was not found in the parameterTable: " + str(lab))
            continue
        if parameterTable.theta_values[i] != v:
            parameterTable.theta_values[i] = v
            parameterTable.changed[parameterTable.theta_index == i] = True


def clonePsydiffModel(model):
    """clonePsydiffModel (src/exposeInlineFunctions.cpp:305-308): deep copy; the device-resident dataset is shared."""
    return copy.deepcopy(model)


def fitModel(psydiffModel, skipUpdate: bool = False, states: bool = True):
    """fitModel (R/modelTemplate.R:165-420): per-person -2LL, latentScores, predictedManifest.
    ``states=False`` (not in the reference) skips the two rows x n / rows x m output matrices — 2x the input bytes.

    Like the reference it overwrites model$m2LL / $latentScores / $predictedManifest in place and clears the
    ``changed`` flags (R/modelTemplate.R:410-414).  Every person is evaluated on every call (stateless parameter
    semantics, quirk Q4); a person whose filter fails numerically gets NaN (quirk Q2)."""
    h = _push_settings(psydiffModel)
    pt = psydiffModel["pars"]["parameterTable"]
    theta = np.ascontiguousarray(pt.theta_values, dtype=np.float64)
    N = pt.theta_index.shape[0]
    rows = psydiffModel["data"]["observations"].shape[0]
    m2ll = np.empty(N)
    if states:
        lat = np.empty((rows, psydiffModel["nlatent"]), order="F")
        pred = np.empty((rows, psydiffModel["nmanifest"]), order="F")
    else:
        lat = pred = None
    _lib.check(_lib.lib().psy_fit(h.ptr, _lib.dptr(theta) if len(theta) else None, 1 if skipUpdate else 0, _lib.dptr(m2ll),
                                  lat.ctypes.data_as(_lib.c_double_p) if states else None,
                                  pred.ctypes.data_as(_lib.c_double_p) if states else None))
    psydiffModel["m2LL"] = m2ll
    if states:
        psydiffModel["latentScores"] = lat
        psydiffModel["predictedManifest"] = pred
    pt.changed[:] = False
    return {"latentScores": lat, "predictedManifest": pred, "m2LL": m2ll}


def fitTotals(psydiffModel, parameterSets, skipUpdate: bool = False):
    """Σ -2LL at several parameter vectors in ONE launch (not in the reference; it is what its callers that evaluate
    psydiffFitInternal at many candidate points would use: GIST's line search, bestOfN).  ``parameterSets``: K x P array in
    the order of getParameterValues(model), or a list of {label: value} dicts.  Returns (totals[K], nonfinite[K]); the
    model's parameter table is not modified."""
    h = _push_settings(psydiffModel)
    pt = psydiffModel["pars"]["parameterTable"]
    if len(parameterSets) and isinstance(parameterSets[0], dict):
        parameterSets = [[ps[l] for l in pt.theta_labels] for ps in parameterSets]
    th = np.ascontiguousarray(parameterSets, dtype=np.float64).reshape(-1, len(pt.theta_labels))
    K = th.shape[0]
    tot, bad = np.empty(K), np.empty(K)
    _lib.check(_lib.lib().psy_fit_many(h.ptr, _lib.dptr(th), K, 1 if skipUpdate else 0, _lib.dptr(tot), _lib.dptr(bad)))
    return tot, bad


_DIRS = {"central": 0, "left": 1, "right": 2}


def _direction(model):
    d = model["direction"]
    if d not in _DIRS:
        raise PsydiffError("Unknown direction argument. Possible are central, left and right")  # R/modelTemplate.R:868-870
    return _DIRS[d]


class NamedVector(Mapping):
    """A named numeric vector — what R hands back from getGradients (R/modelTemplate.R:928-929): labels in first-appearance order
    over one float64 array.  Reads like a dict {label: float} (indexing, iteration, keys / values / items, ==, dict(...)); the
    per-entry Python objects are only made when asked for, so a gradient with 250 000 person-specific entries (config 3) costs
    no 0.1 s of dict building per evaluation.  Read-only; ``dict(v)`` gives a mutable copy, ``np.asarray(v)`` the values."""

    __slots__ = ("_labels", "_values", "_pos")

    def __init__(self, labels, values, pos=None):
        self._labels = labels
        self._values = np.asarray(values, dtype=np.float64)
        self._pos = pos if pos is not None and len(pos) == len(labels) else None

    def _index(self):
        if self._pos is None:
            self._pos = {l: i for i, l in enumerate(self._labels)}
        return self._pos

    def __getitem__(self, label):
        return float(self._values[self._index()[label]])

    def __iter__(self):
        return iter(self._labels)

    def __len__(self):
        return len(self._labels)

    def __contains__(self, label):
        return label in self._index()

    def keys(self):
        return list(self._labels)

    def values(self):
        return self._values.tolist()

    def items(self):
        return list(zip(self._labels, self._values.tolist()))

    def __array__(self, dtype=None, copy=None):
        return self._values if dtype is None else self._values.astype(dtype)

    def __repr__(self):
        head = ", ".join(f"{l!r}: {v!r}" for l, v in zip(self._labels[:6], self._values[:6].tolist()))
        return "NamedVector({" + head + (", ..." if len(self._labels) > 6 else "") + "})"


def getGradients(psydiffModel) -> "NamedVector":
    """getGradients (R/modelTemplate.R:852-930): finite-difference gradient of Σ -2LL for every parameter, with the
    eps[] fallback; all 2P+1 filter sweeps run as one batched launch per eps level.  The model is not modified.  Returns the
    gradient as a named vector (NamedVector: reads like {label: value}, R's names(gradients)))."""
    h = _push_settings(psydiffModel)
    pt = psydiffModel["pars"]["parameterTable"]
    theta = np.ascontiguousarray(pt.theta_values, dtype=np.float64)
    eps = np.ascontiguousarray(psydiffModel["eps"], dtype=np.float64)
    g = np.empty(len(theta))
    tot = C.c_double(0.0)
    _lib.check(_lib.lib().psy_gradients(h.ptr, _lib.dptr(theta), _lib.dptr(eps), len(eps), _direction(psydiffModel),
                                        _lib.dptr(g), C.byref(tot)))
    return NamedVector(pt.theta_labels, g, getattr(pt, "_label_pos", None))


def getGradient(psydiffModel, currentParameterValues, par: int) -> Dict[str, float]:
    """getGradient (R/modelTemplate.R:932-1008): gradient for the single 0-based parameter ``par`` at the given values."""
    h = _push_settings(psydiffModel)
    pt = psydiffModel["pars"]["parameterTable"]
    if isinstance(currentParameterValues, dict):
        theta = np.array([currentParameterValues[l] for l in pt.theta_labels], dtype=np.float64)
    else:
        theta = np.ascontiguousarray(currentParameterValues, dtype=np.float64)
    eps = np.ascontiguousarray(psydiffModel["eps"], dtype=np.float64)
    g = C.c_double(0.0)
    _lib.check(_lib.lib().psy_gradient(h.ptr, _lib.dptr(theta), int(par), _lib.dptr(eps), len(eps), _direction(psydiffModel),
                                       C.byref(g)))
    return {pt.theta_labels[par]: g.value}


def compileParallelGradients(psydiffModel, nCores: int = 1):
    """compileParallelGradients (R/prepareModel.R:572-589).  The reference starts nCores PSOCK workers that each recompile the
    model and receive it again on every call.  Here the workers are GPUs inside the library: the model is compiled (if needed)
    and spread over min(nCores, visible devices) devices — one host thread, persons sharded, partial sums added over NVLink
    (psy_model_set_devices).  ``model["cl"]`` is set so callers that test it (R/optimWrapper.R:169-177) take the same route."""
    if psydiffModel.get("_handle") is None:
        compileModel(psydiffModel)
    n = max(1, min(int(nCores), deviceCount()))
    used = setDevices(psydiffModel, n)
    psydiffModel["cl"] = {"nCores": int(nCores), "backend": "cuda-batched", "devices": used}
    return psydiffModel


def stopParallelGradients(psydiffModel):
    """stopParallelGradients (R/prepareModel.R:598-600): back to the one device the model was compiled on."""
    if psydiffModel.get("cl") is not None and psydiffModel.get("_handle") is not None and len(usedDevices(psydiffModel)) > 1:
        setDevices(psydiffModel, 1)
    psydiffModel["cl"] = None


def getGradientsParallel(psydiffModel) -> "NamedVector":
    """getGradientsParallel (R/prepareModel.R:608-616): same result as getGradients, ordered by names(pars)."""
    return getGradients(psydiffModel)
