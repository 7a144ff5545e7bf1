#!/usr/bin/env python
"""bench.py — SR-UKF person-timesteps/sec (−2LL + full central-FD gradient, FP64) on N B200 GPUs.

One "step" = one evaluation of [Σ −2LL + full finite-difference gradient] of the C4 workload of BASELINE.json
(6-latent coupled Van der Pol model, irregular time intervals, group-specific parameters, T = 100): every person's
2·P_i + 1 = 61 filter instances run as ONE batched launch of the filter kernel, followed by the K3 reduction, the
all-reduce of the partial sums (N > 1) and the D2H read of 5P + 2 doubles.

Weak scaling (default): --persons-per-gpu persons on every GPU (125 000 = the per-GPU shard of the N = 1M x 8-GPU config;
8 GPUs reproduce C4 exactly).  --strong keeps the TOTAL at --total-persons (1 000 000) and gives every GPU 1/N of it.
Inputs (600 MB/GPU of observations) are larger than the 126 MB L2.

After the timed C4 steps the other configs of BASELINE.json (C2: linear OU 100 k persons; C3: Van der Pol with person-specific
parameters under adaptive Dormand-Prince, 250 k persons; C5's model: n = 8, 16 384 persons per GPU) run for a few evaluations
each and land in "configs" with their own kernel time, roofline fraction and oracle spot check (--no-configs skips them).

    python bench.py --gpus 1 --steps 2 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P bench.py --gpus 8 ...
    python bench.py --gpus 8 --single-process    # one process, one host thread, 8 GPUs inside libpsydiff_b200.so
    python bench.py --impl reference             # the CPU path (oracle port of the reference) on the host cores
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SR-UKF person-timesteps/sec (-2LL + FD gradient, FP64)"
UNIT = "person-timesteps/s"
CALIBRATION = os.path.join(ROOT, "profiles", "flop_calibration.json")

# flops of one drift (Ff) / measurement (Fh) evaluation, counted from the model source; they enter the loop-count formula
# (psydiff_b200/flops.py) that stands in when profiles/flop_calibration.json holds no ncu count for the loaded kernel
WORKLOADS = {
    "C4": dict(n=6, m=6, Ff=39, Fh=0, desc="6-latent coupled Van der Pol (n=m=6), irregular dt, 8 groups, rk4 h=min(dt)/30"),
    "C2": dict(n=2, m=2, Ff=6, Fh=8, desc="bivariate linear OU (n=m=2), 5 % missing, rk4 h=1/30"),
    "C3": dict(n=2, m=2, Ff=7, Fh=0, desc="Van der Pol (n=m=2), mu | person (P = N + 7), adaptive Dormand-Prince (dopri5)"),
    "C5": dict(n=8, m=8, Ff=64, Fh=0, desc="4 coupled Van der Pol (n=m=8, the model of config 5), irregular dt, 8 groups, rk4 h=min(dt)/30"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--persons-per-gpu", type=int, default=125_000)
    ap.add_argument("--strong", action="store_true", help="strong scaling: --total-persons split over the GPUs")
    ap.add_argument("--total-persons", type=int, default=1_000_000)
    ap.add_argument("--timepoints", type=int, default=100)
    ap.add_argument("--cpu-persons", type=int, default=0, help="persons in the CPU sample (0 = 16 per host core, about 30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 / C5 legs after the C4 timing")
    ap.add_argument("--single-process", action="store_true", help="--gpus N inside ONE process (psy_model_set_devices) instead of torchrun ranks")
    ap.add_argument("--e2e-steps", type=int, default=3, help="end-to-end repetitions (upload + evaluation + read-back each); the line reports their median and lists all")
    ap.add_argument("--simulate-on-gpu", action="store_true", help="draw the synthetic C4 / C5 observations with torch on this rank's GPU "
                    "(another random stream than the default numpy one; for builder runs at sizes where numpy takes many minutes)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_model(name, N, T, rank, id_offset, simulate_on=None):
    """The keyword arguments of newPsydiff for `N` persons of workload `name` (this rank's shard), as a model object whose host
    copy of the dataset is in R's own layout (column-major: what model$data$observations is and what psy_upload takes)."""
    import psydiff_b200 as pd
    from psydiff_b200 import synth
    if name == "C4":
        kw = synth.config_c4(N=N, T=T, seed=20210132 + rank, id_offset=id_offset, simulate_on=simulate_on)
    elif name == "C2":
        kw = synth.config_c2(N=N, T=T, seed=20210130 + rank)
    elif name == "C3":
        kw = synth.config_c3(N=N, T=T, seed=20210131 + rank, id_offset=id_offset)
    elif name == "C5":
        kw = synth.config_c5(N=N, T=T, seed=20210133 + rank, id_offset=id_offset, simulate_on=simulate_on)
    else:
        raise ValueError(name)
    model = pd.newPsydiff(**kw)
    model["data"]["observations"] = np.asfortranarray(model["data"]["observations"])
    return model


def persons_of(args, world):
    return max(1, args.total_persons // world) if args.strong else args.persons_per_gpu


def cpu_sample(args, model, steps=1, warmup=0, per_thread=16):
    """Time the CPU path (oracle port of the reference algorithm, g++ -O2, no FMA contraction) on a bounded sample of
    the same workload with all host threads; returns (person-timesteps/s, cores, description, seconds per step).
    One person of C4 (61 sweeps x 100 time points) takes about 1.8 s on one thread: 16 persons per thread = about 30 s."""
    from oracle.oracle import Oracle
    cores = os.cpu_count() or 1
    n = args.cpu_persons or per_thread * cores
    n = min(n, model["pars"]["parameterTable"].theta_index.shape[0])
    orc = Oracle(model)
    persons = np.arange(n)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = orc.gradients(nthreads=cores, persons=persons)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    T = args.timepoints
    cpu_sample.last = {"persons": persons, "m2LL": res["m2LL"], "grad": res["grad_clean"]}
    return n * T / sec, cores, (f"first {n} persons of the C4 shard x T={T}, -2LL + full central FD gradient "
                                f"(61 filter sweeps/person), {cores} threads"), sec


def reference_build_sample(timepoints=20):
    """Informational: the REFERENCE'S OWN C++ (oracle/_ref: fitModel of R/modelTemplate.R + inst/include headers, compiled against
    the stand-in Rcpp / Armadillo / odeint headers of oracle/refbuild) timed on one person of the C4 model.  One fitModel sweep
    is timed; getGradients is literally 2P+1 = 61 such sweeps for this person (R/modelTemplate.R:852-930), so the evaluation
    time is 61 x that.  The stand-in containers are not representative of Rcpp/Armadillo speed — this is not the baseline the
    speed-up is quoted against (that is the faster oracle port), it is here so the reference's code is seen to run."""
    try:
        import psydiff_b200 as pd
        from psydiff_b200 import synth
        from oracle import reference_build as rb
        from oracle.oracle import Oracle
        m = pd.newPsydiff(**synth.config_c4(N=1, T=timepoints))
        R = rb.Reference(m)
        R.fit()
        t0 = time.perf_counter()
        f = R.fit()["m2LL"]
        sec = time.perf_counter() - t0
        o = Oracle(m).fit(states=False)["m2LL"]
        sweeps = 2 * m["pars"]["parameterTable"].theta_index.shape[1] + 1
        return {"value": timepoints / (sweeps * sec), "unit": UNIT, "cores": 1,
                "sample": f"1 person x T={timepoints} of the C4 model, one fitModel sweep timed ({sec:.3f} s) x {sweeps} sweeps of getGradients",
                "m2ll_vs_oracle_rel": float(np.max(np.abs(f - o) / np.abs(o))),
                "note": "reference C++ through stand-in Rcpp/Armadillo/odeint containers (oracle/refbuild); not representative of their speed"}
    except Exception as e:                                   # never let the informational leg break the bench line
        return {"unavailable": str(e)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, min(args.warmup, 1))
    # every step is a bounded sample: sized so that the whole run of steps + warm-up stays near three minutes (>= 2 persons
    # per thread, at most the 16 per thread the GPU arm's own cpu_baseline uses)
    per_thread = int(min(16, max(2, 180.0 / (steps + warm) / 1.8)))
    n = args.cpu_persons or per_thread * cores
    model = make_model("C4", max(64, n), args.timepoints, 0, 0)       # only the sample is evaluated
    val, cores, desc, sec = cpu_sample(args, model, steps=steps, warmup=warm, per_thread=per_thread)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 sample: 6-latent coupled Van der Pol, irregular dt, 8 groups, T=100, eps=1e-6 central, rk4 h=min(dt)/30",
                   "note": "the reference's R/Rcpp/Armadillo/odeint stack is absent (no R, Armadillo, Boost): the timed arm is the oracle "
                           "port of its algorithm, which omits the reference's per-call Rcpp::List lookups and is therefore faster "
                           "than it; cpu_baseline.reference_build times the reference's own C++ through stand-in containers"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "reference_build": reference_build_sample()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- the GPU arm ----------------------------------------------------------------------------------------------------
class Ctx:
    """Process layout of this run: torchrun ranks (one process per GPU, NCCL) or ONE process driving --gpus devices."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.in_process = 1
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: psydiff_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        elif args.single_process and args.gpus > 1:
            self.in_process = min(args.gpus, torch.cuda.device_count())
        self.n_gpus = self.world * self.in_process

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        for d in range(self.in_process):
            self.torch.cuda.synchronize(d if self.in_process > 1 else self.local)

    def reduce_max(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def load_calibration():
    try:
        return json.load(open(CALIBRATION))
    except Exception:
        return {}


def kernel_roofline(name, model, n_inst, n_persons, kernel_ms, peak, calib, T):
    """Roofline of the filter kernel of one workload: executed FP64 flops of this rank's launch / its event-timed duration.
    The flop count per instance is ncu's DADD + DMUL + 2 DFMA of the very kernel that is loaded (profiles/flop_calibration.json,
    keyed by the cubin's content hash); when the file has no entry for this kernel the loop-count formula stands in and the
    line says so (`calibration`: "stale ...") — it does not silently reuse a count measured on another kernel."""
    import ctypes as C
    import psydiff_b200 as pd
    from psydiff_b200 import _lib, codegen, flops
    L = _lib.lib()
    w = WORKLOADS[name]
    h = model["_handle"]
    key = os.path.basename(h.cubin).replace(".cubin", "")
    regs, loc, mt = C.c_int(), C.c_int(), C.c_int()
    L.psy_kernel_info(h.ptr, C.byref(regs), C.byref(loc), C.byref(mt))
    dtv = model["data"]["dt"]
    out = {"kernel": "psy_filter_kernel", "kernel_hash": key, "kernel_ms": kernel_ms, "registers": regs.value, "local_bytes": loc.value,
           "block": mt.value, "instances": int(n_inst)}
    fl_formula = None
    if model["integrateFunction"] != "runge_kutta_dopri5":
        hstep = float(model["timeStep"][0])
        qd = np.round(dtv, 3)                      # dt values are continuous; bin to 1e-3 for the step tally
        udt, cnt = np.unique(qd, return_counts=True)
        steps_of = np.array([flops.rk4_step_count(float(d), hstep) if d != 0 else 0 for d in udt])
        # tally over the model's whole dataset (all devices of this process), per person of it
        steps_pp = float((steps_of * cnt).sum()) / model["pars"]["parameterTable"].theta_index.shape[0]
        fl_formula = flops.kernel_flops(w["n"], w["m"], w["Ff"], w["Fh"], steps_pp, T, l_diag=True, dep=codegen.drift_dependency(model["_spec"]))
        out["rk4_steps_per_interval"] = steps_pp / max(T - 1, 1)
        out["flops_per_instance_loop_count"] = fl_formula
        out["flops_per_instance_survey_formula"] = flops.survey_flops(w["n"], w["m"], w["Ff"], w["Fh"], steps_pp, T)
    ent = calib.get(key)
    lanes = int(L.psy_kernel_lanes(h.ptr))
    out["lanes_per_instance_in_rk4"] = lanes
    if ent and ent.get("config") == name:
        fl = float(ent["flops_per_instance"])
        out["calibration"] = f"ncu counters of this kernel ({ent.get('source', CALIBRATION)})"
        if lanes > 1 and fl_formula is not None:
            # lane-pair kernel: the executed count includes the pair's redundant work (centre drift, mean update, padded columns);
            # the roofline numerator is the ALGORITHMIC count of the thread-per-instance formulation
            out["flops_per_instance_executed"] = fl
            fl = fl_formula
            out["calibration"] += "; lane-pair kernel: achieved/frac use the loop-count flops of the thread-per-instance formulation, not the executed count"
        out["traffic"] = float(ent["dram_bytes_per_person"]) * n_persons
        out["traffic_unit"] = "bytes/launch (ncu dram__bytes_read+write of the calibration launch of this kernel, per person x persons)"
        out["ncu_fp64_pipe_active_pct"] = ent.get("fp64_pipe_active_pct")
    else:
        fl = fl_formula
        out["calibration"] = (f"stale: profiles/flop_calibration.json has no ncu count for kernel {key} ({name}); "
                              + ("the loop-count formula of psydiff_b200/flops.py is used" if fl is not None else "no flop count for an adaptive stepper"))
        out["traffic"] = None
        print(f"[bench] WARNING {out['calibration']}", file=sys.stderr, flush=True)
    obs_bytes = model["data"]["observations"].nbytes + dtv.nbytes
    out["hbm_algorithmic_bytes"] = int(obs_bytes + 8 * n_inst)
    out["hbm_staging_gbs"] = (obs_bytes + 8 * n_inst) / (kernel_ms * 1e-3) / 1e9
    if fl is not None:
        out["flops_per_instance"] = fl
        out["achieved"] = fl * n_inst / (kernel_ms * 1e-3) / 1e12
        out["frac"] = out["achieved"] / peak
        if "flops_per_instance_survey_formula" in out:
            out["achieved_survey_formula"] = out["flops_per_instance_survey_formula"] * n_inst / (kernel_ms * 1e-3) / 1e12
    else:
        out["achieved"] = out["frac"] = None
    return out


def oracle_spot_check(ctx, model, k):
    """Every rank checks the -2LL of its first k persons against the CPU oracle on the same inputs; the max over ranks comes back."""
    import psydiff_b200 as pd
    from oracle.oracle import Oracle
    k = min(k, model["pars"]["parameterTable"].theta_index.shape[0])
    persons = np.arange(k)
    nth = max(1, (os.cpu_count() or 1) // max(1, ctx.world))
    ref = Oracle(model).fit(persons=persons, states=False, nthreads=nth)["m2LL"]
    gpu = pd.fitModel(model, states=False)["m2LL"][persons]
    same_pattern = bool(np.array_equal(np.isfinite(ref), np.isfinite(gpu)))
    fin = np.isfinite(ref) & np.isfinite(gpu)
    rel = float(np.max(np.abs(gpu[fin] - ref[fin]) / np.abs(ref[fin]))) if fin.any() else 0.0
    rel = ctx.reduce_max(rel)
    ok = ctx.reduce_sum(0.0 if same_pattern else 1.0) == 0.0
    return {"m2ll_vs_oracle_max_rel": rel, "oracle_persons": int(k) * ctx.world, "same_nonfinite_pattern": ok,
            "ranks_checked": ctx.world}


def compile_on_devices(ctx, model):
    import psydiff_b200 as pd
    from psydiff_b200 import dist as pdist
    if ctx.world > 1:
        pdist.align_parameters(model)
    pd.compileModel(model)
    if ctx.in_process > 1:
        pd.setDevices(model, ctx.in_process)


def run_config(ctx, args, name, n_local, calib, peak, evals=2):
    """One of the non-headline configs: a few gradient evaluations on this run's GPUs, device-timed, with roofline and spot check."""
    import psydiff_b200 as pd
    from psydiff_b200 import _lib, dist as pdist
    L = _lib.lib()
    T = args.timepoints
    t0 = time.perf_counter()
    model = make_model(name, n_local, T, ctx.rank, ctx.rank * n_local)
    compile_on_devices(ctx, model)
    h = model["_handle"].ptr
    setup_s = time.perf_counter() - t0
    n_inst = int(L.psy_n_instances(h))
    pdist.getGradientsSharded(model)                       # warm-up
    ctx.barrier()
    torch = ctx.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    e0.record()
    for _ in range(evals):
        g, total = pdist.getGradientsSharded(model, return_total=True)
        kms.append(L.psy_last_kernel_ms(h))
    e1.record()
    ctx.barrier()
    ms = ctx.reduce_max(e0.elapsed_time(e1)) / evals
    n_total = n_local * ctx.world
    P = len(model["pars"]["parameterTable"].theta_labels)
    out = {"workload": f"{WORKLOADS[name]['desc']}, {n_total} persons x T={T} on {ctx.n_gpus} GPU(s), P={P}, -2LL + full central FD gradient",
           "value": n_total * T / (ms * 1e-3), "unit": UNIT, "ms_per_evaluation": ms, "evaluations": evals, "persons": n_total,
           "n_theta": P, "setup_seconds": setup_s}
    out["roofline"] = kernel_roofline(name, model, n_inst / ctx.in_process, n_local / ctx.in_process, float(np.mean(kms)), peak, calib, T)
    out["frac"] = out["roofline"].get("frac")
    gv = np.array(list(g.values()))
    out["check"] = {"m2ll_total": total, "finite": bool(np.isfinite(total) and np.isfinite(gv).all()), "grad_l2": float(np.linalg.norm(gv))}
    out["check"].update(oracle_spot_check(ctx, model, 16))
    del model
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import ctypes as C
    import psydiff_b200 as pd
    from psydiff_b200 import _lib, dist as pdist

    ctx = Ctx(args)
    torch = ctx.torch
    world, rank = ctx.world, ctx.rank
    L = _lib.lib()
    calib = load_calibration()

    npg = persons_of(args, ctx.n_gpus)
    n_local = npg * ctx.in_process
    model = make_model("C4", n_local, args.timepoints, rank, rank * n_local,
                       simulate_on=f"cuda:{ctx.local}" if args.simulate_on_gpu else None)
    compile_on_devices(ctx, model)
    h = model["_handle"].ptr
    pt = model["pars"]["parameterTable"]
    T, P = args.timepoints, len(pt.theta_labels)
    N_total = n_local * world
    n_inst = int(L.psy_n_instances(h))

    peak = C.c_double(0.0)
    _lib.check(L.psy_measure_fp64_peak(C.byref(peak)))

    def step():
        return pdist.getGradientsSharded(model, return_total=True)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    L.psy_counters_reset()
    kernel_ms = []
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g, total = step()
        kernel_ms.append(L.psy_last_kernel_ms(h))
    e1.record()
    ctx.barrier()
    launches = int(L.psy_launch_count())
    ms_per_step = ctx.reduce_max(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = N_total * T / (ms_per_step * 1e-3)

    # end to end through the public API with HOST buffers: H2D of the dataset + theta, evaluation, D2H of the result;
    # repeated, every repetition device-synchronised on both sides, max over ranks each
    obs_host = model["data"]["observations"]
    h2d = obs_host.nbytes + model["data"]["dt"].nbytes + 8 * P * ctx.in_process
    d2h = 8 * (5 * P + 2)
    e2e_times = []
    for _ in range(max(1, args.e2e_steps)):
        ctx.barrier()
        t0 = time.perf_counter()
        pd.uploadData(model)
        step()
        ctx.barrier()
        e2e_times.append(ctx.reduce_max(time.perf_counter() - t0))
    e2e_s = float(np.median(e2e_times))
    e2e_val = N_total * T / e2e_s

    # parity spot check of the timed path on the very same inputs, on EVERY rank
    spot = oracle_spot_check(ctx, model, 32)

    configs = {}
    if not args.no_configs:
        sizes = {"C2": max(1, 100_000 // ctx.n_gpus) * ctx.in_process, "C3": max(1, 250_000 // ctx.n_gpus) * ctx.in_process,
                 "C5": 16_384 * ctx.in_process}
        for name in ("C2", "C3", "C5"):
            try:
                configs[name] = run_config(ctx, args, name, sizes[name], calib, peak.value)
            except Exception as e:          # a failing side leg must not lose the headline line
                configs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        ctx.close()
        return

    k_ms = float(np.mean(kernel_ms))
    roof = kernel_roofline("C4", model, n_inst / ctx.in_process, n_local / ctx.in_process, k_ms, peak.value, calib, T)
    roof.update({"bound": "fp64", "peak": peak.value, "unit": "TFLOP/s", "kernel_share_of_step": k_ms / ms_per_step,
                 "peak_source": "builder-measured: dependent-free DFMA kernel (psy_measure_fp64_peak) on this GPU; MEASURED_PEAKS.json has no "
                                "FP64 entry (nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)"})
    inst_per_person = n_inst / n_local
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (torch random stream, drawn on the GPU)" if args.simulate_on_gpu else "synthetic",
        "config": {
            "workload": f"C4: {WORKLOADS['C4']['desc']}, P={P} ({inst_per_person:.0f} filter instances/person), {npg} persons/GPU x T={T} "
                        f"({N_total} persons total), -2LL + full central FD gradient, eps=1e-6 "
                        f"(avg {roof.get('rk4_steps_per_interval', 0):.1f} steps/interval)",
            "persons_per_gpu": npg, "timepoints": T, "n_theta": P, "instances_per_gpu": int(n_inst / ctx.in_process),
            "l2": "inputs (observations %.0f MB/GPU) larger than the 126 MB L2" % (obs_host.nbytes / ctx.in_process / 1e6),
            "parallelism": (f"persons sharded over {world} rank(s) x {ctx.in_process} device(s) per process; one all-reduce of {5 * P + 2} doubles "
                            f"per step across ranks" + (", peer copies + in-order sum across the devices of a process" if ctx.in_process > 1 else "")),
        },
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "repetitions": len(e2e_times), "seconds": e2e_times, "statistic": "median of the repetitions"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "configs": configs,
    }
    gvals = np.array(list(g.values()))
    # the timed result itself, so that a bench line can be re-derived elsewhere (tools/bench_total_crosscheck.py)
    line["check"] = {"m2ll_total": total, "finite": bool(np.isfinite(total) and np.isfinite(gvals).all()),
                     "grad_l2": float(np.linalg.norm(gvals)), "grad_head": {k: float(g[k]) for k in list(g)[:4]}}
    line["check"].update(spot)
    if ctx.in_process > 1:
        line["shards"] = pd.shardInfo(model)
    if world == 1 and not args.no_cpu_baseline:
        val, cores, desc, sec = cpu_sample(args, model)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc, "seconds": sec}
        # the reference's fitModel itself is serial (R/modelTemplate.R:244): same port on ONE thread, one person
        from oracle.oracle import Oracle
        t1 = time.perf_counter()
        Oracle(model).gradients(nthreads=1, persons=np.arange(1))
        line["cpu_baseline"]["single_thread_value"] = args.timepoints / (time.perf_counter() - t1)
        line["cpu_baseline"]["reference_build"] = reference_build_sample()
        # the CPU sample doubles as a gradient spot check of the timed GPU path: the same persons as their own model
        ref = cpu_sample.last
        line["check"]["cpu_sample_m2ll_max_rel"] = float(np.max(np.abs(
            pd.fitModel(model, states=False)["m2LL"][ref["persons"]] - ref["m2LL"]) / np.abs(ref["m2LL"])))
    print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON record: whatever libraries print on file descriptor 1 while the bench runs (NCCL's
    # version banner, for one) goes to stderr; the record is written to the real stdout at the end
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = io.StringIO()
    with contextlib.redirect_stdout(_buf):
        main()
    sys.stdout.flush()
    _lines = [ln for ln in _buf.getvalue().splitlines() if ln.strip()]
    for ln in _lines[:-1]:
        print(ln, file=sys.stderr)
    if _lines:
        os.write(_real_stdout, (_lines[-1] + "\n").encode())
