"""Flop / DRAM calibration of the filter kernels bench.py reports on: ncu's executed DADD + DMUL + 2 DFMA thread instructions and
dram bytes of ONE launch per config, keyed by the kernel's content hash (the cubin name), written to profiles/flop_calibration.json.

On the GPU box (tools/gpu_calibrate.sh does this):
    ncu --metrics $METRICS --clock-control none -k regex:psy_filter_kernel -s 1 -c 1 --csv --log-file gpurun_out/calib_c4.csv \
        python tools/run_model.py c4 4096 --reps 2 --meta gpurun_out/calib_c4.meta.json
Here (no GPU needed):
    python tools/calibrate.py gpurun_out/calib_*.csv        # merges into profiles/flop_calibration.json
bench.py looks its loaded kernel up by hash and says "stale" when the file has no entry for it.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "flop_calibration.json")
METRICS = ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,"
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,"
           "l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active")


def _to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * scale.get(unit, 1)


def read(csv_path):
    rows = [r for r in csv.reader(open(csv_path)) if r]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        name, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]].replace(",", "")
        out[name] = (float(val), unit)
    return out


def main(paths):
    calib = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for p in paths:
        meta = json.load(open(p.replace(".csv", ".meta.json")))
        m = read(p)
        dadd, dmul, dfma = (m[f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum"][0] for k in ("dadd", "dmul", "dfma"))
        flops = dadd + dmul + 2 * dfma
        dram = _to_bytes(*m["dram__bytes_read.sum"]) + _to_bytes(*m["dram__bytes_write.sum"])
        t, tu = m["gpu__time_duration.sum"]
        ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1.0)
        calib[meta["kernel_hash"]] = {
            "config": meta["config"], "flops_per_instance": flops / meta["instances"], "dram_bytes_per_person": dram / meta["persons"],
            "persons": meta["persons"], "T": meta["T"], "instances": meta["instances"], "dadd": dadd, "dmul": dmul, "dfma": dfma,
            "ncu_kernel_ms": ms, "ncu_tflops": flops / (ms * 1e-3) / 1e12,
            "fp64_pipe_active_pct": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0],
            "lanes_per_warp_instruction": m["smsp__thread_inst_executed_per_inst_executed.ratio"][0],
            "l1_hit_pct": m["l1tex__t_sector_hit_rate.pct"][0], "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"][0],
            "registers": meta["registers"], "local_bytes": meta["local_bytes"], "block": meta["block"], "flags": meta["flags"],
            "source": f"ncu --metrics (counters) --clock-control none, one launch of tools/run_model.py {meta['config'].lower()} {meta['persons']}; "
                      f"profiles/{os.path.basename(p)}",
        }
        print(meta["config"], meta["kernel_hash"], f"{flops / meta['instances'] / 1e6:.3f} Mflop/instance, {dram / meta['persons']:.0f} DRAM bytes/person, "
              f"{flops / (ms * 1e-3) / 1e12:.2f} TFLOP/s in the capture, FP64 pipe {calib[meta['kernel_hash']]['fp64_pipe_active_pct']:.1f} %")
    json.dump(calib, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
