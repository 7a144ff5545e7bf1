#!/bin/bash
# Round 2, second GPU call (1 GPU): the rewritten host side (device-built lists, streams, shards) through the whole GPU suite,
# a short bench line with every config, timing of psy_set_slots at C4 scale.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02b_pytest_gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/r02b_set_slots_timing.txt
import time, ctypes as C, numpy as np, sys
sys.path.insert(0, '.')
import psydiff_b200 as pd
from psydiff_b200 import synth, _lib
L = _lib.lib()
for N in (125_000, 1_000_000):
    kw = synth.config_c4(N=N, T=2)
    m = pd.newPsydiff(**kw)
    pd.compileModel(m)
    h = m["_handle"].ptr
    pt = m["pars"]["parameterTable"]
    tidx = np.ascontiguousarray(pt.theta_index, dtype=np.int32)
    ts = []
    for _ in range(3):
        t = time.perf_counter(); _lib.check(L.psy_set_slots(h, len(pt.theta_labels), _lib.iptr(tidx))); ts.append(time.perf_counter() - t)
    print(f"psy_set_slots C4 N={N}: instances={L.psy_n_instances(h)} seconds={[round(x, 4) for x in ts]}", flush=True)
    del m
PY
python bench.py --persons-per-gpu 8192 --steps 2 --warmup 3 --cpu-persons 64 > gpurun_out/r02b_bench_8192.json 2> gpurun_out/r02b_bench_8192.err
tail -c 3000 gpurun_out/r02b_bench_8192.json; tail -5 gpurun_out/r02b_bench_8192.err
