#!/bin/bash
# Round 2: lane-pair kernel after the stall-driven changes (W rows accumulate oldest-first, stage update fused into the dA columns,
# source-lane shuffles), n = 8 at 2048 and 4096 persons against the thread-per-instance kernel, n = 10.
mkdir -p gpurun_out
{
for f in "" "-DPSY_PAIR_PARK=1" "-DPSY_PAIR_PARK=0" "-DPSY_PAIR_DIRECT=0" "-DPSY_PAIR_BLOCK=32" "-DPSY_PAIR_BLOCK=128"; do python tools/run_model.py c5 2048 --flags="$f"; done
for f in "" "-DPSY_PAIR=0"; do python tools/run_model.py c5 4096 --flags="$f"; done
python tools/run_model.py k5 1024
} 2>&1 | tee gpurun_out/r02f_pair.txt
