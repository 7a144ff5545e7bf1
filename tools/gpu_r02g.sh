#!/bin/bash
# Round 2: lane-pair kernel, n = 8: W-row order, register caps (3 warps per scheduler), at 2048 persons; best against thread-per-instance at 4096.
mkdir -p gpurun_out
{
for f in "" "-DPSY_PAIR_WORDER=0" "-DPSY_PAIR_MIN_BLOCKS=6" "-DPSY_PAIR_MIN_BLOCKS=6 -DPSY_PAIR_PARK=1" "-DPSY_PAIR_MIN_BLOCKS=5" "-DPSY_PAIR_MIN_BLOCKS=8 -DPSY_PAIR_PARK=1"; do python tools/run_model.py c5 2048 --flags="$f"; done
python tools/run_model.py c5 4096
} 2>&1 | tee gpurun_out/r02g_pair.txt
