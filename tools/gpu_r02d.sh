#!/bin/bash
# Round 2, lane-pair kernel (psy_filter_pair.cuh) on one B200: n = 8 (C5 model), n = 10, and n = 6 (C4) against the thread-per-instance
# kernel; shared-memory parking levels, block sizes.  Kernel times = the library's own CUDA events, min of 3 launches.
mkdir -p gpurun_out
{
echo "# C5 model (n = 8), 2048 persons"
for f in "-DPSY_PAIR=0" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=0" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=1" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=2" \
         "-DPSY_PAIR=1 -DPSY_PAIR_PARK=2 -DPSY_PAIR_BLOCK=32" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=2 -DPSY_PAIR_BLOCK=128" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=1 -DPSY_PAIR_BLOCK=128"; do
  python tools/run_model.py c5 2048 --flags="$f"; done
PSY_SMEM_CARVEOUT=100 python tools/run_model.py c5 2048 --flags="-DPSY_PAIR=1 -DPSY_PAIR_PARK=2"
PSY_SMEM_CARVEOUT=30 python tools/run_model.py c5 2048 --flags="-DPSY_PAIR=1 -DPSY_PAIR_PARK=1"
echo "# C4 (n = 6), 8192 persons"
for f in "-DPSY_PAIR=0" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=0" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=1" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=2"; do python tools/run_model.py c4 8192 --flags="$f"; done
echo "# n = 10, 1024 persons"
for f in "-DPSY_PAIR=0" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=1" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=2"; do python tools/run_model.py k5 1024 --flags="$f"; done
echo "# n = 4 (2 coupled oscillators), 16384 persons"
for f in "-DPSY_PAIR=0" "-DPSY_PAIR=1 -DPSY_PAIR_PARK=0"; do python tools/run_model.py k2 16384 --flags="$f"; done
} 2>&1 | tee gpurun_out/r02d_pair.txt
