#!/bin/bash
# Round 2: ncu --set full of the lane-pair kernel and of the thread-per-instance kernel at n = 8 (C5 model, 2048 persons), plus two variants.
mkdir -p gpurun_out
{
for f in "-DPSY_PAIR_PARK=2" "-DPSY_PAIR_PARK=2 -DPSY_PAIR_DIRECT=1" "-DPSY_PAIR_PARK=2 -DPSY_PAIR_BLOCK=32 -DPSY_PAIR_DIRECT=1"; do python tools/run_model.py c5 2048 --flags="$f"; done
} 2>&1 | tee gpurun_out/r02e_pair.txt
CMD="python tools/run_model.py c5 2048 --reps 2 --flags=-DPSY_PAIR_PARK=2"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02e_n8_pair $CMD > gpurun_out/r02e_ncu_n8_pair.log 2>&1
CMD="python tools/run_model.py c5 2048 --reps 2 --flags=-DPSY_PAIR=0"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02e_n8_single $CMD > gpurun_out/r02e_ncu_n8_single.log 2>&1
ls -la gpurun_out | tail -5
