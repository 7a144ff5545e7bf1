#!/bin/bash
# Round 2, one GPU: RK4 stage-loop unrolling (-DPSY_STAGE_UNROLL=2/4) on C4 (n = 6), the C5 model (n = 8, lane pairs), C2 (n = 2) and n = 10.
mkdir -p gpurun_out
out=gpurun_out/r02m_stage_unroll.txt
: > $out
run() { python tools/run_model.py "$@" 2>&1 | grep -v "^\[" | tail -1 >> $out; }
for f in "" "-DPSY_STAGE_UNROLL=4" "-DPSY_STAGE_UNROLL=4 -DPSY_RK4_SMEM=1" "-DPSY_STAGE_UNROLL=4 -DPSY_BLOCK=32" ""; do run c4 8192 --flags="$f"; done
for f in "" "-DPSY_STAGE_UNROLL=2" "-DPSY_STAGE_UNROLL=4" "-DPSY_STAGE_UNROLL=4 -DPSY_PAIR_PARK=1" ""; do run c5 4096 --flags="$f"; done
for f in "" "-DPSY_STAGE_UNROLL=4"; do run c2 100000 --flags="$f"; done
for f in "" "-DPSY_STAGE_UNROLL=2" "-DPSY_STAGE_UNROLL=4"; do run k5 1024 --flags="$f"; done
cat $out
