#!/bin/bash
# Round 2, one GPU: two RK4 stages per loop body at n = 6, stage unrolling at n = 4, step unrolling at n = 2, the lean dopri5 controller on C3.
mkdir -p gpurun_out
out=gpurun_out/r02n_variants.txt
: > $out
run() { python tools/run_model.py "$@" 2>&1 | grep -v "^\[" | tail -1 >> $out; }
for f in "" "-DPSY_STAGE_UNROLL=2"; do run c4 8192 --flags="$f"; done
for f in "" "-DPSY_STAGE_UNROLL=2" "-DPSY_STAGE_UNROLL=4"; do run k2 16384 --flags="$f"; done
for f in "" "-DPSY_STEP_UNROLL=2" "-DPSY_STAGE_UNROLL=1"; do run c2 100000 --flags="$f"; done
for f in "" "-DPSY_DOPRI_LEAN=1" ""; do run c3 250000 --flags="$f"; done
cat $out
