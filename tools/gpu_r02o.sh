#!/bin/bash
# Round 2, one GPU: register caps of the lean dopri5 kernel (C3) and of the 4-stage C2 kernel; then the whole GPU suite on the new defaults.
mkdir -p gpurun_out
out=gpurun_out/r02o_variants.txt
: > $out
run() { python tools/run_model.py "$@" 2>&1 | grep -v "^\[" | tail -1 >> $out; }
for f in "" "-DPSY_MIN_BLOCKS=6" "-DPSY_MIN_BLOCKS=10" "-DPSY_MIN_BLOCKS=12" "-DPSY_BLOCK=128 -DPSY_MIN_BLOCKS=4" "-DPSY_BLOCK=32 -DPSY_MIN_BLOCKS=16"; do run c3 250000 --flags="$f"; done
for f in "" "-DPSY_MIN_BLOCKS=10" "-DPSY_BLOCK=128" "-DPSY_BLOCK=32"; do run c2 100000 --flags="$f"; done
cat $out
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02o_pytest_gpu.log
