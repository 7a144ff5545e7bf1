#!/bin/bash
# Round 2, one GPU: second rehearsal of examples/c5_gist_fit.py (lambda and GIST's initial step size scaled with N, post-lasso BFGS refit).
mkdir -p gpurun_out
rm -f gpurun_out/r02k_c5_progress.jsonl
timeout 400 python examples/c5_gist_fit.py --persons-per-gpu 32768 --outer 8 --bfgs-iters 2 --progress gpurun_out/r02k_c5_progress.jsonl \
    > gpurun_out/r02k_c5_1gpu_32768.json 2> gpurun_out/r02k_c5.err
tail -c 1500 gpurun_out/r02k_c5_1gpu_32768.json; tail -3 gpurun_out/r02k_c5.err
