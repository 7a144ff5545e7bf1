#!/bin/bash
# Round 2, one GPU: whole GPU suite (fresh log), examples/c5_gist_fit.py at 32 768 persons with GPU-drawn data (rehearsal of the
# 8-GPU run of config 5), and the C4 config at its full size (1M persons x T = 100) on ONE GPU with the final kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02j_pytest_gpu.log
rm -f gpurun_out/r02j_c5_progress.jsonl
timeout 400 python examples/c5_gist_fit.py --persons-per-gpu 32768 --outer 3 --bfgs-iters 1 --progress gpurun_out/r02j_c5_progress.jsonl \
    > gpurun_out/r02j_c5_1gpu_32768.json 2> gpurun_out/r02j_c5.err
tail -c 1200 gpurun_out/r02j_c5_1gpu_32768.json; tail -3 gpurun_out/r02j_c5.err
timeout 600 python bench.py --gpus 1 --strong --total-persons 1000000 --simulate-on-gpu --warmup 1 --steps 1 --e2e-steps 1 --no-configs --no-cpu-baseline \
    > gpurun_out/r02j_bench_1gpu_full_c4_1M_persons.json 2> gpurun_out/r02j_bench_1M.err
tail -c 2500 gpurun_out/r02j_bench_1gpu_full_c4_1M_persons.json; tail -3 gpurun_out/r02j_bench_1M.err
