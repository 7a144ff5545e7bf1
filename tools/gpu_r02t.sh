#!/bin/bash
# Round 2, one GPU: GPU suite and bench line after the gradient result became a NamedVector (no 250 000-entry dict per C3 evaluation).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02t_pytest_gpu.log
python bench.py --persons-per-gpu 8192 --steps 2 --warmup 3 --cpu-persons 64 > gpurun_out/r02t_bench_8192.json 2> gpurun_out/r02t_bench_8192.err
tail -c 300 gpurun_out/r02t_bench_8192.json; tail -n 4 gpurun_out/r02t_bench_8192.err
