#!/bin/bash
# Round 2, TWO GPUs of one box, final code: the multi-GPU tests (in-library path through the Python mirror and the R shim against one
# device), the single-process and the torchrun bench modes, and C4 at its full size strong-scaled over 2 GPUs.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_r_shim.py -m gpu -q -k "several_devices or several_gpus" 2>&1 | tail -4 | tee gpurun_out/r02r_pytest_2gpu.log
python bench.py --gpus 2 --single-process --persons-per-gpu 8192 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_2gpu_single_process.json 2> gpurun_out/r02r_bench_sp.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --strong --total-persons 1000000 --simulate-on-gpu --warmup 1 --steps 1 --e2e-steps 1 --no-configs --no-cpu-baseline > gpurun_out/r02r_bench_2gpu_strong_c4_1M.json 2> gpurun_out/r02r_bench_strong.err
tail -c 700 gpurun_out/r02r_bench_2gpu_single_process.json; echo; tail -c 900 gpurun_out/r02r_bench_2gpu_strong_c4_1M.json; tail -3 gpurun_out/r02r_bench_sp.err gpurun_out/r02r_bench_strong.err
