#!/bin/bash
# Round 2, two GPUs of one box: the multi-GPU path inside libpsydiff_b200.so (one process, one host thread) through the Python mirror and
# through the R shim, against one device; short bench lines in both multi-GPU modes (torchrun ranks / one process).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_r_shim.py -m gpu -q -k "several_devices or several_gpus or lane_pair" 2>&1 | tail -6 | tee gpurun_out/r02i_pytest_2gpu.log
python bench.py --gpus 2 --single-process --persons-per-gpu 8192 --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02i_bench_2gpu_single_process.json 2> gpurun_out/r02i_bench_sp.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --persons-per-gpu 8192 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_2gpu_torchrun.json 2> gpurun_out/r02i_bench_tr.err
tail -c 600 gpurun_out/r02i_bench_2gpu_single_process.json; echo; tail -c 600 gpurun_out/r02i_bench_2gpu_torchrun.json; tail -3 gpurun_out/r02i_bench_sp.err gpurun_out/r02i_bench_tr.err
