#!/bin/bash
# Round 2: whole GPU suite with the lane-pair kernel as the n >= 8 default, ncu counter calibration of the four bench kernels,
# short bench line (8192 persons per GPU) that picks the calibration up.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02h_pytest_gpu.log
bash tools/gpu_calibrate.sh r02h_ > gpurun_out/r02h_calibrate.log 2>&1
python tools/calibrate.py gpurun_out/r02h_calib_*.csv >> gpurun_out/r02h_calibrate.log 2>&1
cp profiles/flop_calibration.json gpurun_out/r02h_flop_calibration.json
python bench.py --persons-per-gpu 8192 --steps 2 --warmup 3 --cpu-persons 64 > gpurun_out/r02h_bench_8192.json 2> gpurun_out/r02h_bench_8192.err
tail -c 1500 gpurun_out/r02h_bench_8192.json; tail -5 gpurun_out/r02h_bench_8192.err; tail -8 gpurun_out/r02h_calibrate.log
