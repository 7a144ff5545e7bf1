#!/bin/bash
# Round 2, one GPU, final kernels: ncu counter calibration of the four bench kernels, bench line with all configs on it, and
# `ncu --set full` captures of the new C3 (lean dopri5 controller) and C2 (four RK4 stages per loop body) kernels.
mkdir -p gpurun_out
bash tools/gpu_calibrate.sh r02p_ > gpurun_out/r02p_calibrate.log 2>&1
python tools/calibrate.py gpurun_out/r02p_calib_*.csv >> gpurun_out/r02p_calibrate.log 2>&1
cp profiles/flop_calibration.json gpurun_out/r02p_flop_calibration.json
python bench.py --persons-per-gpu 8192 --steps 2 --warmup 3 --cpu-persons 64 > gpurun_out/r02p_bench_8192.json 2> gpurun_out/r02p_bench_8192.err
tail -c 1200 gpurun_out/r02p_bench_8192.json; tail -5 gpurun_out/r02p_bench_8192.err; tail -8 gpurun_out/r02p_calibrate.log
CMD="python tools/run_model.py c3 65536 --reps 2"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02p_c3 $CMD > gpurun_out/r02p_ncu_c3.log 2>&1
CMD="python tools/run_model.py c2 32768 --reps 2"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02p_c2 $CMD > gpurun_out/r02p_ncu_c2.log 2>&1
ls -la gpurun_out | tail -12
