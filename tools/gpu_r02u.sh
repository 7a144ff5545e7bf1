#!/bin/bash
# Round 2, one GPU, final kernels: ncu launch list of a short bench command (kernel share of the step) and `ncu --set full` of the C4 kernel.
mkdir -p gpurun_out
CMDB="python bench.py --persons-per-gpu 4096 --steps 1 --warmup 3 --no-configs --no-cpu-baseline --e2e-steps 1"
$CMDB > gpurun_out/r02u_bench_4096.json 2> gpurun_out/r02u_bench_4096.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02u_launches_c4_4096persons.csv $CMDB > gpurun_out/r02u_ncu_launches.log 2>&1
CMD="python tools/run_model.py c4 4096 --reps 2"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02u_c4 $CMD > gpurun_out/r02u_ncu_c4.log 2>&1
tail -c 400 gpurun_out/r02u_bench_4096.json; tail -3 gpurun_out/r02u_ncu_launches.log; tail -3 gpurun_out/r02u_ncu_c4.log; ls -la gpurun_out | grep r02u
