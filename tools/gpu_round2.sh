#!/bin/bash
# First GPU call of the next round (about 6 minutes of box time):  gpurun --timeout 600 -- 'bash tools/gpu_round2.sh'
# Before calling: python tools/warm_kernel_cache.py; python tools/sass_hotloop.py "-DPSY_RSQRT_ROT"   (cubins travel with the snapshot)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
# prepared experiment: rsqrt-based rotations (profiles/r01_variants.txt, eighth series) on C4 and the n = 8 model
python tools/exp_flags.py "" "-DPSY_RSQRT_ROT" 2>&1 | tee gpurun_out/r02_rsqrt_rot.txt
# default bench (C4, 125 000 persons) and the CPU arm
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 2500 gpurun_out/r02_bench_default.json
# launch list + one full capture of the filter kernel on a small shard, source-level sampling included
CMD="python bench.py --persons-per-gpu 4096 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 3 -c 1 -o gpurun_out/r02_prof $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
# afterwards, here:  python tools/ncu_stall_summary.py gpurun_out/r02_prof.ncu-rep
