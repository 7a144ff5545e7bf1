#!/bin/bash
# Round 2, first GPU call: suite sanity, prepared rsqrt experiment, L1/shared carveout sweep of the n = 8 kernel,
# ncu --set full of the n = 8 (C5 model) and dopri5 (C3) kernels.   gpurun --timeout 900 -- 'bash tools/gpu_r02a.sh'
mkdir -p gpurun_out
O=gpurun_out/r02a_experiments.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02a_pytest_gpu.log
{
echo "# rsqrt rotations (prepared in round 1)"
python tools/run_model.py c4 8192
python tools/run_model.py c4 8192 --flags=-DPSY_RSQRT_ROT
python tools/run_model.py c5 2048
python tools/run_model.py c5 2048 --flags=-DPSY_RSQRT_ROT
echo "# n = 8: shared-memory carveout (percent of 228 KB) -> resident blocks and L1 left for the spill frame"
for cv in 25 45 60 75 100; do PSY_SMEM_CARVEOUT=$cv python tools/run_model.py c5 2048; done
for cv in 45 75; do PSY_SMEM_CARVEOUT=$cv python tools/run_model.py c5 2048 --flags=-DPSY_BLOCK=32; done
for cv in 0 25; do PSY_SMEM_CARVEOUT=$cv python tools/run_model.py c5 2048 --flags=-DPSY_RK4_SMEM=0; done
echo "# n = 6: carveout"
for cv in 0 25 50; do PSY_SMEM_CARVEOUT=$cv python tools/run_model.py c4 8192; done
echo "# other configs, final round-1 kernel"
python tools/run_model.py c2 100000
python tools/run_model.py c3 250000
python tools/run_model.py c5 16384 --reps 2
} 2>&1 | tee $O
CMD="python tools/run_model.py c5 2048 --reps 2"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02a_n8 $CMD > gpurun_out/r02a_ncu_n8.log 2>&1
CMD="python tools/run_model.py c3 65536 --reps 2"
ncu --set full --clock-control none --import-source on -k regex:psy_filter_kernel -s 1 -c 1 -o gpurun_out/r02a_c3 $CMD > gpurun_out/r02a_ncu_c3.log 2>&1
ls -la gpurun_out
