#!/bin/bash
# Round 2, kernel variants on one B200 (all timed by the kernel's own CUDA events, min of 3 launches):
#   n = 6 (C4) and n = 8 (C5 model): fused-diffusion RHS, small-live-set RHS with and without chained columns, block sizes;
#   dopri5 (C3) and the n = 2 rk4 kernel (C2): register caps through launch bounds (more resident warps).
mkdir -p gpurun_out
{
echo "# C4 (n = 6), 8192 persons"
for f in "" "-DPSY_RSQRT_ROT=0" "-DPSY_RHS_FUSED=1" "-DPSY_RHS_STREAM=1" "-DPSY_RHS_STREAM=1 -DPSY_CHAIN_COLUMNS=0"; do python tools/run_model.py c4 8192 --flags="$f"; done
echo "# C5 model (n = 8), 2048 persons; default = 32-thread blocks, 45 % carveout"
for f in "" "-DPSY_RHS_FUSED=1" "-DPSY_RHS_STREAM=1" "-DPSY_RHS_STREAM=1 -DPSY_CHAIN_COLUMNS=0" "-DPSY_BLOCK=64"; do python tools/run_model.py c5 2048 --flags="$f"; done
PSY_SMEM_CARVEOUT=60 python tools/run_model.py c5 2048
PSY_SMEM_CARVEOUT=60 python tools/run_model.py c5 2048 --flags="-DPSY_RHS_STREAM=1"
PSY_SMEM_CARVEOUT=30 python tools/run_model.py c5 2048
echo "# n = 10, 1024 persons"
python tools/run_model.py k5 1024
python tools/run_model.py k5 1024 --flags="-DPSY_BLOCK=64"
echo "# C3 (dopri5, n = 2), 250000 persons: register caps"
for f in "" "-DPSY_MIN_BLOCKS=6" "-DPSY_MIN_BLOCKS=8" "-DPSY_MIN_BLOCKS=10" "-DPSY_MIN_BLOCKS=12" "-DPSY_BLOCK=128 -DPSY_MIN_BLOCKS=4"; do python tools/run_model.py c3 250000 --flags="$f"; done
echo "# C2 (rk4, n = 2), 100000 persons"
for f in "" "-DPSY_MIN_BLOCKS=12" "-DPSY_MIN_BLOCKS=16" "-DPSY_BLOCK=128"; do python tools/run_model.py c2 100000 --flags="$f"; done
} 2>&1 | tee gpurun_out/r02c_variants.txt
