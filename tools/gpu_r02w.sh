#!/bin/bash
# Round 2, one GPU: the new device test of the late kernel defaults (stage unrolling, lean dopri5 controller).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "stage_unrolling" 2>&1 | tail -5 | tee gpurun_out/r02w_pytest_gpu_new_test.log
