#!/bin/bash
# Round 2, EIGHT GPUs: config 5 of BASELINE.json at its stated size — GIST lasso + post-lasso BFGS refit of the 8-latent model,
# N = 1 000 000 persons x T = 100 sharded over 8 x B200 (torchrun ranks, one all-reduce of 5P+2 doubles per sweep).
mkdir -p gpurun_out
rm -f gpurun_out/r02l_c5_8gpu_progress.jsonl
nvidia-smi --query-gpu=index,name,memory.used --format=csv,noheader > gpurun_out/r02l_gpus.txt 2>&1
timeout 570 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
    examples/c5_gist_fit.py --persons-per-gpu 125000 --outer 8 --bfgs-iters 2 --progress gpurun_out/r02l_c5_8gpu_progress.jsonl \
    > gpurun_out/r02l_c5_8gpu_1M.json 2> gpurun_out/r02l_c5.err
echo "rc=$?"
tail -c 1800 gpurun_out/r02l_c5_8gpu_1M.json; tail -5 gpurun_out/r02l_c5.err; tail -n 3 gpurun_out/r02l_c5_8gpu_progress.jsonl | cut -c1-300
