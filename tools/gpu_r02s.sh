#!/bin/bash
# Round 2, one GPU, final state: smoke(), the whole GPU suite, the bench line with all configs (8192-person C4 shard).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02s_smoke.log
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02s_pytest_gpu.log
python bench.py --persons-per-gpu 8192 --steps 2 --warmup 3 --cpu-persons 64 > gpurun_out/r02s_bench_8192.json 2> gpurun_out/r02s_bench_8192.err
tail -c 600 gpurun_out/r02s_bench_8192.json; tail -n 4 gpurun_out/r02s_bench_8192.err
