#!/bin/bash
# Round 2, one GPU: host-side cost of the end-to-end step (upload + evaluation) at 8192 and 125 000 persons.
mkdir -p gpurun_out
python tools/exp_upload.py 8192 125000 2>&1 | grep -v "^\[" | tee gpurun_out/r02q_upload_timing.txt
