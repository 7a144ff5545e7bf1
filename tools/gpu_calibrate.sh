#!/bin/bash
# ncu counter pass of the filter kernel of every bench config (one launch each); parse afterwards with tools/calibrate.py.
# usage on the GPU box: bash tools/gpu_calibrate.sh [tag]     (outputs gpurun_out/<tag>calib_<cfg>.{csv,meta.json})
TAG=${1:-}
mkdir -p gpurun_out
METRICS=$(python -c "import sys; sys.path.insert(0,'tools'); import calibrate; print(calibrate.METRICS)")
for spec in "c4 4096" "c2 32768" "c3 65536" "c5 2048"; do
  set -- $spec
  python tools/run_model.py $1 $2 --reps 2 > /dev/null 2>&1    # plain run first: the command must exit 0 without ncu
  ncu --metrics $METRICS --clock-control none -k regex:psy_filter_kernel -s 1 -c 1 --csv --log-file gpurun_out/${TAG}calib_$1.csv \
      python tools/run_model.py $1 $2 --reps 2 --meta gpurun_out/${TAG}calib_$1.meta.json > gpurun_out/${TAG}calib_$1.log 2>&1
  tail -1 gpurun_out/${TAG}calib_$1.log
done
