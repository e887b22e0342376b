#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <launch-skip> <bench args...>   -- one `ncu --set full` capture of one launch
tag=$1; kre=$2; skip=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$kre --launch-skip $skip --launch-count 1 \
  -o gpurun_out/$tag -f python bench.py --no-cpu-baseline --no-fp64-peak --no-workloads "$@" > gpurun_out/${tag}.log 2>&1
echo "$tag rc=$?"
