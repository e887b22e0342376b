#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <launch-skip> <units-per-launch> <bench args...>
# One `ncu --set full` capture of one launch, summarised ON THE BOX (tools/ncu_summary.py) so that only text travels back:
# gpurun_out/ is capped at 64 MiB and a report with imported source is 20-40 MiB.
tag=$1; kre=$2; skip=$3; units=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:$kre --launch-skip $skip --launch-count 1 \
  -o /tmp/$tag -f python bench.py --no-cpu-baseline --no-fp64-peak --no-workloads "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "$tag ncu rc=$?"
python tools/ncu_summary.py /tmp/$tag.ncu-rep gpurun_out/$tag.txt $units > /dev/null 2> gpurun_out/${tag}_summary.err
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/${tag}_sass.csv.gz
ls -la /tmp/$tag.ncu-rep gpurun_out/$tag.txt
