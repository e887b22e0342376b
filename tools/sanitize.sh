#!/bin/bash
# compute-sanitizer over a short env-step run of every kernel (tools/sanitize.py): racecheck (shared-memory hazards of the
# multi-warp frames' exchange buffers and of the per-env group exchange), then memcheck.   usage: tools/sanitize.sh <tag> [steps]
tag=${1:-san}; steps=${2:-6}
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck memcheck; do
  timeout ${SANITIZE_TIMEOUT:-75} $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize.py $steps > $out/${tag}_${tool}.log 2>&1
  echo "rc=$?" >> $out/${tag}_${tool}.log
  grep -E "^\[sanitize\]|ERROR SUMMARY|RACECHECK SUMMARY|rc=" $out/${tag}_${tool}.log | tail -20
done
