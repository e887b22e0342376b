#!/bin/bash
# One GPU experiment session: parity tests on the in-tree library (falling back to a second library when they fail),
# a soak run, interleaved per-kernel times of library variants (tools/exp_variants.py), controller micro-benchmarks.
# usage: tools/gpu_exp.sh <tag> <fallback.so> <variant.so> [<variant.so> ...]
tag=$1; fb=$2; shift 2
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/${tag}_gpu.txt
python -m pytest tests -q -m gpu -x > $out/${tag}_pytest.log 2>&1; rc=$?; echo "rc=$rc" >> $out/${tag}_pytest.log
if [ $rc -ne 0 ] && [ -n "$fb" ] && [ -f "$fb" ]; then
  ACS_LIB=$PWD/$fb python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_fallback.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest_fallback.log
fi
timeout 400 python tools/soak.py 600 > $out/${tag}_soak.log 2>&1; echo "rc=$?" >> $out/${tag}_soak.log
python tools/exp_variants.py --splits=-1 --rounds 1 --steps 30 --warm 100 \
  --cases 1v1:4096,1v1:65536,2v2_shoot:8192,1v1_shoot:16384,4v4:4096 "$@" > $out/${tag}_variants.log 2>&1
python tools/exp_controller_time.py > $out/${tag}_controller.log 2>&1
tail -4 $out/${tag}_pytest.log; tail -3 $out/${tag}_soak.log; cat $out/${tag}_variants.log; tail -12 $out/${tag}_controller.log
