#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + one full capture of the dominant kernel.
# usage: tools/gpu_round.sh <tag>
tag=${1:-rX}
out=gpurun_out
nvidia-smi -L > $out/${tag}_gpu.txt
python -m pytest tests -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "rc=$?" >> $out/${tag}_smoke.log
python bench.py --steps 100 --warmup 10 > $out/${tag}_bench.log 2>&1
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.log 2>&1
for w in 1v1_shoot 2v2_shoot 4v4; do
  python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline > $out/${tag}_bench_$w.log 2>&1
done
python bench.py --envs 65536 --steps 30 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_1v1_65536.log 2>&1
if [ $? -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_launches.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_substeps --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_substeps -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_post --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_post -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_full_post.log 2>&1
fi
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench.log | cut -c1-600
