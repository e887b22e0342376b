#!/bin/bash
# One GPU session: parity tests, smoke, bench lines, ncu launch lists + full captures of the dominant kernels.
# usage: tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_gpu.txt
if [ -z "$2" ]; then
  python -m pytest tests -q -m gpu -x > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "rc=$?" >> $out/${tag}_smoke.log
fi
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.log 2>&1
python bench.py --steps 200 --warmup 10 > $out/${tag}_bench.log 2>&1
rc=$?
for w in 1v1_shoot 2v2_shoot 4v4; do
  python bench.py --workload $w --steps 100 --warmup 60 --no-cpu-baseline --no-workloads > $out/${tag}_bench_$w.log 2>&1
done
python bench.py --envs 65536 --steps 50 --warmup 5 --no-cpu-baseline --no-workloads > $out/${tag}_bench_1v1_65536.log 2>&1
if [ $rc -eq 0 ]; then
  B="python bench.py --no-cpu-baseline --no-fp64-peak --no-workloads"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file $out/${tag}_launches.csv \
    $B --steps 5 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $out/${tag}_launches_2v2shoot.csv \
    $B --workload 2v2_shoot --steps 5 --warmup 60 > $out/${tag}_ncu_launches2.log 2>&1
  # (regex:k_env_substeps matches whichever substep kernel the batch size selects: k_env_substeps[_split|_split3|_split4])
  ncu --set full --clock-control none --import-source on -k regex:k_env_substeps --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_substeps -f $B --steps 5 --warmup 3 > $out/${tag}_ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_post --launch-skip 60 --launch-count 1 \
    -o $out/${tag}_post -f $B --steps 100 --warmup 3 > $out/${tag}_ncu_post.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_substeps --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_substeps_65536 -f $B --envs 65536 --steps 5 --warmup 3 > $out/${tag}_ncu_full2.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_missiles --launch-skip 62 --launch-count 1 \
    -o $out/${tag}_missiles_2v2shoot -f $B --workload 2v2_shoot --steps 5 --warmup 60 > $out/${tag}_ncu_full3.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_substeps --launch-skip 62 --launch-count 1 \
    -o $out/${tag}_substeps_2v2shoot -f $B --workload 2v2_shoot --steps 5 --warmup 60 > $out/${tag}_ncu_full4.log 2>&1
fi
tail -3 $out/${tag}_pytest.log
tail -3 $out/${tag}_smoke.log
for f in $out/${tag}_bench*.log; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], d.get("impl", "b200"), d["config"].get("scenario"), d["config"].get("envs_per_gpu"),
          "value %.2fM e2e %.2fM ms/step %s sub %s post %s frac %s" % (
              d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("ms_per_step"), r.get("kernel_ms"), r.get("post_ms"), r.get("frac")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
