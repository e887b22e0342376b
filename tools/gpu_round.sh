#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + one full capture of the dominant kernel.
# usage: tools/gpu_round.sh <tag>
tag=${1:-rX}
out=gpurun_out
nvidia-smi -L > $out/${tag}_gpu.txt
python -c "import jsbsim" > $out/${tag}_probe_jsbsim.txt 2>&1; echo "rc=$?" >> $out/${tag}_probe_jsbsim.txt
python -m pytest tests -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "rc=$?" >> $out/${tag}_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.log 2>&1
python bench.py --steps 200 --warmup 10 > $out/${tag}_bench.log 2>&1
for w in 1v1_shoot 2v2_shoot 4v4 heading; do
  python bench.py --workload $w --steps 100 --warmup 25 --no-cpu-baseline > $out/${tag}_bench_$w.log 2>&1
done
python bench.py --envs 65536 --steps 50 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_1v1_65536.log 2>&1
python bench.py --envs 262144 --steps 30 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_1v1_262144.log 2>&1
if [ $? -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_launches.log 2>&1
  # (regex:k_env_substeps matches whichever substep kernel the batch size selects: k_env_substeps[_split|_split3|_split4])
  ncu --set full --clock-control none --import-source on -k regex:k_env_substeps --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_substeps -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_post --launch-skip 260 --launch-count 1 \
    -o $out/${tag}_post -f python bench.py --steps 300 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_post.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_env_substeps --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_substeps_65536 -f python bench.py --envs 65536 --steps 5 --warmup 3 --no-cpu-baseline --no-fp64-peak > $out/${tag}_ncu_full2.log 2>&1
fi
tail -3 $out/${tag}_pytest.log
for f in $out/${tag}_bench*.log; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], d.get("impl", "b200"), d["config"].get("scenario"), d["config"].get("envs_per_gpu"),
          "value %.2fM e2e %.2fM ms/step %s sub %s post %s reset %s fp64frac %s" % (
              d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("ms_per_step"), r.get("kernel_ms"), r.get("post_ms"), r.get("reset_ms"),
              r.get("fp64", {}).get("frac")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
