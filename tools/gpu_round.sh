#!/bin/bash
# One GPU session: parity tests, smoke, bench lines, ncu launch lists + summarised full captures of the dominant kernels.
# usage: tools/gpu_round.sh <tag> [skip-tests]      (only text comes back: gpurun_out/ is capped at 64 MiB)
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
python bench.py --envs 65536 --steps 50 --warmup 5 --no-cpu-baseline --no-workloads > $out/${tag}_bench_1v1_65536.log 2>&1
if [ $rc -eq 0 ]; then
  B="python bench.py --no-cpu-baseline --no-fp64-peak --no-workloads"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file $out/${tag}_launches_4096envs.csv \
    $B --steps 5 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_2v2shoot.csv \
    $B --workload 2v2_shoot --steps 5 --warmup 20 > $out/${tag}_ncu_launches2.log 2>&1
  # (regex:k_env_substeps matches whichever substep kernel the batch size selects: k_env_substeps[_split|_split3|_split4])
  tools/ncu_capture.sh ${tag}_k_env_substeps_split3_4096envs k_env_substeps 4 8192 --steps 5 --warmup 3
  tools/ncu_capture.sh ${tag}_k_env_post_4096envs k_env_post 60 8192 --steps 100 --warmup 3
  tools/ncu_capture.sh ${tag}_k_env_substeps_65536envs k_env_substeps 4 131072 --envs 65536 --steps 5 --warmup 3
  tools/ncu_capture.sh ${tag}_k_env_substeps_2v2shoot k_env_substeps 62 32768 --workload 2v2_shoot --steps 5 --warmup 60
  tools/ncu_capture.sh ${tag}_k_env_missiles_2v2shoot k_env_missiles 62 32768 --workload 2v2_shoot --steps 5 --warmup 60
  tools/ncu_capture.sh ${tag}_k_env_post_2v2shoot k_env_post 62 32768 --workload 2v2_shoot --steps 5 --warmup 60
  rm -f $out/*_sass.csv.gz
fi
tail -3 $out/${tag}_pytest.log
tail -4 $out/${tag}_smoke.log
for f in $out/${tag}_bench*.log; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], d.get("impl", "b200"), d["config"].get("scenario"), d["config"].get("envs_per_gpu"),
          "value %.2fM e2e %.2fM ms/step %s sub %s post %s frac %s" % (
              d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("ms_per_step"), r.get("kernel_ms"), r.get("post_ms"), r.get("frac")))
    for k, w in (d.get("workloads") or {}).items():
        if isinstance(w, dict) and "value" in w:
            print("   ", k, "value %.2fM e2e %.2fM ms/step %.4f sub %.4f missiles %.4f post %.4f frac %.4f" % (
                w["value"] / 1e6, w["e2e"]["value"] / 1e6, w["ms_per_step"], w["kernel_ms"], w["missiles_ms"], w["post_ms"], w["roofline_fp64_frac"] or 0))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
