"""CPU test of the bench contract's reference arm: `bench.py --impl reference` runs the oracle port on the host cores and
prints ONE JSON line with the keys the driver reads (no GPU involved; under torchrun only rank 0 prints)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-rounds", "2"], capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("agent-steps/sec") and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["scenario"] == "1v1/NoWeapon/Selfplay"


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []
