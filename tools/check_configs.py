#!/usr/bin/env python
"""Build-container check: every yaml the reference ships (envs/JSBSim/configs/**) resolves to the same TaskSpec as the
yaml of the same name shipped in aircombat_selfplay_b200/configs (written by tools/make_configs.py)."""
import dataclasses
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from aircombat_selfplay_b200.tasks import CONFIG_DIR, load_spec  # noqa: E402

REF = Path("/root/reference/envs/JSBSim/configs")
bad = 0
for f in sorted(REF.glob("*/*.yaml")):
    name = f"{f.parent.name}/{f.stem}"
    try:
        a = load_spec(name, config_dir=str(REF))
    except NotImplementedError as e:
        print(f"[skip] {name}: {str(e)[:70]}")
        continue
    if not (CONFIG_DIR / f"{name}.yaml").exists():
        print(f"[MISSING] {name}"); bad += 1; continue
    b = load_spec(name)
    da, db = dataclasses.asdict(a), dataclasses.asdict(b)
    diff = {k: (da[k], db[k]) for k in da if da[k] != db[k]}
    if diff:
        bad += 1
        print(f"[DIFF] {name}: {diff}")
    else:
        print(f"[ok] {name}")
sys.exit(1 if bad else 0)
