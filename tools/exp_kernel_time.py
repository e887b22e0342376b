"""Tuning experiment: k_env_substeps time for one library build (ACS_LIB) at several batch sizes."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec

def run(config, n_envs, steps=20, warm=5, split=None):
    spec = load_spec(config, substeps_override=12)
    b = EnvBatch(spec, n_envs, seed=0)
    if split is not None:
        b.set_option("frame_split", split)
    b.reset()
    rng = np.random.default_rng(0)
    A = spec.n_agents
    acts = torch.tensor(np.concatenate([rng.integers(0, 41, (steps + warm, n_envs, A, 3)), rng.integers(0, 30, (steps + warm, n_envs, A, 1)),
                                        (rng.random((steps + warm, n_envs, A, spec.shoot_dim)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32), device="cuda")
    for t in range(warm):
        b.step(acts[t], auto_reset=True)
    torch.cuda.synchronize()
    b.set_timing(True)
    for t in range(steps):
        b.step(acts[warm + t], auto_reset=True)
    ms, n = b.get_timing()
    b.close()
    return ms["substeps"] / n, ms["post"] / n, ms["reset"] / n

tag = os.environ.get("ACS_LIB", "default").split("/")[-1]
sizes = [int(x) for x in os.environ.get("EXP_SIZES", "4096,16384,65536").split(",")]
for split in (0, 1, 2, 3):
    out = [tag, f"split={split}"]
    for config, n in [("1v1/NoWeapon/Selfplay", n) for n in sizes] + [("2v2/ShootMissile/HierarchySelfplay", 8192)]:
        s, p, r = run(config, n, split=split)
        A = load_spec(config).n_agents
        out.append(f"{config.split('/')[0]}x{n}: sub {s:.3f} ms post {p:.3f} reset {r:.3f} -> {n * A / (s + p + r) / 1e3:.1f} M/s")
    print(" | ".join(out), flush=True)
