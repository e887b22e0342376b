import sys
sys.path.insert(0, "/root/repo")
import os
os.environ["EXP_SIZES"] = ""
import numpy as np, torch
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec
def run(config, n_envs, steps=30, warm=40, split=None):
    spec = load_spec(config, substeps_override=12)
    b = EnvBatch(spec, n_envs, seed=0)
    if split is not None: b.set_option("frame_split", split)
    b.reset()
    rng = np.random.default_rng(0)
    A = spec.n_agents
    acts = torch.tensor(np.concatenate([rng.integers(0, 41, (steps + warm, n_envs, A, 3)), rng.integers(0, 30, (steps + warm, n_envs, A, 1)),
                                        (rng.random((steps + warm, n_envs, A, spec.shoot_dim)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32), device="cuda")
    for t in range(warm): b.step(acts[t], auto_reset=True)
    torch.cuda.synchronize()
    b.set_timing(True)
    for t in range(steps): b.step(acts[warm + t], auto_reset=True)
    ms, n = b.get_timing()
    names, mi = b.arena("ms_i")
    live = float((mi[names.index("status")] == 0).float().mean()) if "status" in names else -1
    b.close()
    return ms["substeps"] / n, ms["post"] / n, ms["reset"] / n, live
for cfg in ("2v2/NoWeapon/Selfplay", "2v2/ShootMissile/HierarchySelfplay", "1v1/ShootMissile/Selfplay", "scenario2/scenario2"):
    for n in (8192,):
        s, p, r, live = run(cfg, n, split=0)
        print(f"{cfg} x{n}: sub {s:.3f} post {p:.3f} reset {r:.3f} live-missile-slot-frac {live:.3f}", flush=True)
