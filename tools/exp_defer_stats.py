"""Tuning experiment: how many envs take the deferred / the lockstep missile path, live missiles per aircraft, chaff."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec
for cfg, n in (("scenario3/scenario3", 4096), ("2v2/ShootMissile/HierarchySelfplay", 8192), ("1v1/ShootMissile/Selfplay", 16384)):
    spec = load_spec(cfg, substeps_override=12)
    b = EnvBatch(spec, n, seed=0)
    b.set_option("frame_split", 0)
    b.reset()
    rng = np.random.default_rng(0)
    A = spec.n_agents
    acc = []
    for t in range(120):
        a = np.concatenate([rng.integers(0, 41, (n, A, 3)), rng.integers(0, 30, (n, A, 1)), (rng.random((n, A, spec.shoot_dim)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32)
        b.step(torch.tensor(a, device="cuda"), auto_reset=True)
        if t >= 60 and t % 10 == 0:
            ni, ei = b.arena("env_i"); mn, mi = b.arena("ms_i"); an, ai = b.arena("ac_i")
            st = mi[mn.index("status")].view(n, A, -1)
            launched = (st == 0).sum(-1).float()
            mode = ei[ni.index("deferred")]
            acc.append((float((mode == 1).float().mean()), float((launched.sum(1) > 0).float().mean()), float(launched.mean()),
                        float(launched.max()), float((ai[an.index("chaff_state")] == 1).view(n, A).any(1).float().mean()),
                        float((ai[an.index("n_launched")]).float().mean()), float((mode == 2).float().mean())))
    m = np.mean(acc, axis=0)
    print(f"{cfg} x{n}: envs on the fast missile path {m[0]:.3f}, on the lockstep path {m[6]:.4f}  envs with LAUNCHED missiles {m[1]:.3f}  LAUNCHED per aircraft mean {m[2]:.3f} max {m[3]:.0f}  "
          f"envs with active chaff {m[4]:.3f}  launched slots per aircraft {m[5]:.2f}", flush=True)
    b.close()
