"""Tuning experiment: how many missile slots are launched / in flight per aircraft in the weapon configs (steady state)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec
for cfg, n in (("1v1/ShootMissile/Selfplay", 4096), ("2v2/ShootMissile/HierarchySelfplay", 2048), ("scenario3/scenario3", 1024)):
    spec = load_spec(cfg, substeps_override=12)
    b = EnvBatch(spec, n, seed=0)
    b.reset()
    rng = np.random.default_rng(0)
    A = spec.n_agents
    for t in range(150):
        act = torch.tensor(np.concatenate([rng.integers(0, 41, (n, A, 3)), rng.integers(0, 30, (n, A, 1)),
                                           (rng.random((n, A, spec.shoot_dim)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32), device="cuda")
        b.step(act, auto_reset=True)
    names_i, ai = b.arena("ac_i"); names_m, mi = b.arena("ms_i")
    nl = ai[names_i.index("n_launched")].cpu().numpy()
    S = mi.shape[1] // nl.shape[0]
    st = mi[names_m.index("status")].cpu().numpy().reshape(-1, S)
    det = mi[names_m.index("detached")].cpu().numpy().reshape(-1, S)
    vals, cnt = np.unique(st, return_counts=True)
    print(cfg, "slots/aircraft", S, "n_launched mean %.2f max %d" % (nl.mean(), nl.max()), "status histogram", dict(zip(vals.tolist(), cnt.tolist())),
          "detached", int(det.sum()), "chaff active frac %.3f" % float((ai[names_i.index("chaff_state")] == 1).float().mean()))
    # per-warp: max over lanes of launched slots (what a warp iterates per substep)
    w = nl[: (nl.shape[0] // 32) * 32].reshape(-1, 32)
    print("   per-warp max n_launched: mean %.2f ; warps with any launched: %.2f" % (w.max(1).mean(), (w.max(1) > 0).mean()))
    b.close()
