"""Tuning experiment: where the end-to-end VecEnv.step time goes (host wall clock per stage, 4096-env headline config)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.env_wrappers import BatchedVecEnv

n = 4096
ve = BatchedVecEnv("1v1/NoWeapon/Selfplay", n, device=0, seed=0, substeps=12, copy=False)
rng = np.random.default_rng(0)
acts = np.concatenate([rng.integers(0, 41, (300, n, 2, 3)), rng.integers(0, 30, (300, n, 2, 1))], axis=-1).astype(np.int32)
ve.reset()
for t in range(20):
    ve.step(acts[t])
T = {"put": 0.0, "step": 0.0, "fetch": 0.0, "post": 0.0, "sync_only": 0.0}
N = 200
torch.cuda.synchronize()
t_all = time.perf_counter()
for t in range(N):
    t0 = time.perf_counter()
    with torch.cuda.device(ve.core.device):
        ve._put_actions(acts[20 + t])
        t1 = time.perf_counter()
        ve.core.step(ve._act_dev)
        t2 = time.perf_counter()
        ve._host.copy_(ve.core.batch.out_buf, non_blocking=True)
        t2b = time.perf_counter()
        torch.cuda.current_stream(ve.core.device).synchronize()
        t3 = time.perf_counter()
    ve._pending = True
    out = ve.step_wait.__wrapped__(ve) if hasattr(ve.step_wait, "__wrapped__") else None
    t4 = time.perf_counter()
    T["put"] += t1 - t0; T["step"] += t2 - t1; T["fetch"] += t2b - t2; T["sync_only"] += t3 - t2b; T["post"] += t4 - t3
total = time.perf_counter() - t_all
print("per step (us):", {k: round(v / N * 1e6, 1) for k, v in T.items()}, "total", round(total / N * 1e6, 1))
# plain loop for comparison
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(N):
    ve.step(acts[20 + t])
print("ve.step per step (us):", round((time.perf_counter() - t0) / N * 1e6, 1))
