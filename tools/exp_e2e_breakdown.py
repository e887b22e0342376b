"""Tuning experiment: where the end-to-end VecEnv.step time goes (host wall clock per stage, 4096-env headline config)."""
import os as _os; _os.environ.setdefault("ACS_ALLOW_RANDOM_CONTROLLER", "1")   # synthetic controller weights: timing / soak only
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
import time
N = 300
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(N): fn(t)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / N * 1e6
print("ve.step                      %.1f us" % timed(lambda t: ve.step(acts[20 + t % 250])))
print("  _put_actions(h2d=False)    %.1f us" % timed(lambda t: ve._put_actions(acts[20 + t % 250], h2d=False)))
def replay_sync(t):
    ve._g.replay(); torch.cuda.current_stream().synchronize()
print("  graph replay + sync        %.1f us" % timed(replay_sync))
g2 = torch.cuda.CUDAGraph()
ve.core._warm_for_capture()
with torch.cuda.graph(g2):
    ve.core._step_body(ve._act_dev)
def replay2(t):
    g2.replay(); torch.cuda.current_stream().synchronize()
print("  kernels-only graph + sync  %.1f us" % timed(replay2))
def h2d(t):
    ve._act_dev.copy_(ve._act_host, non_blocking=True); torch.cuda.current_stream().synchronize()
print("  H2D only + sync            %.1f us" % timed(h2d))
def d2h(t):
    ve._host.copy_(ve.core.batch.out_buf, non_blocking=True); torch.cuda.current_stream().synchronize()
print("  D2H only + sync            %.1f us" % timed(d2h))
def wait_only(t):
    ve._pending = True; ve._graphed = True; ve.step_wait()
print("  step_wait python only      %.1f us" % timed(wait_only))
