"""Tuning experiment: host-facing step time with the output-slot pool (copy=True) against the single buffer (copy=False)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.env_wrappers import BatchedVecEnv
n = 4096
ve = BatchedVecEnv("1v1/NoWeapon/Selfplay", n, device=0, seed=0, substeps=12)
ve.reset()
rng = np.random.default_rng(0)
acts = rng.integers(0, 30, (64, n, 2, 4)).astype(np.int32)

def loop(hold, steps=300):
    out = None
    t_async = t_wait = 0.0
    for t in range(20):
        out = ve.step(acts[t % 64])
        if not hold: out = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(steps):
        a = time.perf_counter(); ve.step_async(acts[t % 64]); b = time.perf_counter(); o = ve.step_wait(); c = time.perf_counter()
        t_async += b - a; t_wait += c - b
        out = o if hold else None
        o = None
    tot = (time.perf_counter() - t0) / steps * 1e3
    return tot, t_async / steps * 1e3, t_wait / steps * 1e3, len(ve._slots)

for copy, hold in ((False, True), (True, False), (True, True), (False, True), (True, True)):
    ve.copy = copy
    print(f"copy={copy} hold_last={hold}: total %.3f ms  step_async %.3f  step_wait %.3f  slots %d" % loop(hold), flush=True)
