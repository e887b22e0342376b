"""Tuning experiment (needs a -DACS_SPLIT_PROFILE build, ACS_LIB=...): cycle stamps of one pair of the two-warp frame.

NOTE: ptxas schedules the clock reads freely relative to BAR.SYNC, so the stamps do not delimit the barrier waits
reliably; the per-instruction `stall_barrier` samples of an ncu capture (tools/ncu_summary.py) are
what the role-balance numbers in DESIGN.md come from.  Kept for coarse per-frame totals only."""
import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200 import capi
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec

spec = load_spec("1v1/NoWeapon/Selfplay", substeps_override=12)
n = 4096
b = EnvBatch(spec, n, seed=0)
b.set_option("frame_split", 1)
b.reset()
rng = np.random.default_rng(0)
for t in range(6):
    act = torch.tensor(np.concatenate([rng.integers(0, 41, (n, 2, 3)), rng.integers(0, 30, (n, 2, 1))], axis=-1).astype(np.int32), device="cuda")
    b.step(act, auto_reset=True)
torch.cuda.synchronize()
out = (ctypes.c_longlong * 16)()
capi.lib().acs_debug_split_profile.argtypes = [ctypes.c_void_p]
assert capi.lib().acs_debug_split_profile(out) == 0
a, bb = list(out[:8]), list(out[8:])
t0 = a[6]
print(os.environ.get("ACS_LIB"), "timeline of frame 6 of pair 0 (cycles after role A's loop top)")
print("A: loop top 0 | arrive b1 %d release %d | arrive b2 %d release %d | arrive b3 %d release %d | next loop top %d" % tuple(x - t0 for x in a[:6] + [a[7]]))
print("B:            | arrive b1 %d release %d | arrive b2 %d release %d | arrive b3 %d release %d" % tuple(x - t0 for x in bb[:6]))
