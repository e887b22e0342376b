"""Tuning experiment (needs a -DACS_FRAME_PROFILE build, ACS_LIB=...): cycles per FDM stage of one thread.

The stamps serialise the stages (no overlap across them), so the total is ~20 % above the uninstrumented frame and a
stage run alone on its own warp (multi-warp frames) can differ from its share here; use for relative sizes."""
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
b.set_option("frame_split", 0)
b.reset()
rng = np.random.default_rng(0)
steps = 6
for t in range(steps):
    act = torch.tensor(np.concatenate([rng.integers(0, 41, (n, 2, 3)), rng.integers(0, 30, (n, 2, 1))], axis=-1).astype(np.int32), device="cuda")
    b.step(act, auto_reset=True)
torch.cuda.synchronize()
out = (ctypes.c_longlong * 16)()
capi.lib().acs_debug_frame_profile.argtypes = [ctypes.c_void_p]
assert capi.lib().acs_debug_frame_profile(out) == 0
names = ["propagate", "gravity", "atmosphere", "fcs", "massbalance", "auxiliary", "engine+fuel", "aero DRAG", "aero SIDE", "aero LIFT",
         "aero ROLL", "aero PITCH", "aero YAW", "accelerations"]
frames = steps * 12
tot = sum(out[:14])
for nm, c in zip(names, out):
    print(f"{nm:14s} {c / frames:8.0f} cycles/frame {100 * c / tot:5.1f}%")
print("total", tot / frames)
