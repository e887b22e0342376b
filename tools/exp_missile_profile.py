"""Tuning experiment (needs a -DACS_MISSILE_PROFILE build, ACS_LIB=...): cycles per warp of k_env_missiles' lockstep section
(threatened envs) and of the rest (independent missiles), in steady state under bench.py's action distribution."""
import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200 import capi
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec

lib = capi.lib()
lib.acs_debug_missile_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_longlong * 8)()
for cfg, n in (("1v1/ShootMissile/Selfplay", 16384), ("2v2/ShootMissile/HierarchySelfplay", 8192), ("scenario3/scenario3", 4096)):
    spec = load_spec(cfg, substeps_override=12)
    b = EnvBatch(spec, n, seed=0)
    b.set_option("frame_split", 0)
    b.reset()
    rng = np.random.default_rng(0)
    A = spec.n_agents
    warm, steps = 100, 20
    for t in range(warm + steps):
        a = np.concatenate([rng.integers(0, 41, (n, A, 3)), rng.integers(0, 30, (n, A, 1)), (rng.random((n, A, spec.shoot_dim)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32)
        if t == warm:
            lib.acs_debug_missile_profile(out, 1)
            b.set_timing(True)
        b.step(torch.tensor(a, device="cuda"), auto_reset=True)
    ms, k = b.get_timing()
    lib.acs_debug_missile_profile(out, 1)
    warps = n * (1 << (A - 1).bit_length()) // 32
    print(f"{cfg} x{n}: k_env_missiles {ms['missiles'] / k * 1e3:.1f} us; per step: {out[2] / steps:.0f} of {warps} warps ran the lockstep section, "
          f"mean {out[0] / max(out[2], 1):.0f} max {out[1]} cycles; rest of the kernel ({out[5] / steps:.0f} warps) mean {out[3] / max(out[5], 1):.0f} "
          f"max {out[4]} cycles; longest warp {out[6]} cycles", flush=True)
    b.close()
