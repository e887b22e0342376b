"""A short env-step run meant to be executed UNDER compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize.py [steps]

Every substep kernel (one thread per aircraft and the two- / three- / four-warp frames, whose role warps exchange through
shared memory behind named barriers), k_env_missiles, k_env_post with the fused template reset, and the computed reset
kernels of the heading task are launched on small RAGGED batches (env counts that are no multiple of a warp or block),
with weapons, chaff and auto-reset in play.  The low-level action API is used, so no PyTorch network kernels run under the
tool.  Prints what was launched; the tool's own summary says whether it saw a hazard.

NOTE: the GPU pool this repository was built on refuses compute-sanitizer ("closed on this pool": runs under it have left GPUs
needing a reset), so no sanitizer log is committed; run without the tool the script is a ragged-batch soak of every kernel
(finite outputs, fault counter 0), which is how it was exercised here (tools/sanitize.sh is the wrapper for a pool that allows it).
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec
from tests.env_parity import close_init_states, low_init_states, random_actions

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
splits = [int(x) for x in os.environ.get("SANITIZE_SPLITS", "0,1,2,3").split(",")]
total = 0
for cfg, n, init, mode, shoot_p in (("scenario2/scenario2", 37, "close", "smooth", 0.7),        # 2v2: missiles, chaff, guns
                                    ("1v1/NoWeapon/Selfplay", 70, "low", "dive", 0.0),         # crashes -> fused template reset
                                    ("singlecontrol/heading", 45, None, "random", 0.0)):       # computed resets, re-targeting
    spec = load_spec(cfg, substeps_override=12)
    for split in splits:
        rng = np.random.default_rng(3)
        b = EnvBatch(spec, n, seed=1, device_share_obs=True)
        b.set_option("frame_split", split)
        if init == "close":
            b.set_init_states(close_init_states(spec, rng, dist_km=(3.0, 6.0)))
        elif init == "low":
            b.set_init_states(low_init_states(spec, h_ft=8450.0))
        b.reset()
        resets = 0
        for t in range(steps):
            act = torch.tensor(random_actions(rng, spec, n, mode=mode, shoot_p=shoot_p), device="cuda")
            obs, share, rew, done, info = b.step(act, auto_reset=True)
            resets += int(b.env_done.sum())
            assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
        torch.cuda.synchronize()
        names, ei = b.arena("env_i")
        faults = int(ei[names.index("faults")].sum())
        mnames, mi = b.arena("ms_i")
        used = int((mi[mnames.index("status")] >= 0).sum()) if mi.numel() else 0      # MS_INACTIVE = -1
        print(f"[sanitize] {cfg} x {n} envs, frame_split {split} (effective {b.get_option('frame_split_effective')}): {steps} steps, "
              f"{resets} env resets, {used} missiles launched, fault counter {faults}", flush=True)
        assert faults == 0
        total += steps
        b.close()
print(f"[sanitize] done: {total} env steps launched")
