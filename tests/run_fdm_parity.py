"""Ad-hoc GPU parity + timing report (the pytest -m gpu tests assert the same quantities)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from oracle.fdm import OracleFdm
from aircombat_selfplay_b200.capi import FdmBatch
from tests.fdm_parity import oracle_named_state, random_ics, random_controls, field_scale

def compare(tag, fb, oracles):
    st = fb.get_state().cpu().numpy(); out = fb.get_outputs().cpu().numpy()
    worst = {}
    for i, f in enumerate(oracles):
        d = oracle_named_state(f)
        for k, name in enumerate(fb.state_names):
            if name not in d: continue
            err = abs(st[k, i] - d[name]) / field_scale(name, d[name])
            if err > worst.get(name, (0,))[0]: worst[name] = (err, st[k, i], d[name])
        for k, name in enumerate(fb.output_names):
            if name not in d: continue
            err = abs(out[k, i] - d[name]) / max(1.0, abs(d[name]))
            if err > worst.get("out:" + name, (0,))[0]: worst["out:" + name] = (err, out[k, i], d[name])
    top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:8]
    print(f"[{tag}] max rel err = {top[0][1][0]:.3e}")
    for name, (e, a, b) in top: print(f"    {name:40s} err={e:.3e} gpu={a:.15g} oracle={b:.15g}")

n = 64
rng = np.random.default_rng(0)
ic = random_ics(rng, n)
fb = FdmBatch(n, 1)
fb.reset(torch.tensor(ic, device="cuda"))
oracles = [OracleFdm() for _ in range(n)]
for f, c in zip(oracles, ic): f.reset(*c)
compare("reset", fb, oracles)
u = random_controls(rng, n)
fb.set_controls(torch.tensor(u, device="cuda")); fb.run(1)
for f, c in zip(oracles, u): f.set_controls(*c); f.run(1)
compare("1 frame", fb, oracles)
fb.run(11)
for f in oracles: f.run(11)
compare("12 frames", fb, oracles)
for step in range(50):
    u = random_controls(rng, n)
    fb.set_controls(torch.tensor(u, device="cuda")); fb.run(12)
    for f, c in zip(oracles, u): f.set_controls(*c); f.run(12)
    if step in (0, 9, 49): compare(f"{(step+1)*12+12} frames", fb, oracles)
# timing
N = 262144
fb2 = FdmBatch(N, 1)
ic2 = torch.tensor(random_ics(rng, N), device="cuda")
fb2.reset(ic2); fb2.set_controls(torch.tensor(random_controls(rng, N), device="cuda"))
torch.cuda.synchronize()
for _ in range(3): fb2.run(12)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fb2.run(12)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"k_fdm_run N={N} K=12: {ms:.3f} ms/step -> {N/ms*1e3:.3e} agent-steps/s, {N*12/ms*1e3:.3e} frames/s")
e0.record(); fb2.reset(ic2); e1.record(); torch.cuda.synchronize()
print(f"k_fdm_reset N={N}: {e0.elapsed_time(e1):.3f} ms")
