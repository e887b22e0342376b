#!/usr/bin/env python
"""Generates tests/golden/controller.npz from the REFERENCE's BaselineActor (envs/JSBSim/model/baseline_actor.py, torch +
numpy only, so it imports here unmodified): a seeded random-init state dict, a sequence of inputs, and the actions / GRU
states the reference module produces step by step.  Build container only (reads /root/reference)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference/envs/JSBSim/model")
import baseline_actor  # noqa: E402

torch.manual_seed(1234)
ref = baseline_actor.BaselineActor()
with torch.no_grad():      # LayerNorm / bias defaults are 1 / 0: randomise them so a swapped weight cannot hide
    for p in ref.parameters():
        if p.dim() == 1:
            p.add_(0.3 * torch.randn_like(p))
ref.eval()
N, T = 6, 8
rng = np.random.default_rng(0)
x = rng.normal(0, 1, (T, N, 12)).astype(np.float32)
h = np.zeros((N, 1, 128), dtype=np.float32)
acts, hs = [], []
for t in range(T):
    a, h_t = ref(x[t], h)
    h = h_t.detach().numpy()
    acts.append(a.numpy())
    hs.append(h[:, 0].copy())
out = {"x": x, "actions": np.stack(acts), "h": np.stack(hs)}
for k, v in ref.state_dict().items():
    out["sd:" + k] = v.numpy()
np.savez_compressed(ROOT / "tests" / "golden" / "controller.npz", **out)
print("wrote controller.npz", {k: v.shape for k, v in out.items() if not k.startswith("sd:")})
