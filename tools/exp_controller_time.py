"""Tuning experiment: time of the PyTorch side of a hierarchical step (controller + action plumbing) vs the env kernels."""
import os as _os; _os.environ.setdefault("ACS_ALLOW_RANDOM_CONTROLLER", "1")   # synthetic controller weights: timing / soak only
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.envs import BatchedEnv
n = 8192
env = BatchedEnv("2v2/ShootMissile/HierarchySelfplay", n, seed=0, substeps=12)
env.reset()
A = env.n_agents
rng = np.random.default_rng(0)
acts = torch.tensor(np.concatenate([rng.integers(0, 3, (n, A, 1)), rng.integers(0, 5, (n, A, 1)), rng.integers(0, 3, (n, A, 1)),
                                    (rng.random((n, A, env.act_dim - 3)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32), device="cuda")
for _ in range(20): env.step(acts)
def timed(fn, reps=50):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
print("whole step (graph)            %.3f ms" % timed(lambda: env.step(acts)))
g = torch.cuda.CUDAGraph()
env._warm_for_capture()
with torch.cuda.graph(g):
    env.low_level_actions(env._act_in)
print("low_level_actions (graph)     %.3f ms" % timed(g.replay))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:14]
for r in rows: print("  %-70s n=%3d  %.1f us each" % (r.key[:70], r.count, r.device_time_total / max(r.count, 1)))

# ---- micro-benchmarks of the pieces (alternatives for the LayerNorms and the K = 12 input layer)
import torch.nn.functional as F
N = n * A
x = torch.randn(N, 128, device="cuda"); w = torch.ones(128, device="cuda"); b = torch.zeros(128, device="cuda")
def ln_manual(x):
    var, mean = torch.var_mean(x, dim=-1, correction=0, keepdim=True)
    return torch.addcmul(b, (x - mean) * torch.rsqrt(var + 1e-5), w)
for name, fn in (("F.layer_norm", lambda: F.layer_norm(x, (128,), w, b)), ("var_mean + addcmul", lambda: ln_manual(x)),
                 ("native_layer_norm [N/8, 8, 128]", lambda: F.layer_norm(x.view(-1, 8, 128), (128,), w, b))):
    fn(); print("  %-34s %.1f us" % (name, 1e3 * timed(fn, 100)))
print("  max |F.layer_norm - manual| = %.2e" % float((F.layer_norm(x, (128,), w, b) - ln_manual(x)).abs().max()))
xt = x.t().contiguous(); wc = w[:, None].contiguous(); bc = b[:, None].contiguous(); ones = torch.full((128, 1), 1.0 / 128, device="cuda")
def ln_t(xt):           # transposed layout [128, N]: the statistics reduce over the strided dimension, N stays contiguous
    var, mean = torch.var_mean(xt, dim=0, correction=0, keepdim=True)
    return torch.addcmul(bc, (xt - mean) * torch.rsqrt(var + 1e-5), wc)
def ln_gemv(x):         # moments through two matrix-vector products
    m = x @ ones; d = x - m; v = (d * d) @ ones
    return torch.addcmul(b, d * torch.rsqrt(v + 1e-5), w)
for name, fn in (("transposed var_mean + addcmul", lambda: ln_t(xt)), ("var_mean over dim 0 alone", lambda: torch.var_mean(xt, dim=0, correction=0, keepdim=True)),
                 ("moments by gemv", lambda: ln_gemv(x)), ("x @ ones alone", lambda: x @ ones)):
    fn(); print("  %-34s %.1f us" % (name, 1e3 * timed(fn, 100)))
print("  max |F.layer_norm - transposed| = %.2e, gemv %.2e" % (float((F.layer_norm(x, (128,), w, b) - ln_t(xt).t()).abs().max()),
                                                              float((F.layer_norm(x, (128,), w, b) - ln_gemv(x)).abs().max())))
x12 = torch.randn(N, 12, device="cuda"); w1 = torch.randn(128, 12, device="cuda"); b1 = torch.randn(128, device="cuda")
x16 = F.pad(x12, (0, 4)); w16 = F.pad(w1, (0, 4))
for name, fn in (("linear K=12", lambda: F.linear(x12, w1, b1)), ("linear K=16 (padded)", lambda: F.linear(x16, w16, b1)),
                 ("addmm K=12", lambda: torch.addmm(b1, x12, w1.t()))):
    fn(); print("  %-34s %.1f us" % (name, 1e3 * timed(fn, 100)))
