"""What-if for the low-level controller: the same PyTorch controller with TF32 tensor-core matmuls
allowed -- how much faster, and how many arg-max decisions it changes against the fp32 arithmetic the reference uses.
Not a product setting: the controller's discrete outputs are compared exactly with the reference's (tests/golden/controller.npz)."""
import os as _os; _os.environ.setdefault("ACS_ALLOW_RANDOM_CONTROLLER", "1")
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from aircombat_selfplay_b200.controller import make_controller

N = 32768
ctl = make_controller("cuda", allow_random=True)
g = torch.Generator(device="cuda"); g.manual_seed(0)


def timed(fn, reps=50):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps


def rollout(tf32, steps=20):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    gg = torch.Generator(device="cuda"); gg.manual_seed(1)
    h = torch.zeros(N, 128, device="cuda")
    acts = []
    for _ in range(steps):
        x = torch.randn(N, 12, device="cuda", generator=gg) * 0.5
        a, h = ctl(x, h)
        acts.append(a.clone())
    return torch.stack(acts), h


x = torch.randn(N, 12, device="cuda", generator=g) * 0.5
h0 = torch.zeros(N, 128, device="cuda")
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for _ in range(3): ctl(x, h0)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = ctl(x, h0)
    print("controller forward, %d rows, allow_tf32=%s: %.3f ms (graph replay)" % (N, tf32, timed(gr.replay)))
a32, h32 = rollout(False)
a32b, _ = rollout(False)
atf, htf = rollout(True)
print("fp32 run-to-run arg-max mismatches: %d of %d" % (int((a32 != a32b).sum()), a32.numel()))
for t in (0, 4, 9, 19):
    print("step %2d: TF32 changes %.3f %% of the arg-max classes (rows with any change %.3f %%)" % (
        t, 100.0 * float((a32[t] != atf[t]).float().mean()), 100.0 * float((a32[t] != atf[t]).any(dim=-1).float().mean())))
print("max |h_fp32 - h_tf32| after 20 recurrent steps: %.2e" % float((h32 - htf).abs().max()))
