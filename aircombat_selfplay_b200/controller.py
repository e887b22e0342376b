"""Batched low-level flight controller for the hierarchical tasks.

The reference's hierarchical tasks turn a high-level action (delta altitude / heading / velocity class) into stick and
throttle commands with a small recurrent policy, ``BaselineActor`` (reference envs/JSBSim/model/baseline_actor.py:88-110):
MLP(12 -> 128 -> 128, ReLU + LayerNorm) -> one GRU layer (128) -> LayerNorm -> four categorical heads [41, 41, 41, 30],
arg-max actions.  The reference evaluates it on the CPU with batch 1, once per agent per step, inside
``normalize_action`` (reference envs/JSBSim/tasks/singlecombat_task.py:223-256).  Here it is ONE PyTorch call over all
``n_envs * n_agents`` rows on the simulator's GPU, between "last step's observation" and "substeps" -- the policy side
stays in PyTorch by design (BASELINE.json north_star); the physics kernels are the hand-written part.

``load_reference_checkpoint`` reads the reference's ``baseline_model.pt`` state dict (key names of ``BaselineActor``).
"""
from __future__ import annotations

import math
import os
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

HIDDEN = 128
HEAD_DIMS = (41, 41, 41, 30)

# reference envs/JSBSim/tasks/singlecombat_task.py:216-218
NORM_DELTA_ALTITUDE = (0.1, 0.0, -0.1)
NORM_DELTA_HEADING = (-math.pi / 6, -math.pi / 12, 0.0, math.pi / 12, math.pi / 6)
NORM_DELTA_VELOCITY = (0.05, 0.0, -0.05)


class LowLevelController(nn.Module):
    def __init__(self, input_dim: int = 12):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, HIDDEN)
        self.ln1 = nn.LayerNorm(HIDDEN)
        self.fc2 = nn.Linear(HIDDEN, HIDDEN)
        self.ln2 = nn.LayerNorm(HIDDEN)
        self.gru = nn.GRUCell(HIDDEN, HIDDEN)      # one time step of the reference's 1-layer nn.GRU (same weight layout)
        self.norm = nn.LayerNorm(HIDDEN)
        self.heads = nn.Linear(HIDDEN, sum(HEAD_DIMS))   # the four logits_net layers stacked row-wise

    def _padded_heads(self):
        """The four logit heads as one [4 * 41, 128] weight: head k occupies rows [41 k, 41 k + HEAD_DIMS[k]), the padding
        rows have zero weight and a -1e30 bias, so ONE arg-max over a [N, 4, 41] view picks the four classes (the
        throttle head has 30 classes) instead of four reduction kernels over slices."""
        w, b = self.heads.weight, self.heads.bias
        key = (w.data_ptr(), w._version, b._version, w.device)
        if getattr(self, "_hp_key", None) != key:
            m = max(HEAD_DIMS)
            wp = torch.zeros((len(HEAD_DIMS) * m, HIDDEN), dtype=w.dtype, device=w.device)
            bp = torch.full((len(HEAD_DIMS) * m,), -1e30, dtype=b.dtype, device=b.device)
            o = 0
            for k, d in enumerate(HEAD_DIMS):
                wp[k * m:k * m + d] = w[o:o + d]
                bp[k * m:k * m + d] = b[o:o + d]
                o += d
            self._hp, self._hp_key = (wp, bp), key
        return self._hp

    @torch.no_grad()
    def forward(self, obs12: torch.Tensor, h: torch.Tensor):
        """obs12 [N, 12] float32, h [N, 128] float32 -> (actions [N, 4] int32, a transposed view of a [4, N] tensor; h' [N, 128])."""
        x = self.ln1(F.relu_(self.fc1(obs12)))
        x = self.ln2(F.relu_(self.fc2(x)))
        h = self.gru(x, h)
        wp, bp = self._padded_heads()
        # logits^T = Wp hn^T + bp: the same dot products as the four logits_net layers, laid out [4 * 41, N] so that the
        # arg-max reduces over a strided dimension with N contiguous (coalesced; the [N, 4, 41] layout reduced 41 adjacent
        # floats per thread group: 55 us of a 0.56 ms controller call at 32 768 rows)
        logits_t = torch.addmm(bp.unsqueeze(1), wp, self.norm(h).t())
        # Categorical(logits).probs.argmax == logits.argmax (first maximum on ties, as torch.argmax)
        return logits_t.view(len(HEAD_DIMS), max(HEAD_DIMS), -1).argmax(dim=1).to(torch.int32).t(), h

    def load_reference_state_dict(self, sd: dict):
        """Maps ``BaselineActor.state_dict()`` keys (reference envs/JSBSim/model/baseline_actor.py) onto this module."""
        m = {
            "fc1.weight": sd["base.mlp.fc.0.weight"], "fc1.bias": sd["base.mlp.fc.0.bias"],
            "ln1.weight": sd["base.mlp.fc.2.weight"], "ln1.bias": sd["base.mlp.fc.2.bias"],
            "fc2.weight": sd["base.mlp.fc.3.weight"], "fc2.bias": sd["base.mlp.fc.3.bias"],
            "ln2.weight": sd["base.mlp.fc.5.weight"], "ln2.bias": sd["base.mlp.fc.5.bias"],
            "gru.weight_ih": sd["rnn.gru.weight_ih_l0"], "gru.weight_hh": sd["rnn.gru.weight_hh_l0"],
            "gru.bias_ih": sd["rnn.gru.bias_ih_l0"], "gru.bias_hh": sd["rnn.gru.bias_hh_l0"],
            "norm.weight": sd["rnn.norm.weight"], "norm.bias": sd["rnn.norm.bias"],
            "heads.weight": torch.cat([sd[f"act.action_outs.{k}.logits_net.weight"] for k in range(4)], dim=0),
            "heads.bias": torch.cat([sd[f"act.action_outs.{k}.logits_net.bias"] for k in range(4)], dim=0),
        }
        self.load_state_dict(m)


def find_checkpoint(path=None, config_dir=None):
    """``baseline_model.pt`` lookup (the reference loads ``<envs/JSBSim>/model/baseline_model.pt`` and raises when it is
    absent, envs/JSBSim/tasks/singlecombat_task.py:208-213): an explicit ``path`` or ``$ACS_BASELINE_MODEL`` MUST exist;
    otherwise ``<package>/model/baseline_model.pt``, then -- when the yaml files come from a reference checkout
    (``config_dir=<...>/envs/JSBSim/configs``) -- that checkout's ``model/baseline_model.pt``.  None when nothing is found."""
    for what, c in (("controller_path", path), ("$ACS_BASELINE_MODEL", os.environ.get("ACS_BASELINE_MODEL"))):
        if c:
            if not Path(c).exists():
                raise FileNotFoundError(f"{what} = {c!r} does not exist (low-level controller checkpoint baseline_model.pt)")
            return Path(c)
    cands = [Path(__file__).resolve().parent / "model" / "baseline_model.pt"]
    if config_dir:
        cands.append(Path(config_dir).resolve().parent / "model" / "baseline_model.pt")
    for c in cands:
        if c.exists():
            return c
    return None


def make_controller(device, path=None, seed: int = 0, config_dir=None, allow_random: bool = False) -> LowLevelController:
    """The controller on ``device`` with the weights of the reference's ``baseline_model.pt``.

    The checkpoint is not part of this repository.  Without one this RAISES, as the reference's ``torch.load`` does --
    every hierarchical task (scenario1/2/3, wvr, maneuver_curriculum, Hierarchy* yamls, the scripted opponents) would
    otherwise fly a random low-level policy without saying so.  ``allow_random=True`` (or ``ACS_ALLOW_RANDOM_CONTROLLER=1``)
    is the explicit opt-in for benchmarks and shape tests: a seeded random init of the same architecture, i.e. the same
    arithmetic and memory traffic, not the trained flight behaviour."""
    ck = find_checkpoint(path, config_dir)
    g = torch.Generator().manual_seed(seed)
    ctl = LowLevelController()
    if ck is not None:
        ctl.load_reference_state_dict(torch.load(str(ck), map_location="cpu"))
        ctl.checkpoint = str(ck)
    else:
        if not (allow_random or os.environ.get("ACS_ALLOW_RANDOM_CONTROLLER") == "1"):
            from .capi import AcsError
            raise AcsError("baseline_model.pt (the low-level controller of the hierarchical tasks) was not found: pass "
                           "controller_path=, set $ACS_BASELINE_MODEL, copy it to aircombat_selfplay_b200/model/, or point "
                           "config_dir at a reference checkout's envs/JSBSim/configs; allow_random_controller=True opts in to "
                           "random-init weights (benchmarks / shape tests only)")
        for name, p in ctl.named_parameters():      # every parameter from the seeded generator: instances are identical
            if p.dim() > 1:
                p.data = torch.randn(p.shape, generator=g) / math.sqrt(p.shape[-1])
            elif name.startswith(("ln", "norm")) and name.endswith("weight"):
                p.data = torch.ones_like(p)
            else:
                p.data = 0.1 * torch.randn(p.shape, generator=g)
        ctl.checkpoint = None
    return ctl.to(device).eval()


def hierarchical_luts(device):
    """The three class -> delta lookup tables as ONE device tensor of 11 entries plus the per-column offsets (created once:
    building them inside a step would be a pageable H2D copy, which a CUDA-graph capture forbids)."""
    lut = torch.tensor(NORM_DELTA_ALTITUDE + NORM_DELTA_HEADING + NORM_DELTA_VELOCITY, dtype=torch.float64, device=device)
    off = torch.tensor([0, len(NORM_DELTA_ALTITUDE), len(NORM_DELTA_ALTITUDE) + len(NORM_DELTA_HEADING)], dtype=torch.int64, device=device)
    return lut, off


def hierarchical_input(high: torch.Tensor, obs: torch.Tensor, force_climb_below_m=None, luts=None) -> torch.Tensor:
    """input_obs of the reference (singlecombat_task.py:234-246 / multiplecombat_task.py:171-178).

    high [N, 3] integer classes; obs [N, D] float64 current observations (obs[:, 0] = altitude / 5000).  The 1v1 task
    forces the "climb" class below 3500 m (singlecombat_task.py:235-237).  One gather for the three deltas, one
    concatenation, one cast."""
    lut, off = luts if luts is not None else hierarchical_luts(obs.device)
    idx = high.to(torch.int64) + off
    if force_climb_below_m is not None:
        idx[:, 0] = torch.where(obs[:, 0] * 5000.0 < force_climb_below_m, 0, idx[:, 0])
    return torch.cat([lut[idx], obs[:, :9]], dim=-1).to(torch.float32)
