"""Rule-based opponents of the ``use_baseline`` scenarios, batched.

The reference's ``*_vs_pursue`` / ``*_vs_maneuver`` yamls replace the enemy team's RL action by a scripted agent
(reference envs/JSBSim/model/baseline.py): the agent turns the tactical situation into a (delta altitude, delta heading,
delta velocity) request, normalises it into the 12-vector of the low-level controller, and lets the same ``BaselineActor``
GRU produce stick / throttle classes (its own recurrent state, reset with the episode).  The reference does this per
aircraft per step in Python with batch 1; here the request is computed for every red aircraft of every environment with a
few tensor ops on views of the simulator's device state, and the controller call is shared with the hierarchical RL
agents' (one batched call for all rows).

Functions take plain tensors so that the CPU tests can feed them from the oracle env.
"""
from __future__ import annotations

import math

import torch

# ManeuverAgent('triangle') schedule (baseline.py:116-131): heading offsets every 30 s, altitude 6000 m, u 243 m/s
TRIANGLE = (math.pi / 3, math.pi, -math.pi / 3)
TURN_INTERVAL_S = 30.0


def in_range_rad(a: torch.Tensor) -> torch.Tensor:
    """utils.in_range_rad: (-pi, pi] (reference envs/JSBSim/utils/utils.py:112-117; python % is floor-mod)."""
    a = torch.remainder(a, 2 * math.pi)
    return torch.where(a > math.pi, a - 2 * math.pi, a)


def pursue_request(ego_pos, ego_vel, ego_u, tgt_pos, tgt_u):
    """PursueAgent.set_delta_value (baseline.py:80-100).  pos = (north, east, up) [m], vel = (v_n, v_e, v_d) [m/s],
    u = body-x speed [m/s]; all [N, ...] float64.  Returns (delta_altitude [m], delta_heading [rad], delta_velocity [m/s])."""
    delta_altitude = tgt_pos[:, 2] - ego_pos[:, 2]
    ego_v = torch.sqrt(ego_vel[:, 0] ** 2 + ego_vel[:, 1] ** 2)
    dx, dy = tgt_pos[:, 0] - ego_pos[:, 0], tgt_pos[:, 1] - ego_pos[:, 1]
    R = torch.sqrt(dx ** 2 + dy ** 2)
    proj = dx * ego_vel[:, 0] + dy * ego_vel[:, 1]
    ao = torch.acos(torch.clamp(proj / (R * ego_v + 1e-8), -1.0, 1.0))
    side = torch.sign(ego_vel[:, 0] * dy - ego_vel[:, 1] * dx)
    return delta_altitude, ao * side, tgt_u - ego_u


def maneuver_request(step, init_heading, cur_heading, altitude_m, u_mps, time_interval):
    """ManeuverAgent('triangle').set_delta_value with dodge_missile = False (baseline.py:133-151).  ``step`` [N] int64
    counts this agent's calls since the episode reset; ``init_heading`` is the heading at its first call."""
    # step_list = arange(1, 301) * 30 / time_interval; i = first index with step <= step_list[i]; the list is 300 long,
    # a python for-loop that never breaks leaves i at the last index
    per = TURN_INTERVAL_S / time_interval
    i = torch.clamp(torch.ceil(step.to(torch.float64) / per).to(torch.int64) - 1, 0, 299) % 3
    # TRIANGLE[i] without a lookup tensor (creating one here would be a host->device copy inside a CUDA-graph capture)
    offset = torch.where(i == 0, TRIANGLE[0], torch.where(i == 1, TRIANGLE[1], TRIANGLE[2])).to(torch.float64)
    delta_heading = init_heading + offset - cur_heading
    return 6000.0 - altitude_m, delta_heading, 243.0 - u_mps


def controller_input(delta_altitude, delta_heading, delta_velocity, ego9):
    """BaselineAgent.get_observation (baseline.py:44-62): [dh/1000, in_range_rad(dpsi), dv/340, ego9...] -> float32 [N, 12];
    ``ego9`` = (alt/5000, sin roll, cos roll, sin pitch, cos pitch, u/340, v/340, w/340, vc/340)."""
    head = torch.stack([delta_altitude / 1000.0, in_range_rad(delta_heading), delta_velocity / 340.0], dim=-1)
    return torch.cat([head, ego9], dim=-1).to(torch.float32)


class DeviceState:
    """Zero-copy views of the simulator's aircraft arenas, shaped [n_envs, n_agents] per field."""

    def __init__(self, batch):
        self.B, self.A = batch.n_envs, batch.n_agents
        names_d, t_d = batch.arena_view("ac_d")
        names_o, t_o = batch.arena_view("out")
        self._d = {n: t_d[k].view(self.B, self.A) for k, n in enumerate(names_d)}
        self._o = {n: t_o[k].view(self.B, self.A) for k, n in enumerate(names_o)}

    def pos(self, idx):
        return torch.stack([self._d["pos_n"][:, idx], self._d["pos_e"][:, idx], self._d["pos_u"][:, idx]], dim=-1)

    def vel(self, idx):
        return torch.stack([self._d["vel_n"][:, idx], self._d["vel_e"][:, idx], self._d["vel_d"][:, idx]], dim=-1)

    def u(self, idx):
        return self._d["u_mps"][:, idx]

    def heading(self, idx):
        return self._o["heading_rad"][:, idx]

    def altitude(self, idx):
        return self._d["h_sl_m"][:, idx]

    def ego9(self, idx):
        roll, pitch = self._o["roll_rad"][:, idx], self._o["pitch_rad"][:, idx]
        d = self._d
        return torch.stack([d["h_sl_m"][:, idx] / 5000, torch.sin(roll), torch.cos(roll), torch.sin(pitch), torch.cos(pitch),
                            d["u_mps"][:, idx] / 340, d["v_mps"][:, idx] / 340, d["w_mps"][:, idx] / 340, d["vc_mps"][:, idx] / 340], dim=-1)


class RuleOpponents:
    """The red team's scripted agents of one env batch.  ``kind`` = 'pursue' | 'maneuver' (the yaml ``baseline_type``;
    'loiter' is not handled by the reference's own ``load_agent`` either and raises the same NotImplementedError)."""

    def __init__(self, kind: str, env_kind: str, n_ego: int, n_enm: int, time_interval: float, state, n_envs: int, device):
        """``state``: anything with DeviceState's accessors (the CPU tests pass a view of the oracle envs)."""
        if kind not in ("pursue", "maneuver") or (kind == "maneuver" and env_kind != "1v1"):
            raise NotImplementedError(f"baseline_type {kind!r} (reference load_agent / load_agents raise here too)")
        self.kind, self.n_ego, self.n_enm, self.time_interval = kind, n_ego, n_enm, time_interval
        self.state = state
        self.step = torch.zeros((n_envs, n_enm), dtype=torch.int64, device=device)
        self.init_heading = torch.zeros((n_envs, n_enm), dtype=torch.float64, device=device)
        self.has_init = torch.zeros((n_envs, n_enm), dtype=torch.bool, device=device)

    def reset(self, env_mask=None):
        if env_mask is None:
            self.step.zero_(); self.has_init.zero_()
        else:
            keep = ~env_mask.bool().view(-1, 1)
            self.step.mul_(keep); self.has_init.logical_and_(keep)

    def inputs(self) -> torch.Tensor:
        """float32 [n_envs, n_enm, 12]: the controller input of every red aircraft for this step."""
        st, rows = self.state, []
        for k in range(self.n_enm):
            me = self.n_ego + k
            if self.kind == "pursue":
                # 1v1: PursueAgent.get_action(env, task) chases agents[0]; N-v-N: baseline_agent[k].get_action(env, task, k)
                # chases the k-th aircraft of the agents dict, i.e. blue k (E/tasks/scenario2_task.py:49-53)
                req = pursue_request(st.pos(me), st.vel(me), st.u(me), st.pos(k), st.u(k))
            else:
                cur = st.heading(me)
                self.init_heading[:, k] = torch.where(self.has_init[:, k], self.init_heading[:, k], cur)
                self.has_init[:, k] = True
                req = maneuver_request(self.step[:, k], self.init_heading[:, k], cur, st.altitude(me), st.u(me), self.time_interval)
                self.step[:, k] += 1
            rows.append(controller_input(*req, st.ego9(me)))
        return torch.stack(rows, dim=1)
