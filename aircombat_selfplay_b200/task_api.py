"""The Task plugin surface of the reference (reference envs/JSBSim/tasks/task_base.py:8-122) over the device simulator.

In the reference a Task is a Python object whose methods the env calls once per agent per step: ``normalize_action``,
``step`` (weapons), ``get_obs``, ``get_reward`` (a list of reward-function objects), ``get_termination`` (a list of
termination-condition objects).  Here those computations run inside ``k_env_substeps`` / ``k_env_post`` for every env at
once (a Python callback cannot run in a kernel), driven by the ``AcsTaskConfig`` the same yaml resolves to.  ``env.task``
keeps the reference's *surface* -- same attribute and method names, argument order and return shapes -- as a VIEW of
what the device computed for the last step:

    task.num_agents, task.observation_space, task.action_space, task.share_observation_space
    task.reward_functions          [RewardFunction]      class name, reward_scale, is_potential, hyper-parameters, order
    task.termination_conditions    [TerminationCondition] class name, order; get_termination(task, env, agent_id, info)
    task.get_obs(env, agent_id)                 the agent's row of the current observation
    task.get_reward(env, agent_id, info)        (last step's reward of the agent, info)
    task.get_termination(env, agent_id, info)   (last step's done flag, info + done_condition)
    task.normalize_action(env, agent_id, action) the four normalised commands the kernel applies for this action (a pure
                                                function here: the hierarchical controller's recurrent state is NOT advanced)
    task.reset(env) / task.step(env)            no-ops: the device step already did both
    task._check_missile_warning(env, agent_id)  first live incoming missile (position, velocity) or None

``agent_id`` is the reference's uid string ("A0100") or the agent's index.  For a batched env the views return one row
per env.  Adding a NEW reward / termination / observation packer means a new ``ACS_R_*`` / ``ACS_T_*`` / ``ACS_OBS_*``
enum value in include/acs.h, its case in ``reward_one`` / ``agent_termination`` / ``write_obs``
(csrc/env_kernels.cuh), its scalar restatement in oracle/env_oracle.py and the class-name entry in taskspec.py /
tasks.py -- INTEGRATION.md walks through it.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import taskspec as ts

REWARD_CLASS_NAME = {v: k for k, v in ts.REWARD_CLASS.items()}
TERMINATION_CLASS_NAME = {ts.T_UNREACH_HEADING: "UnreachHeading", ts.T_EXTREME_STATE: "ExtremeState", ts.T_OVERLOAD: "Overload",
                          ts.T_LOW_ALTITUDE: "LowAltitude", ts.T_TIMEOUT: "Timeout", ts.T_SAFE_RETURN: "SafeReturn"}
_REWARD_PARAMS = {ts.R_ALTITUDE: ("safe_altitude", "danger_altitude", "Kv"),
                  ts.R_POSTURE: ("orientation_version", "range_version", "target_dist"), ts.R_RELATIVE_ALTITUDE: ("KH",)}


class RewardFunction:
    """Descriptor of one reward class of the task, in evaluation order (reference envs/JSBSim/reward_functions/
    reward_function_base.py:8-76).  The value is computed in ``reward_one`` (csrc/env_kernels.cuh)."""

    def __init__(self, task: "Task", index: int, spec: ts.RewardSpec):
        self._task, self.index, self.kind = task, index, spec.kind
        self.name = REWARD_CLASS_NAME[spec.kind]
        self.reward_scale = spec.scale
        self.is_potential = bool(spec.potential)
        vals = (spec.p0, spec.p1, spec.p2)
        self.params = {k: (f"v{int(v)}" if k.endswith("_version") else v) for k, v in zip(_REWARD_PARAMS.get(spec.kind, ()), vals)}
        self.reward_item_names = [self.name]

    def pre_rewards(self, env=None):
        """``BaseRewardFunction.pre_rewards`` of a potential-based reward (reward_function_base.py:56): [n_envs, A] tensor."""
        core = self._task._core
        names, t = core.batch.arena("ac_d")
        return t[names.index(f"pre_reward{self.index}")].view(core.n_envs, core.n_agents)

    def reset(self, task, env):        # done by the device reset
        pass

    def get_reward(self, task, env, agent_id):
        raise NotImplementedError(f"{self.name}.get_reward: per-class values are summed inside k_env_post; task.get_reward "
                                  "returns the agent's total, pre_rewards() the potential memory")

    def __repr__(self):
        return f"{self.name}(scale={self.reward_scale}, potential={self.is_potential}, {self.params})"


class TerminationCondition:
    """Descriptor of one termination class of the task, in evaluation order (reference envs/JSBSim/termination_conditions/
    termination_condition_base.py).  ``get_termination`` reports what the device decided for the last step."""

    def __init__(self, task: "Task", index: int, kind: int):
        self._task, self.index, self.kind = task, index, kind
        self.name = TERMINATION_CLASS_NAME[kind]

    def get_termination(self, task, env, agent_id, info: Optional[dict] = None):
        """(done, success, info): done iff THIS condition ended the agent's episode in the last step (conditions are
        evaluated in order and the first hit short-circuits, task_base.py:104-110)."""
        info = {} if info is None else info
        row = self._task._info_row(agent_id)
        cause = row[..., 0]
        done = cause == self.kind
        # the reference's aggregate `success` survives only when the agent ends ALIVE through SafeReturn (safe_return.py:41-47:
        # every enemy down, no missile inbound); every other ending reports success False (timeout.py:31, low_altitude.py:33 ...)
        success = done & (self.kind == ts.T_SAFE_RETURN) & (row[..., 1] == 0)
        if self._task._single:
            done, success = bool(done), bool(success)
            if done:
                from .envs import DONE_CONDITIONS
                info["done_condition"] = DONE_CONDITIONS[self.kind]
        return done, success, info

    def __repr__(self):
        return self.name


class Task:
    """``env.task`` -- see the module docstring."""

    def __init__(self, core, single: bool):
        self._core, self._single = core, single
        self.config = core.config
        self.name = core.task_name
        sp = core.spec
        self.reward_functions: List[RewardFunction] = [RewardFunction(self, i, r) for i, r in enumerate(sp.rewards)]
        self.termination_conditions: List[TerminationCondition] = [TerminationCondition(self, i, t) for i, t in enumerate(sp.terminations)]
        self.observation_space = core.observation_space
        self.share_observation_space = core.share_observation_space
        self.action_space = core.action_space
        self.use_baseline = bool(sp.use_baseline)
        self.use_artillery = bool(sp.use_artillery)
        self.max_attack_angle, self.max_attack_distance = sp.max_attack_angle, sp.max_attack_distance
        self.min_attack_interval = sp.min_attack_interval
        # the registry entry this task resolves to (obs packer, launch rule, hierarchical flag ...), tasks.TASKS[name]
        self.descriptor: Dict = dict(core.task_desc)

    # mapping access keeps older call sites (`env.task["hier"]`) working
    def __getitem__(self, k):
        return self.descriptor[k]

    @property
    def num_agents(self) -> int:
        return self._core.n_agents

    # ------------------------------------------------------------------ helpers
    def _index(self, agent_id) -> int:
        if isinstance(agent_id, str):
            return (self._core.ego_ids + self._core.enm_ids).index(agent_id)
        return int(agent_id)

    def _pick(self, t: torch.Tensor, agent_id):
        x = t[:, self._index(agent_id)]
        return x[0] if self._single else x

    def _info_row(self, agent_id):
        return self._pick(self._core.batch.info, agent_id).cpu().numpy()

    # ------------------------------------------------------------------ the reference's methods
    def reset(self, env=None):
        """task.reset(env): reward-function resets, weapon counters, lock windows -- all part of the device reset."""

    def step(self, env=None):
        """task.step(env): weapon launches / artillery -- part of the device step (task_step_agent, csrc/env_kernels.cuh)."""

    def get_obs(self, env, agent_id) -> np.ndarray:
        """The agent's current observation ([D]; batched env: [n_envs, D])."""
        return self._pick(self._core.batch.obs, agent_id).cpu().numpy().copy()

    def get_reward(self, env, agent_id, info: Optional[dict] = None):
        info = {} if info is None else info
        r = self._pick(self._core.batch.rewards, agent_id).cpu().numpy()
        return (float(r) if self._single else r.copy()), info

    def get_termination(self, env, agent_id, info: Optional[dict] = None):
        info = {} if info is None else info
        d = self._pick(self._core.batch.dones, agent_id).cpu().numpy().astype(bool)
        if self._single:
            d = bool(d)
            cause = int(self._info_row(agent_id)[0])
            if d and cause >= 0:
                from .envs import DONE_CONDITIONS
                info["done_condition"] = DONE_CONDITIONS[cause]
        return d, info

    def normalize_action(self, env, agent_id, action) -> np.ndarray:
        """The four normalised commands (aileron, elevator, rudder in [-1, 1], throttle in [0.4, 0.9]) the step applies for
        ``action`` -- the host mirror of ``decode_action`` (csrc/env_kernels.cuh; reference heading_task.py:102-110,
        singlecombat_task.py:141-153).  Hierarchical tasks: the low-level controller is evaluated on the agent's current
        observation and recurrent state WITHOUT advancing that state (the reference's call has that side effect; the
        state that counts advances inside ``env.step``)."""
        core = self._core
        a = np.asarray(action).ravel()
        k = self._index(agent_id)
        if core.hier:
            from .controller import hierarchical_input
            high = torch.as_tensor(a[:3].astype(np.int64), device=core.device).view(1, 3)
            obs = core.batch.obs[0:1, k]
            x = hierarchical_input(high, obs, core._climb_below, core._luts)
            low, _ = core.controller(x, core.rnn.view(core.n_envs, core.n_agents, -1)[0:1, k].clone())
            a = low[0].cpu().numpy()
        a = a[:4].astype(np.float64)
        out = np.zeros(4)
        if core.spec.act_kind == ts.ACT_HEADING:
            out[0:3] = a[0:3] * 2. / (41 - 1.) - 1.
            out[3] = a[3] * 0.5 / (30 - 1.) + 0.4
        else:
            out[0:3] = a[0:3] / 20 - 1.
            out[3] = a[3] / 58 + 0.4
        return np.clip(out, [-1.0, -1.0, -1.0, 0.0], [1.0, 1.0, 1.0, 0.9])       # the catalog clip on set (catalog.py:192-197)

    def _check_missile_warning(self, env, agent_id):
        """First live missile in the agent's ``under_missiles`` (append order), as ``AircraftSimulator.check_missile_warning``
        (simulatior.py:321-325): ``{"position": [3], "velocity": [3], "shooter": index}`` of env 0, or None."""
        core = self._core
        k = self._index(agent_id)
        ni, mi = core.batch.arena("ms_i")
        nd, md = core.batch.arena("ms_d")
        A, S = core.n_agents, max(1, core.spec.n_missile_slots)
        mi, md = mi.cpu().numpy(), md.cpu().numpy()
        best = None
        for j in range(A):
            if (j < core.spec.n_ego) == (k < core.spec.n_ego):
                continue
            for s in range(S):
                c = j * S + s                                   # env 0
                if mi[ni.index("status"), c] == 0 and mi[ni.index("target"), c] == k and not mi[ni.index("detached"), c]:
                    born = mi[ni.index("born"), c]
                    if best is None or born < best[0]:
                        best = (born, c, j)
        if best is None:
            return None
        _, c, j = best
        return {"position": np.array([md[nd.index(f)][c] for f in ("pos_n", "pos_e", "pos_u")]),
                "velocity": np.array([md[nd.index(f)][c] for f in ("vel_n", "vel_e", "vel_u")]), "shooter": j}

    def __repr__(self):
        return f"Task({self.name!r}, rewards={self.reward_functions}, terminations={self.termination_conditions})"
