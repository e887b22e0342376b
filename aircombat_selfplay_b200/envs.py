"""Host-side mirror of the reference's environment classes over the batched device simulator.

``BatchedEnv``        n_envs environments of one scenario on one GPU; torch tensors in / out; auto-reset on device.
``SingleControlEnv`` / ``SingleCombatEnv`` / ``MultipleCombatEnv``
                      the reference's per-env classes (reference envs/JSBSim/envs/*.py): same constructor argument
                      (yaml config name), same gym-style ``reset``/``step`` shapes, ``seed``, ``agents``, ``ego_ids`` /
                      ``enm_ids``, ``num_agents``, ``observation_space`` / ``action_space`` (/ ``share_observation_space``),
                      ``time_interval``, ``current_step`` -- each is a batch of ONE env on the GPU.
The vectorised contract the runners consume lives in ``env_wrappers.py``.

There is no CPU fallback: constructing any of these without a CUDA device or without the built ``libacs.so`` raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import spaces
from . import taskspec as ts
from .capi import INFO_DIM, AcsError, EnvBatch
from .controller import hierarchical_input, hierarchical_luts, make_controller
from .tasks import TASKS, load_spec, parse_config

# done_condition strings of the reference's termination classes (without colorama codes), by ACS_T_* id
DONE_CONDITIONS = {
    ts.T_UNREACH_HEADING: "unreach_heading", ts.T_EXTREME_STATE: "extreme_state", ts.T_OVERLOAD: "overload",
    ts.T_LOW_ALTITUDE: "low_altitude", ts.T_TIMEOUT: "timeout", ts.T_SAFE_RETURN: "safe_return",
}


def action_space_for(task_name: str):
    """reference load_action_space of each Task class (E/tasks/*.py)."""
    t = TASKS[task_name]
    if task_name == "hierarchical_multiplecombat_shoot_nearest":      # E/tasks/multiplecombat_task.py:217-218
        return spaces.MultiDiscrete([3, 5, 3, 2])
    low = spaces.MultiDiscrete([3, 5, 3]) if t["hier"] else spaces.MultiDiscrete([41, 41, 41, 30])
    if t["shoot"] == 0:
        return low
    if t["shoot"] == 1:
        return spaces.Tuple([low, spaces.Discrete(2)])
    return spaces.Tuple([low, spaces.MultiDiscrete([2, 2, 2, 2])])


class BatchedEnv:
    """``n_envs`` environments of one yaml scenario on one GPU.

    ``step(actions)`` takes the task's own action layout -- int tensor ``[n_envs, n_agents, act_dim]`` where ``act_dim``
    is 4 (direct stick/throttle classes) or 3 (hierarchical classes) plus the shoot entries -- runs the batched low-level
    controller for hierarchical tasks, then the device step with auto-reset.  Outputs are views of one device buffer
    (``self.batch.out_buf``) that the next step overwrites.
    """

    def __init__(self, config_name: str, n_envs: int, device: int = 0, seed: int = 0, env_offset: int = 0,
                 config_dir: Optional[str] = None, substeps: Optional[int] = None, controller_path: Optional[str] = None,
                 auto_reset: bool = True, use_cuda_graph: Optional[bool] = None, allow_random_controller: bool = False,
                 curriculum_rule: Optional[int] = None, curriculum_threshold: Optional[float] = None,
                 curriculum_window: Optional[int] = None):
        if not torch.cuda.is_available():
            raise AcsError("CUDA is not available; the simulator has no CPU fallback")
        self.config_name = config_name
        self.config = parse_config(config_name, config_dir)
        self.spec = load_spec(config_name, config_dir, substeps)
        self.task_name = self.spec.name
        self.task_desc = TASKS[self.task_name]
        for k, x in (("curriculum_rule", curriculum_rule), ("curriculum_threshold", curriculum_threshold), ("curriculum_window", curriculum_window)):
            if x is not None:          # overrides of the per-env win-rate record / stage rule (set_curriculum_stages)
                setattr(self.spec, k, type(getattr(self.spec, k))(x))
        if self.spec.use_baseline and not self.task_desc["hier"]:
            raise NotImplementedError("use_baseline is wired for the hierarchical task families this fork ships (scenario1/2/3, "
                                      "wvr, maneuver_curriculum); the LAG-style SingleCombatTask path calls "
                                      "baseline_agent.get_action with a signature the fork's agents no longer have")
        self.n_envs, self.n_agents = n_envs, self.spec.n_agents
        self.device = torch.device("cuda", device)
        self.auto_reset = auto_reset
        self.seed_value = int(seed)
        with torch.cuda.device(self.device):
            self.batch = EnvBatch(self.spec, n_envs, seed=seed, device=device, env_offset=env_offset)
        self.hier = bool(self.task_desc["hier"])
        self.high_dim = 3 if self.hier else 4
        self.act_dim = self.high_dim + self.spec.shoot_dim
        self.observation_space = spaces.Box(low=-10, high=10.0, shape=(self.spec.obs_dim,))
        self.share_observation_space = spaces.Box(low=-10, high=10.0, shape=(self.n_agents * self.spec.obs_dim,))
        self.action_space = action_space_for(self.task_name)
        uids = list(self.config["aircraft_configs"].keys())
        self.ego_ids = [u for u in uids if u[0] == uids[0][0]]
        self.enm_ids = [u for u in uids if u[0] != uids[0][0]]
        self._low = torch.zeros((n_envs, self.n_agents, 4 + self.spec.shoot_dim), dtype=torch.int32, device=self.device)
        if self.hier:
            self.controller = make_controller(self.device, controller_path, config_dir=config_dir, allow_random=allow_random_controller)
            self._luts = hierarchical_luts(self.device)
            self.rnn = torch.zeros((n_envs * self.n_agents, 128), dtype=torch.float32, device=self.device)
            # 1v1 hierarchical tasks force a climb below 3500 m (E/tasks/singlecombat_task.py:235-237)
            self._climb_below = 3500.0 if self.task_desc["env"] == "1v1" else None
        self.opponents = None
        if self.spec.use_baseline:
            from .opponents import DeviceState, RuleOpponents
            self.opponents = RuleOpponents(self.spec.baseline_type, self.task_desc["env"], self.spec.n_ego, self.spec.n_enm,
                                           self.spec.substeps / self.spec.sim_freq, DeviceState(self.batch), n_envs, self.device)
        self._was_reset = False
        # hierarchical tasks: the whole step -- controller (~40 small PyTorch kernels) + the env kernels -- is captured once
        # into a CUDA graph and replayed, one launch per step instead of a launch-bound train of tiny kernels.  Plain
        # tasks launch their four kernels directly (a graph launch costs more than it saves there: measured +20%).
        self.use_cuda_graph = self.hier if use_cuda_graph is None else bool(use_cuda_graph)
        self._graph = None
        self._epoch = 0             # bumped whenever a captured step goes stale (host wrappers hold graphs of their own)
        self._act_in = torch.zeros((n_envs, self.n_agents, self.act_dim), dtype=torch.int32, device=self.device)
        self._timing = False
        self.curriculum_angle = 0
        from .task_api import Task
        self.task = Task(self, single=False)       # the reference's Task plugin surface as a view of the device state

    # ------------------------------------------------------------------ reference-shaped properties
    @property
    def num_agents(self) -> int:
        return self.n_agents

    @property
    def time_interval(self) -> float:
        return self.spec.substeps / self.spec.sim_freq

    @property
    def max_steps(self) -> int:
        return self.spec.max_steps

    def seed(self, seed: int = 0):
        self.seed_value = int(seed or 0)
        self.batch.set_seed(self.seed_value)
        self._graph = None          # the seed is a kernel parameter baked into the captured step
        self._epoch += 1
        return [seed]

    def set_init_states(self, init_states):
        """Per-aircraft initial conditions used by the following resets (reference reset_simulators /
        reset_simulators_curriculum, E/envs/singlecombat_env.py:45-122)."""
        self.batch.set_init_states(init_states)
        for k, row in enumerate(init_states):        # keep the host copy of the config in step (state_dict reads it)
            for j in range(12):
                self.batch.cfg.init_state[k][j] = float(row[j])
        self._graph = None
        self._epoch += 1

    def set_curriculum_angle(self, angle: int):
        """Curriculum stage of the *_curriculum / wvr / maneuver_curriculum tasks: the following resets start from
        ``reset_simulators_curriculum(angle)`` (E/envs/singlecombat_env.py:87-122, multiplecombat_env.py:185-248).  The
        reference keeps one stage counter per env process, advanced by its own win-rate record; here the stage is
        batch-wide and advanced by the caller (the per-step ``info`` carries what a win-rate record needs)."""
        from .tasks import curriculum_init_states
        if not self.spec.curriculum:
            raise AcsError(f"task {self.task_name!r} has no curriculum reset")
        self.curriculum_angle = int(angle)
        self.set_init_states(curriculum_init_states(self.spec.env_kind, self.spec.yaml_init_states, self.curriculum_angle))

    def set_curriculum_stages(self, angles):
        """Per-env curriculum, as the reference keeps it per env process (``curriculum_angle``, ``record``, ``winning_rate`` of
        Scenario*_curriculum / WVRTask, E/tasks/scenario2_task.py:172-223): stage ``k`` resets through
        ``reset_simulators_curriculum(angles[k])``.  Every env carries its own stage and its own record of the last
        ``window`` episode outcomes in the ``env_i`` arena ("stage", "curriculum_record", "curriculum_count"); the device
        updates the record when an episode ends and applies the rule before the auto-reset that follows.
        The rule is a constructor argument (``curriculum_rule``): 1 = the reference's rule verbatim (``rate >= threshold and
        len(record) > 20`` -- never true, the record is capped at 20 entries, so the reference's stage never moves), 2 =
        advance when the record is full, 0 = record only; default: the task's (1).  Stages can also be written directly
        (``set_env_stages``)."""
        from .tasks import curriculum_init_states
        if not self.spec.curriculum:
            raise AcsError(f"task {self.task_name!r} has no curriculum reset")
        angles = [int(a) for a in angles]
        self.curriculum_angles = angles
        self.curriculum_angle = angles[0]
        self.set_init_states(curriculum_init_states(self.spec.env_kind, self.spec.yaml_init_states, angles[0]))
        for k, a in enumerate(angles[1:], start=1):
            self.batch.set_stage_init_states(k, curriculum_init_states(self.spec.env_kind, self.spec.yaml_init_states, a))

    def set_env_stages(self, stages: torch.Tensor):
        """Writes every env's curriculum stage (int tensor [n_envs]); it takes effect at the env's next reset."""
        names, ei = self.batch.arena("env_i")
        ei[names.index("stage")] = stages.to(device=self.device, dtype=torch.int32)
        self.batch.set_arena("env_i", ei)

    def curriculum_state(self):
        """(stage, wins in the record, episodes in the record) per env, as device tensors."""
        names, ei = self.batch.arena("env_i")
        bits = ei[names.index("curriculum_record")]
        wins = torch.zeros_like(bits)
        for k in range(31):
            wins += (bits >> k) & 1
        return ei[names.index("stage")], wins, ei[names.index("curriculum_count")]

    def close(self):
        self.batch.close()

    # ------------------------------------------------------------------ checkpoint / resume (the reference never saves env state)
    def state_dict(self):
        """Everything a bit-exact resume needs: the simulator's arenas, the last step's outputs (the hierarchical
        controller reads the observation), the controller's recurrent state, the scripted opponents' counters."""
        from .capi import ARENAS
        sd = {f"arena:{k}": self.batch.arena(k)[1] for k in ARENAS}
        sd["out_buf"] = self.batch.out_buf.clone()
        if self.hier:
            sd["rnn"] = self.rnn.clone()
        if self.opponents is not None:
            sd.update({"opp:step": self.opponents.step.clone(), "opp:init_heading": self.opponents.init_heading.clone(),
                       "opp:has_init": self.opponents.has_init.clone()})
        sd["seed"] = self.seed_value
        # what the NEXT auto-resets start from: the per-aircraft initial conditions in force (set_init_states / curriculum stage)
        sd["init_states"] = torch.tensor([[float(x) for x in self.batch.cfg.init_state[k]] for k in range(self.n_agents)], dtype=torch.float64)
        sd["curriculum_angle"] = int(self.curriculum_angle)
        sd["curriculum_angles"] = list(getattr(self, "curriculum_angles", []))     # per-env stages: the stage templates to rebuild
        return sd

    def load_state_dict(self, sd):
        from .capi import ARENAS
        if int(sd["seed"]) != self.seed_value:
            self.seed(int(sd["seed"]))            # NB: seed() restarts the episode counters; the arenas restore them below
        if "init_states" in sd:                   # rebuilds the reset template(s) before the arenas are restored
            if sd.get("curriculum_angles"):
                self.set_curriculum_stages(sd["curriculum_angles"])
            else:
                self.set_init_states(sd["init_states"].cpu().numpy())
            self.curriculum_angle = int(sd.get("curriculum_angle", 0))
        for k in ARENAS:
            self.batch.set_arena(k, sd[f"arena:{k}"].to(self.device).contiguous())
        self.batch.out_buf.copy_(sd["out_buf"])
        if self.hier:
            self.rnn.copy_(sd["rnn"])
        if self.opponents is not None:
            self.opponents.step.copy_(sd["opp:step"]); self.opponents.init_heading.copy_(sd["opp:init_heading"])
            self.opponents.has_init.copy_(sd["opp:has_init"])
        self._was_reset = True

    # ------------------------------------------------------------------ device API
    def reset(self, env_mask: Optional[torch.Tensor] = None):
        with torch.cuda.device(self.device):
            if self.hier:
                if env_mask is None:
                    self.rnn.zero_()
                else:
                    self.rnn.view(self.n_envs, self.n_agents, 128)[env_mask.bool()] = 0
            if self.opponents is not None:
                self.opponents.reset(env_mask)
            self._was_reset = True
            return self.batch.reset(env_mask)

    def low_level_actions(self, actions: torch.Tensor) -> torch.Tensor:
        """normalize_action's discrete half: the task's action rows -> low-level rows [n_envs, A, 4 + shoot_dim]."""
        B, A = self.n_envs, self.n_agents
        if not self.hier:
            return actions if actions.dtype == torch.int32 else actions.to(torch.int32)
        high = actions[..., :3].reshape(B * A, 3)
        obs = self.batch.obs.view(B * A, -1)
        x = hierarchical_input(high, obs, self._climb_below, self._luts)
        if self.opponents is not None:      # the red team's controller input comes from its scripted agent, not from the action
            x.view(B, A, 12)[:, self.spec.n_ego:] = self.opponents.inputs()
        low, h = self.controller(x, self.rnn)
        self.rnn.copy_(h)                                  # in place: the buffer is captured by the CUDA graph
        self._low.view(B * A, -1)[:, :4].copy_(low)        # one strided copy: `low` is a transposed view
        if self.spec.shoot_dim:
            self._low[..., 4:] = actions[..., 3:].to(torch.int32)
            if self.opponents is not None:  # scripted aircraft shoot everything when use_artillery, else nothing (scenario2_task.py:52-57)
                self._low[:, self.spec.n_ego:, 4:] = 1 if self.spec.use_artillery else 0
        return self._low

    def _step_body(self, actions: torch.Tensor):
        low = self.low_level_actions(actions)
        out = self.batch.step(low.contiguous(), auto_reset=self.auto_reset)
        if self.hier and self.auto_reset:
            # task.reset re-zeroes the controller's recurrent state of the envs that were just reset
            self.rnn.view(self.n_envs, self.n_agents, 128).mul_((1 - self.batch.env_done.view(-1, 1, 1)).to(torch.float32))
            if self.opponents is not None:
                self.opponents.reset(self.batch.env_done)
        return out

    def _warm_for_capture(self):
        if self.hier:      # warm the controller's cuBLAS / LayerNorm paths outside the capture, leaving no trace in the state
            keep = [self.rnn, self._low]
            if self.opponents is not None:
                keep += [self.opponents.step, self.opponents.init_heading, self.opponents.has_init]
            saved = [t.clone() for t in keep]
            for _ in range(3):
                self.low_level_actions(self._act_in)
            for t, s0 in zip(keep, saved):
                t.copy_(s0)
        torch.cuda.synchronize(self.device)

    def _capture(self):
        self._warm_for_capture()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._graph_out = self._step_body(self._act_in)
        self._graph = g

    def set_timing(self, on: bool):
        """Per-kernel CUDA-event timing (bench).  Events recorded inside a captured graph cannot be queried, so timed
        steps launch eagerly; the caller keeps the GPU busy ahead of them (bench.py) so the intervals hold no launch gaps."""
        self._timing = bool(on)
        self.batch.set_timing(on)

    def step(self, actions: torch.Tensor):
        """actions: integer tensor [n_envs, n_agents, act_dim] on this env's device."""
        if not self._was_reset:
            raise AcsError("step() called before reset()")
        assert actions.shape == (self.n_envs, self.n_agents, self.act_dim), (tuple(actions.shape), self.act_dim)
        with torch.cuda.device(self.device):
            if not self.use_cuda_graph or self._timing:
                return self._step_body(actions if actions.dtype == torch.int32 else actions.to(torch.int32))
            if actions.data_ptr() != self._act_in.data_ptr():
                self._act_in.copy_(actions)
            if self._graph is None:
                self._capture()
            self._graph.replay()
            return self._graph_out


def info_dict(info, heading: bool) -> dict:
    """One env's ``info`` dict from its [A, ACS_INFO_DIM] int rows.  Keys as the reference: ``current_step`` always
    (envs/JSBSim/envs/env_base.py:132), ``done_condition`` when an agent terminated this step, ``heading_turn_counts`` only
    when UnreachHeading is what terminated it (termination_conditions/unreach_heading.py:60-62) -- the runner appends that
    key whenever it is present (runner/jsbsim_runner.py:55-57), so it must not appear on ordinary steps."""
    d = {"current_step": int(info[0, 2])}
    causes = [int(c) for c in info[:, 0]]
    if any(c >= 0 for c in causes):
        d["done_condition"] = [DONE_CONDITIONS.get(c, "") for c in causes]
    if heading and causes[0] == ts.T_UNREACH_HEADING:
        d["heading_turn_counts"] = int(info[0, 3])
    return d


class LazyInfo(dict):
    """``infos[i]`` of the VecEnv contract: a dict VIEW over the step's info arrays, materialised on access, so that
    building ``n_envs`` Python dicts is not on the step path.  It subclasses ``dict`` (the reference's tests assert
    ``isinstance(infos[0], dict)``) but keeps no items of its own: every reader -- ``[]``, ``in``, ``get``, iteration,
    ``keys / items / values``, ``len``, ``==``, ``copy()``, pickling -- goes through ``info_dict``.  Because ``__iter__``
    is overridden, CPython's ``dict(info)``, ``{**info}`` and ``f(**info)`` take the generic mapping path (``keys()`` +
    ``[]``) instead of copying the empty storage, and ``json.dumps(info)`` calls ``items()``: all of them see the content."""
    __slots__ = ("_src", "_i")

    def __init__(self, src, i):
        # one placeholder item in the C-level storage: json's C encoder writes "{}" for a dict whose storage is empty
        # before it ever asks for items(); with a non-empty storage it calls the overridden items()
        super().__init__(current_step=None)
        self._src, self._i = src, i

    def _materialise(self):
        return info_dict(self._src["info"][self._i], self._src["heading"])

    def __getitem__(self, k):
        return self._materialise()[k]

    def __contains__(self, k):
        return k in self._materialise()

    def get(self, k, default=None):
        return self._materialise().get(k, default)

    def keys(self):
        return self._materialise().keys()

    def items(self):
        return self._materialise().items()

    def values(self):
        return self._materialise().values()

    def __iter__(self):
        return iter(self._materialise())

    def __len__(self):
        return len(self._materialise())

    def __bool__(self):
        return True

    def copy(self):
        return self._materialise()

    def __reduce__(self):
        return (dict, (self._materialise(),))

    def __repr__(self):
        return repr(self._materialise())

    def __eq__(self, other):
        return self._materialise() == other

    def __ne__(self, other):
        return self._materialise() != other

    __hash__ = None


class AircraftView:
    """Read/poke access to one aircraft of env 0 -- the slice of ``AircraftSimulator`` (reference
    envs/JSBSim/core/simulatior.py:89-325) the reference's tests and render scripts touch.  Off the hot path: every
    access copies an arena from the device."""

    def __init__(self, env: "_SingleEnvBase", uid: str, index: int):
        self._env, self.uid, self.index = env, uid, index
        self.partners: List["AircraftView"] = []
        self.enemies: List["AircraftView"] = []
        cfg = env.core.config["aircraft_configs"][uid]
        self.color = cfg.get("color", "Red")
        self.model = cfg.get("model", "f16")
        self.num_missiles = int(cfg.get("missile", 0))

    def _field(self, arena, name):
        names, t = self._env.core.batch.arena(arena)
        return t[names.index(name), self.index]

    def _set_field(self, arena, name, value):
        b = self._env.core.batch
        names, t = b.arena(arena)
        t[names.index(name), self.index] = value
        b.set_arena(arena, t)

    @property
    def status(self) -> int:
        return int(self._field("ac_i", "status"))

    is_alive = property(lambda s: s.status == 0)
    is_crash = property(lambda s: s.status == 1)
    is_shotdown = property(lambda s: s.status == 2)

    def crash(self):
        self._set_field("ac_i", "status", 1)

    def shotdown(self):
        self._set_field("ac_i", "status", 2)

    @property
    def bloods(self) -> float:
        return float(self._field("ac_d", "bloods"))

    def get_position(self):
        return np.array([float(self._field("ac_d", k)) for k in ("pos_n", "pos_e", "pos_u")])

    def get_velocity(self):
        return np.array([float(self._field("ac_d", k)) for k in ("vel_n", "vel_e", "vel_d")])

    def get_geodetic(self):
        return np.array([float(self._field("out", "lon_deg")), float(self._field("out", "lat_geod_deg")),
                         float(self._field("ac_d", "h_sl_m"))])

    def get_rpy(self):
        return np.array([float(self._field("out", k)) for k in ("roll_rad", "pitch_rad", "heading_rad")])

    def get_sim_time(self):
        return float(self._field("fdm", "sim_time"))

    def get_property_value(self, name: str):
        """FDM outputs / state by this package's field names (``acs_output_field_name`` / ``acs_state_field_name``)."""
        b = self._env.core.batch
        for arena in ("out", "fdm", "ac_d"):
            names, t = b.arena(arena)
            if name in names:
                return float(t[names.index(name), self.index])
        raise KeyError(name)


class _SingleEnvBase:
    """One environment (a device batch of 1) with the reference's numpy-facing gym-style API."""
    ENV_KIND = None

    def __init__(self, config_name: str, device: int = 0, config_dir: Optional[str] = None, substeps: Optional[int] = None,
                 controller_path: Optional[str] = None, allow_random_controller: bool = False):
        self.core = BatchedEnv(config_name, 1, device=device, config_dir=config_dir, substeps=substeps,
                               controller_path=controller_path, auto_reset=False, allow_random_controller=allow_random_controller)
        kind = self.core.task_desc["env"]
        from .task_api import Task
        self.task = Task(self.core, single=True)
        if self.ENV_KIND is not None and kind != self.ENV_KIND:
            raise NotImplementedError(f"Unknown taskname: {self.core.task_name}")   # load_task of the reference env classes
        self.config_name = config_name
        self.config = self.core.config
        self.current_step = 0
        self._jsbsims: Dict[str, AircraftView] = {}
        for k, uid in enumerate(self.core.ego_ids + self.core.enm_ids):
            self._jsbsims[uid] = AircraftView(self, uid, k)
        for u, a in self._jsbsims.items():
            for w, b in self._jsbsims.items():
                if u == w:
                    continue
                (a.partners if u[0] == w[0] else a.enemies).append(b)

    # reference-shaped attributes
    agents = property(lambda s: s._jsbsims)
    ego_ids = property(lambda s: s.core.ego_ids)
    enm_ids = property(lambda s: s.core.enm_ids)
    num_agents = property(lambda s: s.core.num_agents)
    observation_space = property(lambda s: s.core.observation_space)
    share_observation_space = property(lambda s: s.core.share_observation_space)
    action_space = property(lambda s: s.core.action_space)
    time_interval = property(lambda s: s.core.time_interval)
    max_steps = property(lambda s: s.core.max_steps)
    sim_freq = property(lambda s: s.core.spec.sim_freq)
    agent_interaction_steps = property(lambda s: s.core.spec.substeps)

    def seed(self, seed=None):
        return self.core.seed(seed)

    def close(self):
        self.core.close()

    def render(self, mode="txt", filepath="./JSBSimRecording.txt.acmi"):
        """TacView text log (reference envs/JSBSim/envs/env_base.py:207-250); other modes raise as in the reference."""
        if mode != "txt":
            raise NotImplementedError
        if getattr(self, "_acmi", None) is None:
            from .tacview import AcmiWriter
            self._acmi = AcmiWriter(self.core, 0)
        self._acmi.write(filepath, self.current_step * self.time_interval)

    def _actions(self, action):
        """Legal inputs as in the reference: list / tuple / ndarray of per-agent actions, Tuple-space samples included."""
        rows = []
        for a in action:
            if isinstance(a, (tuple, list)):
                rows.append(np.concatenate([np.atleast_1d(np.asarray(x)).ravel() for x in a]))
            else:
                rows.append(np.asarray(a).ravel())
        arr = np.stack(rows).astype(np.int32)
        assert arr.shape == (self.num_agents, self.core.act_dim), (arr.shape, self.core.act_dim)
        return torch.from_numpy(arr).to(self.core.device).unsqueeze(0)

    def _info(self, info_t):
        return info_dict(info_t[0].cpu().numpy(), self.core.spec.obs_kind == ts.OBS_HEADING)


class _PlainEnv(_SingleEnvBase):
    """BaseEnv.reset/step (reference envs/JSBSim/envs/env_base.py:98-173): obs [A, D]; rewards, dones [A, 1]; info."""

    def reset(self) -> np.ndarray:
        self.current_step = 0
        obs, _ = self.core.reset()
        return obs[0].cpu().numpy()

    def step(self, action):
        obs, _, rew, done, info = self.core.step(self._actions(action))
        self.current_step += 1
        return (obs[0].cpu().numpy(), rew[0].cpu().numpy().reshape(-1, 1), done[0].cpu().numpy().astype(bool).reshape(-1, 1),
                self._info(info))


class SingleControlEnv(_PlainEnv):
    """reference envs/JSBSim/envs/singlecontrol_env.py"""
    ENV_KIND = "control"


class SingleCombatEnv(_PlainEnv):
    """reference envs/JSBSim/envs/singlecombat_env.py"""
    ENV_KIND = "1v1"


class MultipleCombatEnv(_SingleEnvBase):
    """reference envs/JSBSim/envs/multiplecombat_env.py: reset -> (obs, share_obs); step -> (obs, share_obs, rewards,
    dones, info)."""
    ENV_KIND = "nvn"

    def reset(self):
        self.current_step = 0
        obs, share = self.core.reset()
        return obs[0].cpu().numpy(), share[0].cpu().numpy()

    def step(self, action):
        obs, share, rew, done, info = self.core.step(self._actions(action))
        self.current_step += 1
        return (obs[0].cpu().numpy(), share[0].cpu().numpy(), rew[0].cpu().numpy().reshape(-1, 1),
                done[0].cpu().numpy().astype(bool).reshape(-1, 1), self._info(info))
