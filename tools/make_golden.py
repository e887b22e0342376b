#!/usr/bin/env python
"""Generates tests/golden/env_*.npz: trajectories of the REFERENCE's own Python env layer.

Run in the build container only (it imports /root/reference, which does not travel to the GPU box):

    python tools/make_golden.py

What is real and what is shimmed.  The reference's env / task / reward / termination / missile / chaff code
(/root/reference/envs/JSBSim/{envs,tasks,reward_functions,termination_conditions,core,utils}) is imported and executed
UNMODIFIED.  Its third-party dependencies are absent from this image, so they are replaced by shims:
  * ``jsbsim``     -> ``FGFDMExec`` over the CPU oracle FDM (oracle/fdm.py); the FDM arithmetic is therefore NOT pinned by
                      these vectors (FDM parity stays "unpinned") -- everything above it is;
  * ``pymap3d``    -> geodetic2ned / ned2geodetic restated from the published algorithm (same formulas as the oracle);
  * ``gymnasium``  -> the package's own space classes + a seeding stub; ``colorama`` -> empty strings.
The reference's hierarchical tasks call a GRU controller inside normalize_action; the C-ABI boundary sits below that
controller, so the two ``Hierarchical*Task.normalize_action`` methods are patched to the direct stick/throttle mapping
(the non-hierarchical base-class code, reference tasks/singlecombat_task.py:141-153).  Random draws (np.random.rand in
the chaff test, env.np_random.uniform in the heading task) are replaced by the keyed generator of the oracle so that the
restatement can be compared draw for draw.
"""
from __future__ import annotations

import json
import math
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from oracle import env_oracle as eo            # noqa: E402
from oracle.fdm import OracleFdm               # noqa: E402
from aircombat_selfplay_b200 import spaces as my_spaces   # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

# ----------------------------------------------------------------------------- shims
_SNAP = {
    "position/long-gc-deg": "lon_deg", "position/lat-geod-deg": "lat_geod_deg", "position/h-sl-ft": "h_sl_ft",
    "attitude/roll-rad": "roll_rad", "attitude/pitch-rad": "pitch_rad", "attitude/heading-true-rad": "heading_rad",
    "attitude/psi-deg": "psi_deg", "velocities/v-north-fps": "v_north_fps", "velocities/v-east-fps": "v_east_fps",
    "velocities/v-down-fps": "v_down_fps", "velocities/u-fps": "u_fps", "velocities/v-fps": "v_fps", "velocities/w-fps": "w_fps",
    "velocities/vc-fps": "vc_fps", "velocities/p-rad_sec": "p_rad_sec", "velocities/q-rad_sec": "q_rad_sec",
    "velocities/r-rad_sec": "r_rad_sec", "velocities/eci-velocity-mag-fps": "eci_velocity_mag_fps",
    "accelerations/n-pilot-x-norm": "n_pilot_x", "accelerations/n-pilot-y-norm": "n_pilot_y",
    "accelerations/n-pilot-z-norm": "n_pilot_z", "simulation/sim-time-sec": "sim_time",
}
_CMD = ["fcs/aileron-cmd-norm", "fcs/elevator-cmd-norm", "fcs/rudder-cmd-norm", "fcs/throttle-cmd-norm"]
_IC = ["ic/long-gc-deg", "ic/lat-geod-deg", "ic/h-sl-ft", "ic/psi-true-deg", "ic/u-fps", "ic/v-fps", "ic/w-fps", "ic/p-rad_sec",
       "ic/q-rad_sec", "ic/r-rad_sec", "ic/phi-deg", "ic/theta-deg"]


class _Engine:
    def init_running(self):
        pass


class _Propulsion:
    def get_num_engines(self):
        return 1

    def get_engine(self, j):
        return _Engine()

    def get_steady_state(self):
        return True


class FGFDMExec:
    """The slice of jsbsim.FGFDMExec the reference touches (core/simulatior.py:165-188,223,261,295,313)."""

    def __init__(self, root_dir=None):
        self.dt = 1.0 / 120.0
        self.props = {k: 0.0 for k in _IC + _CMD}
        self.fdm = None
        self._cache = None

    def set_debug_level(self, n):
        pass

    def load_model(self, name):
        assert name == "f16"
        return True

    def query_property_catalog(self, check):
        return "\n".join(f"{n} (RW)" for n in list(_SNAP) + _CMD + _IC) + "\n"

    def set_dt(self, dt):
        self.dt = dt

    def run_ic(self):
        self.fdm = OracleFdm(self.dt, 1.0 / 120.0)   # FCS components latch the load-time dt (1/120), DESIGN.md F11
        self.fdm.reset(*[self.props[k] for k in _IC])
        self.fdm.set_controls(*[self.props[k] for k in _CMD])
        self._cache = None
        return True

    def get_propulsion(self):
        return _Propulsion()

    def run(self):
        self.fdm.set_controls(*[self.props[k] for k in _CMD])
        self.fdm.run(1)
        self._cache = None
        return True

    def get_sim_time(self):
        return self.get_property_value("simulation/sim-time-sec")

    def get_property_value(self, name):
        if name in _SNAP:
            if self._cache is None:
                self._cache = self.fdm.snapshot_dict()
            return float(self._cache[_SNAP[name]])
        return float(self.props.get(name, 0.0))

    def set_property_value(self, name, value):
        self.props[name] = float(value)


def _geodetic2ned(lat, lon, h, lat0, lon0, h0, ell=None, deg=True):
    n, e, u = eo.LLA2NEU(lon, lat, h, lon0, lat0, h0)
    return n, e, -u


def _ned2geodetic(n, e, d, lat0, lon0, h0, ell=None, deg=True):
    u = -d
    la, lo = math.radians(lat0), math.radians(lon0)
    x0, y0, z0 = eo.geodetic2ecef(lat0, lon0, h0)
    t = math.cos(la) * u - math.sin(la) * n
    w = math.sin(la) * u + math.cos(la) * n
    x = x0 + math.cos(lo) * t - math.sin(lo) * e
    y = y0 + math.sin(lo) * t + math.cos(lo) * e
    z = z0 + w
    a, b = eo.WGS84_A, eo.WGS84_B
    r = math.sqrt(x * x + y * y + z * z)
    E = math.sqrt(a * a - b * b)
    uq = math.sqrt(0.5 * (r * r - E * E) + 0.5 * math.hypot(r * r - E * E, 2 * E * z))
    Q = math.hypot(x, y)
    huE = math.hypot(uq, E)
    beta = math.atan(huE / uq * z / Q)
    beta += ((b * uq - a * huE + E * E) * math.sin(beta)) / (a * huE / math.cos(beta) - E * E * math.cos(beta))
    lat = math.atan(a / b * math.tan(beta))
    lon = math.atan2(y, x)
    return math.degrees(lat), math.degrees(lon), eo.NEU2alt(n, e, u, lon0, lat0, h0)


class _KeyedRandom:
    """env.np_random replacement: uniform(lo, hi) consumes the oracle's keyed draws in call order (reset: heading,
    altitude, velocity; every UnreachHeading re-target: heading, altitude, velocity)."""

    def __init__(self, seed, env_index):
        self.seed, self.env_index = seed, env_index
        self.episode, self.turn, self.k = -1, 0, 0
        self.in_reset = False

    def begin_reset(self):
        self.episode += 1
        self.turn, self.k, self.in_reset = 0, 0, True

    def end_reset(self):
        self.in_reset, self.k = False, 0

    def uniform(self, lo, hi):
        if self.in_reset:
            d = eo.u01(self.seed, self.env_index, eo.RNG_RESET, self.episode, self.k)
            self.k += 1
        else:
            d = eo.u01(self.seed, self.env_index, eo.RNG_HEADING, self.episode, self.turn, self.k)
            self.k += 1
            if self.k == 3:
                self.k, self.turn = 0, self.turn + 1
        return lo + (hi - lo) * d


def install_shims():
    js = types.ModuleType("jsbsim")
    js.FGFDMExec = FGFDMExec
    sys.modules["jsbsim"] = js
    pm = types.ModuleType("pymap3d")
    pm.geodetic2ned, pm.ned2geodetic = _geodetic2ned, _ned2geodetic
    sys.modules["pymap3d"] = pm
    col = types.ModuleType("colorama")
    class _Empty(type):
        def __getattr__(cls, name):
            return ""
    col.Fore = _Empty("Fore", (), {})
    col.Back = col.Style = col.Fore
    sys.modules["colorama"] = col
    gym = types.ModuleType("gymnasium")
    gym.Env = type("Env", (), {})
    gym.Space = object
    sp = types.ModuleType("gymnasium.spaces")
    for n in ("Box", "Discrete", "MultiDiscrete", "Tuple"):
        setattr(sp, n, getattr(my_spaces, n))
    gym.spaces = sp
    ut = types.ModuleType("gymnasium.utils")
    seeding = types.ModuleType("gymnasium.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.default_rng(seed), seed)
    ut.seeding = seeding
    gym.utils = ut
    sys.modules.update({"gymnasium": gym, "gymnasium.spaces": sp, "gymnasium.utils": ut, "gymnasium.utils.seeding": seeding})
    for name in ("matplotlib", "matplotlib.pyplot", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.agent = None
                sys.modules[name] = m
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def import_reference():
    install_shims()
    sys.path.insert(0, str(REF))
    import torch
    real_load = torch.load
    torch.load = lambda *a, **k: {}
    from envs.JSBSim.model import baseline_actor
    # every BaselineActor the reference builds (the tasks' low-level policy, the scripted agents' actor) gets the seeded
    # weights of tests/golden/controller.npz instead of baseline_model.pt (which the reference loads onto 'cuda')
    _g = np.load(GOLDEN / "controller.npz")
    _sd = {k[3:]: torch.from_numpy(_g[k]) for k in _g.files if k.startswith("sd:")}
    _orig_lsd = torch.nn.Module.load_state_dict
    baseline_actor.BaselineActor.load_state_dict = lambda self, sd, *a, **k: _orig_lsd(self, _sd)
    from envs.JSBSim.envs import env_base, multiplecombat_env, singlecombat_env, singlecontrol_env
    from envs.JSBSim.tasks import multiplecombat_task, singlecombat_task
    del real_load      # torch.load stays stubbed: the hierarchical Task constructors load baseline_model.pt onto 'cuda'

    def direct(self, env, agent_id, action):   # the non-hierarchical mapping (reference tasks/singlecombat_task.py:141-153)
        norm_act = np.zeros(4)
        norm_act[0] = action[0] / 20 - 1.
        norm_act[1] = action[1] / 20 - 1.
        norm_act[2] = action[2] / 20 - 1.
        norm_act[3] = action[3] / 58 + 0.4
        return norm_act
    singlecombat_task.HierarchicalSingleCombatTask.normalize_action = direct
    multiplecombat_task.HierarchicalMultipleCombatTask.normalize_action = direct

    def shoot_nearest(self, env, agent_id, action):   # reference tasks/multiplecombat_task.py:274-276 with 4 low-level ints
        self._shoot_action[agent_id] = action[4] > 0
        return direct(self, env, agent_id, action[:4])
    multiplecombat_task.HierarchicalMultipleCombatShootTask.normalize_action = shoot_nearest
    return env_base, singlecontrol_env, singlecombat_env, multiplecombat_env


# ----------------------------------------------------------------------------- cases
def base_config(task, n_per_team, K, missile=None, extra=None, lat_gap=0.1, h_ft=20000.0, max_steps=1000, dh_enemy=0.0):
    acs = {}
    for team, psi, color in (("A", 0.0, "Blue"), ("B", 180.0, "Red")):
        if n_per_team == 0 and team == "B":
            continue
        for k in range(max(n_per_team, 1)):
            d = {"color": color, "model": "f16",
                 "init_state": {"ic_long_gc_deg": 120.0 + 0.01 * k + (0.003 if team == "B" else 0.0),
                                "ic_lat_geod_deg": 60.0 if team == "A" else 60.0 + lat_gap,
                                "ic_h_sl_ft": h_ft + (dh_enemy if team == "B" else 0.0), "ic_psi_true_deg": psi, "ic_u_fps": 800.0}}
            if missile is not None:
                d["missile"] = missile
            acs[f"{team}0{k + 1}00"] = d
    cfg = {"task": task, "sim_freq": 60, "agent_interaction_steps": K, "max_steps": max_steps, "altitude_limit": 2500,
           "acceleration_limit_x": 10.0, "acceleration_limit_y": 10.0, "acceleration_limit_z": 10.0,
           "battle_field_center": [120.0, 60.0, 0.0], "aircraft_configs": acs,
           "PostureReward_scale": 15.0, "PostureReward_potential": True, "PostureReward_orientation_version": "v2",
           "PostureReward_range_version": "v3", "AltitudeReward_safe_altitude": 4.0, "AltitudeReward_danger_altitude": 3.5,
           "AltitudeReward_Kv": 0.2, "max_attack_angle": 45, "max_attack_distance": 14000, "min_attack_interval": 25}
    cfg.update(extra or {})
    return cfg


def heading_config():
    cfg = base_config("heading", 0, 6, max_steps=10000)
    cfg["aircraft_configs"]["A0100"].update({"max_heading_increment": 180, "max_altitude_increment": 7000,
                                             "max_velocities_u_increment": 100, "check_interval": 30})
    return cfg


def actions_for(rng, A, shoot_dim, T, mode, shoot_p=0.3):
    act = np.zeros((T, A, 4 + shoot_dim), dtype=np.int64)
    if mode == "random":
        act[..., 0:3] = rng.integers(0, 41, (T, A, 3))
        act[..., 3] = rng.integers(0, 30, (T, A))
    elif mode == "smooth":
        act[..., 0:3] = 20 + rng.integers(-3, 4, (T, A, 3))
        act[..., 3] = rng.integers(10, 30, (T, A))
    elif mode == "dive":
        act[..., 0:4] = np.array([20, 40, 20, 29])
    elif mode == "straight":
        act[..., 0:4] = np.array([20, 19, 20, 0])
    elif mode == "chase":        # ego at full throttle behind an idling enemy on the same track: closes into gun range
        act[..., 0:4] = np.array([20, 19, 20, 0])
        act[:, : A // 2, 3] = 29
    if shoot_dim:
        act[..., 4:] = rng.random((T, A, shoot_dim)) < shoot_p
    return act


CASES = [
    # name, env kind, reference Task class (None = the env's own load_task), config, T, action mode, shoot_dim
    ("heading_random", "control", None, heading_config(), 60, "random", 0),
    ("heading_straight", "control", None, heading_config(), 40, "straight", 0),
    ("1v1_noweapon_random", "1v1", "SingleCombatTask", base_config("singlecombat", 1, 12), 40, "random", 0),
    ("1v1_noweapon_close", "1v1", "SingleCombatTask", base_config("singlecombat", 1, 12, lat_gap=0.05, dh_enemy=400.0), 60, "smooth", 0),
    ("1v1_noweapon_crash", "1v1", "SingleCombatTask", base_config("singlecombat", 1, 12, h_ft=9200.0), 40, "dive", 0),
    ("1v1_artillery_close", "1v1", "SingleCombatTask", base_config("singlecombat", 1, 12, lat_gap=0.03, extra={"use_artillery": True}), 60, "straight", 0),
    ("1v1_dodge_close", "1v1", "SingleCombatDodgeMissileTask",
     base_config("singlecombat_dodge_missile", 1, 12, missile=4, lat_gap=0.08, extra={"MissilePostureReward_scale": 30}), 80, "smooth", 0),
    ("1v1_shoot_close", "1v1", "SingleCombatShootMissileTask", base_config("singlecombat_shoot", 1, 12, missile=4, lat_gap=0.07, dh_enemy=300.0), 80, "smooth", 1),
    ("2v2_noweapon_random", "nvn", None, base_config("multiplecombat", 2, 12), 30, "random", 0),
    ("2v2_noweapon_crash", "nvn", None, base_config("multiplecombat", 2, 12, h_ft=9200.0), 40, "dive", 0),
    ("2v2_shoot_nearest_close", "nvn", "HierarchicalMultipleCombatShootTask:multiplecombat_task",
     base_config("hierarchical_multiplecombat_shoot_nearest", 2, 12, missile=4, lat_gap=0.07, extra={"MissilePostureReward_scale": 30}), 80, "smooth", 1),
    ("scenario2_close", "nvn", None, base_config("scenario2", 2, 6, missile=2, lat_gap=0.07, dh_enemy=300.0, max_steps=9000), 150, "smooth", 4),
    ("scenario2_random", "nvn", None, base_config("scenario2", 2, 6, missile=2, max_steps=9000), 40, "random", 4),
    ("scenario2_nvn_close", "nvn", None, base_config("scenario2_nvn", 2, 6, missile=2, lat_gap=0.06, max_steps=9000), 150, "smooth", 4),
    ("scenario3_close", "nvn", None, base_config("scenario3", 4, 6, missile=2, lat_gap=0.06, dh_enemy=-300.0, max_steps=9000), 120, "smooth", 4),
    ("scenario3_nvn_random", "nvn", None, base_config("scenario3_nvn", 4, 6, missile=2, max_steps=9000), 30, "random", 4),
    ("approach_random", "control", None, base_config("approach", 0, 6, max_steps=10000), 40, "random", 0),
    ("scenario2_rwr_close", "nvn", None, base_config("scenario2_rwr", 2, 6, missile=2, lat_gap=0.07, dh_enemy=200.0, max_steps=9000), 100, "smooth", 4),
    ("scenario3_rwr_random", "nvn", None, base_config("scenario3_rwr", 4, 6, missile=2, max_steps=9000), 30, "random", 4),
    # curriculum resets (stage 0: tail chase 11 km behind a north-bound enemy) + the automatic gun of WVRTask /
    # Maneuver_curriculum: -5 blood per step inside 3 km and 5 degrees until the enemy is SHOTDOWN by blood
    ("wvr_chase", "1v1", None, base_config("wvr", 1, 6, extra={"use_artillery": True}, max_steps=9000), 1500, "chase_cl", 0),
    ("maneuver_curriculum_chase", "1v1", None, base_config("maneuver_curriculum", 1, 6, max_steps=9000), 1500, "chase_cl", 0),
    ("scenario2_curriculum_random", "nvn", None, base_config("scenario2_curriculum", 2, 6, missile=2, max_steps=9000), 40, "random", 4),
    # scripted red team (use_baseline): PursueAgent / ManeuverAgent drive the enemy through the BaselineActor GRU
    ("scenario2_vs_pursue", "nvn", None, base_config("scenario2", 2, 6, missile=2, lat_gap=0.08, max_steps=9000,
                                                     extra={"use_baseline": True, "baseline_type": "pursue", "use_artillery": True}), 120, "smooth", 4),
    ("scenario1_vs_pursue", "1v1", None, base_config("scenario1", 1, 6, missile=2, lat_gap=0.08, max_steps=9000,
                                                     extra={"use_baseline": True, "baseline_type": "pursue", "use_artillery": False}), 120, "smooth", 4),
    ("scenario1_vs_maneuver", "1v1", None, base_config("scenario1", 1, 6, missile=2, max_steps=9000,
                                                       extra={"use_baseline": True, "baseline_type": "maneuver", "use_artillery": False}), 400, "smooth", 4),
    ("wvr_vs_pursue", "1v1", None, base_config("wvr", 1, 6, max_steps=9000,
                                               extra={"use_baseline": True, "baseline_type": "pursue", "use_artillery": True}), 150, "smooth", 0),
    # guns (R < 3 km, AO < 5 deg: -5 blood per burst until SHOTDOWN by blood) and many chaff bursts
    ("scenario2_gun_chaff", "nvn", None, base_config("scenario2", 2, 6, missile=30, lat_gap=0.04, max_steps=9000), 120, "straight", 4, 0.8),
    ("scenario3_gun_chaff", "nvn", None, base_config("scenario3", 4, 6, missile=6, lat_gap=0.05, max_steps=9000), 100, "smooth", 4, 0.6),
    # the chaff decoy draw on BOTH sides of 0.85 (CHAFF_DRAWS below): missiles that fly through a cloud unharmed for a
    # while, some of them on to their target, and missiles that are decoyed
    ("scenario2_chaff_mixed", "nvn", None, base_config("scenario2", 2, 6, missile=30, lat_gap=0.04, max_steps=9000), 120, "straight", 4, 0.8),
    ("scenario3_chaff_mixed", "nvn", None, base_config("scenario3", 4, 6, missile=6, lat_gap=0.05, max_steps=9000), 100, "smooth", 4, 0.6),
]

# np.random.rand() of the chaff test (reference envs/JSBSim/envs/env_base.py:153), by case and call index.  Every other
# case keeps the constant 0.5 (always below 0.85: the first chaff in range decoys the missile).
CHAFF_DRAWS = {
    "scenario2_chaff_mixed": lambda k: 0.5 if (k % 41) == 40 else 0.85 + 0.1 * ((k * 7) % 10) / 10.0,    # mostly >= 0.85, 0.85 itself included
    "scenario3_chaff_mixed": lambda k: 0.2 if (k % 13) == 12 else 0.95,
}


def run_case(mods, name, kind, task_cls, cfg, T, mode, shoot_dim, shoot_p=0.3, seed=7):
    env_base, singlecontrol_env, singlecombat_env, multiplecombat_env = mods
    import envs.JSBSim.tasks as ref_tasks
    from envs.JSBSim.tasks import multiplecombat_task as mct
    EnvConfig = type("EnvConfig", (object,), dict(cfg))
    for m in (env_base,):
        m.parse_config = lambda filename, _c=EnvConfig: _c
    cls = {"control": singlecontrol_env.SingleControlEnv, "1v1": singlecombat_env.SingleCombatEnv,
           "nvn": multiplecombat_env.MultipleCombatEnv}[kind]
    if task_cls is not None:
        if ":" in task_cls:
            cname, _ = task_cls.split(":")
            tcls = getattr(mct, cname)
        else:
            tcls = getattr(ref_tasks, task_cls)

        class Patched(cls):
            def load_task(self):
                self.task = tcls(self.config)
        cls = Patched
    env = cls("golden")
    keyed = _KeyedRandom(seed, 0)
    env.np_random = keyed
    draws = []

    def chaff_draw(*a):        # chaff draw (env_base.py:153); the consumed sequence is stored, the oracle test replays it
        v = CHAFF_DRAWS[name](len(draws)) if name in CHAFF_DRAWS else 0.5
        draws.append(v)
        return v
    np.random.rand = chaff_draw
    A = len(cfg["aircraft_configs"])
    rng = np.random.default_rng(seed)
    pilot = None
    if mode == "chase_cl":
        # closed-loop action generator: an altitude / wings-level hold flown on the ORACLE env stepped alongside, ego on
        # high throttle behind an idling enemy -- only a way to produce an action sequence that reaches gun range; the
        # recorded actions are then the inputs of both sides of the comparison
        from aircombat_selfplay_b200.tasks import build_spec
        pilot = eo.OracleEnv(build_spec(cfg), seed=seed, env_index=0)
        pilot.reset()
        acts = np.zeros((T, A, 4), dtype=np.int64)
    else:
        acts = actions_for(rng, A, shoot_dim, T, mode, shoot_p)
    keyed.begin_reset()
    r = env.reset()
    keyed.end_reset()
    obs0 = r[0] if kind == "nvn" else r
    obs, rews, dones, status, share = [np.asarray(obs0, dtype=np.float64)], [], [], [], []
    cond, missiles, bloods, chaffs = [], [], [], []
    steps_done = 0
    for t in range(T):
        if pilot is not None:
            for k, sim in enumerate(pilot.sims):
                e = 19 + int(round(np.clip(0.01 * (sim.h_sl_m - 6096) - 0.1 * sim.velocity[2], -8, 8)))
                al = 20 - int(round(np.clip(sim.posture[0] * 15, -8, 8)))
                acts[t, k] = [al, e, 20, 29 if k < A // 2 else 0]
            pilot.step(acts[t])
        a = acts[t].astype(np.float64) if shoot_dim else acts[t]
        out = env.step(a)
        if kind == "nvn":
            o, s, rw, d, info = out
            share.append(np.asarray(s, dtype=np.float64))
        else:
            o, rw, d, info = out
        obs.append(np.asarray(o, dtype=np.float64)); rews.append(np.asarray(rw, dtype=np.float64).reshape(-1))
        dones.append(np.asarray(d).reshape(-1).astype(bool))
        status.append([0 if s.is_alive else (1 if s.is_crash else 2) for s in env.agents.values()])
        bloods.append([float(s.bloods) for s in env.agents.values()])
        missiles.append(len(env._tempsims))
        chaffs.append(len(env._chaffsims))
        steps_done += 1
        if np.all(dones[-1]):
            break
    out = {"config": json.dumps(cfg), "kind": kind, "seed": seed, "actions": acts[:steps_done], "obs": np.stack(obs),
           "rewards": np.stack(rews), "dones": np.stack(dones), "status": np.array(status), "bloods": np.array(bloods),
           "n_missiles": np.array(missiles), "n_chaffs": np.array(chaffs), "chaff_draws": np.array(draws, dtype=np.float64)}
    if share:
        out["share_obs"] = np.stack(share)
    if kind == "control":
        out["turn_counts"] = np.array(env.heading_turn_counts)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"env_{name}.npz", **out)
    ev = f"steps={steps_done} missiles={max(missiles) if missiles else 0} chaffs={max(chaffs) if chaffs else 0} draws={len(draws)} (>=0.85: {sum(d >= 0.85 for d in draws)}) status={sorted(set(np.array(status).ravel().tolist()))} " \
         f"min_blood={np.min(bloods):.1f} done={bool(np.all(dones[-1]))} sum_rew={np.sum(rews):.3f}"
    print(f"[golden] {name}: {ev}")


def main():
    import contextlib
    import io
    only = sys.argv[1:]
    with contextlib.redirect_stdout(io.StringIO()):
        mods = import_reference()
    for case in CASES:
        if only and case[0] not in only:
            continue
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                run_case(mods, *case)
        except Exception:
            sys.stdout.write(buf.getvalue()[-2000:])
            raise
        lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("[golden]")]
        print("\n".join(lines))


if __name__ == "__main__":
    main()
