"""Task registry: yaml scenario config -> ``TaskSpec``.

Mirrors what the reference's Task classes resolve to (reference envs/JSBSim/tasks/*.py, "E/tasks" below) once their
Python class composition (MRO) is followed: observation packer, action layout, ordered reward functions with their
``{ClassName}_{key}`` hyper-parameters (E/reward_functions/reward_function_base.py:14-15), ordered termination
conditions, weapon-launch rule, reward gating and step ordering.  The yaml schema is the reference's own
(E/configs/**.yaml), so its config files can be used unchanged (``config_dir=`` of the env constructors).
"""
from __future__ import annotations

import math
import os
from pathlib import Path
from typing import Dict

import yaml

from . import taskspec as ts
from .taskspec import RewardSpec, TaskSpec

CONFIG_DIR = Path(__file__).resolve().parent / "configs"

_IC_KEYS = ["ic_long_gc_deg", "ic_lat_geod_deg", "ic_h_sl_ft", "ic_psi_true_deg", "ic_u_fps", "ic_v_fps", "ic_w_fps",
            "ic_p_rad_sec", "ic_q_rad_sec", "ic_r_rad_sec", "ic_phi_deg", "ic_theta_deg"]
# AircraftSimulator.clear_defalut_condition (E/core/simulatior.py:192-208)
_IC_DEFAULT = [120.0, 60.0, 20000.0, 0.0, 800.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


def parse_config(name_or_path: str, config_dir=None) -> dict:
    """Same lookup rule as E/utils/utils.py:7-23 (``<configs>/<name>.yaml``), returning the raw dict."""
    cands = []
    if os.path.isfile(name_or_path):
        cands.append(Path(name_or_path))
    for d in ([Path(config_dir)] if config_dir else []) + [CONFIG_DIR]:
        cands.append(d / f"{name_or_path}.yaml")
    for c in cands:
        if c.exists():
            with open(c, "r", encoding="utf-8") as f:
                return yaml.load(f, Loader=yaml.FullLoader)
    raise FileNotFoundError(f"config {name_or_path!r} not found (looked in {[str(c) for c in cands]})")


def init_state_row(init_state: dict):
    row = list(_IC_DEFAULT)
    for k, v in (init_state or {}).items():
        if k in _IC_KEYS:
            row[_IC_KEYS.index(k)] = float(v)
    return row


def _reward(cfg: dict, cls: str) -> RewardSpec:
    kind = ts.REWARD_CLASS[cls]
    r = RewardSpec(kind=kind, scale=float(cfg.get(f"{cls}_scale", 1.0)), potential=bool(cfg.get(f"{cls}_potential", False)))
    if kind == ts.R_ALTITUDE:       # E/reward_functions/altitude_reward.py:14-16
        r.p0, r.p1, r.p2 = float(cfg.get(f"{cls}_safe_altitude", 4.0)), float(cfg.get(f"{cls}_danger_altitude", 3.5)), float(cfg.get(f"{cls}_Kv", 0.2))
    elif kind == ts.R_POSTURE:      # E/reward_functions/posture_reward.py:18-20
        r.p0 = float(str(cfg.get(f"{cls}_orientation_version", "v2")).lstrip("v"))
        r.p1 = float(str(cfg.get(f"{cls}_range_version", "v3")).lstrip("v"))
        r.p2 = float(cfg.get(f"{cls}_target_dist", 3.0))
    elif kind == ts.R_RELATIVE_ALTITUDE:   # E/reward_functions/relative_altitude_reward.py:16
        r.p0 = float(cfg.get(f"{cls}_KH", 1.0))
    return r


_SCENARIO_REWARDS = ["AltitudeReward", "CombatGeometryReward", "EventDrivenReward", "GunBEHITReward", "GunTargetTailReward",
                     "GunWEZDOTReward", "GunWEZReward", "PostureReward", "RelativeAltitudeReward", "MissilePostureReward",
                     "ShootPenaltyReward"]
_TERMS_1V1 = [ts.T_LOW_ALTITUDE, ts.T_EXTREME_STATE, ts.T_OVERLOAD, ts.T_SAFE_RETURN, ts.T_TIMEOUT]      # E/tasks/singlecombat_task.py:34-40
_TERMS_NVN = [ts.T_SAFE_RETURN, ts.T_EXTREME_STATE, ts.T_OVERLOAD, ts.T_LOW_ALTITUDE, ts.T_TIMEOUT]      # E/tasks/multiplecombat_task.py:33-39
_TERMS_HEADING = [ts.T_UNREACH_HEADING, ts.T_EXTREME_STATE, ts.T_OVERLOAD, ts.T_LOW_ALTITUDE, ts.T_TIMEOUT]  # E/tasks/heading_task.py:20-26
_TERMS_NO_SAFE_RETURN = [ts.T_LOW_ALTITUDE, ts.T_EXTREME_STATE, ts.T_OVERLOAD, ts.T_TIMEOUT]              # E/tasks/approach_task.py:24-29, WVR_task.py:32-37

# task name -> (env kind, obs packer, reward classes, launch rule, shoot_dim, hierarchical, high-level action dims)
# env kind: "control" (SingleControlEnv), "1v1" (SingleCombatEnv), "nvn" (MultipleCombatEnv)
_R_BASE = ["AltitudeReward", "PostureReward", "EventDrivenReward"]
_R_DODGE = ["PostureReward", "MissilePostureReward", "AltitudeReward", "EventDrivenReward"]
_R_SHOOT = ["PostureReward", "AltitudeReward", "EventDrivenReward", "ShootPenaltyReward"]
_R_SHOOT_NVN = ["PostureReward", "AltitudeReward", "EventDrivenReward", "ShootPenaltyReward", "MissilePostureReward"]
_R_MANEUVER = ["AltitudeReward", "CombatGeometryReward", "EventDrivenReward", "GunBEHITReward", "GunTargetTailReward",
               "GunWEZDOTReward", "GunWEZReward", "PostureReward", "RelativeAltitudeReward"]       # singlecombat_task.py:267-277
_R_WVR = ["PostureReward", "AltitudeReward", "EventDrivenReward", "CombatGeometryReward", "GunBEHITReward", "GunTargetTailReward",
          "GunWEZReward", "GunWEZDOTReward"]                                                        # WVR_task.py:21-30
TASKS: Dict[str, dict] = {
    # SingleControlEnv (E/envs/singlecontrol_env.py:16-23)
    "heading": dict(env="control", obs=ts.OBS_HEADING, rewards=["HeadingReward", "AltitudeReward"], launch=ts.L_NONE, shoot=0, hier=False),
    "approach": dict(env="control", obs=ts.OBS_HEADING, rewards=["AltitudeReward"], launch=ts.L_NONE, shoot=0, hier=False,
                     terms=_TERMS_NO_SAFE_RETURN),                                                                                      # ApproachTask
    # SingleCombatEnv: LAG task names (the fork's load_task dropped these branches; the classes exist)
    "singlecombat": dict(env="1v1", obs=ts.OBS_1V1, rewards=_R_BASE, launch=ts.L_NONE, shoot=0, hier=False),                                  # SingleCombatTask
    "hierarchical_singlecombat": dict(env="1v1", obs=ts.OBS_1V1, rewards=_R_BASE, launch=ts.L_NONE, shoot=0, hier=True),                      # HierarchicalSingleCombatTask
    "singlecombat_dodge_missile": dict(env="1v1", obs=ts.OBS_1V1_MISSILE, rewards=_R_DODGE, launch=ts.L_RULE_LOCK, shoot=0, hier=False),      # SingleCombatDodgeMissileTask
    "hierarchical_singlecombat_dodge_missile": dict(env="1v1", obs=ts.OBS_1V1_MISSILE, rewards=_R_DODGE, launch=ts.L_RULE_LOCK, shoot=0, hier=True),
    "singlecombat_shoot": dict(env="1v1", obs=ts.OBS_1V1_MISSILE, rewards=_R_SHOOT, launch=ts.L_RL_SINGLE, shoot=1, hier=False),              # SingleCombatShootMissileTask
    "hierarchical_singlecombat_shoot": dict(env="1v1", obs=ts.OBS_1V1_MISSILE, rewards=_R_SHOOT, launch=ts.L_RL_SINGLE, shoot=1, hier=True),
    # SingleCombatEnv: this fork's names (E/envs/singlecombat_env.py:19-36)
    "scenario1": dict(env="1v1", obs=ts.OBS_1V1_MISSILE, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True),
    "scenario1_curriculum": dict(env="1v1", obs=ts.OBS_1V1_MISSILE, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, curriculum=True),
    "scenario1_rwr": dict(env="1v1", obs=ts.OBS_1V1_RWR, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True),
    "scenario1_rwr_curriculum": dict(env="1v1", obs=ts.OBS_1V1_RWR, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, curriculum=True),
    "maneuver_curriculum": dict(env="1v1", obs=ts.OBS_1V1, rewards=_R_MANEUVER, launch=ts.L_AUTO_GUN, shoot=0, hier=True, curriculum=True),
    "wvr": dict(env="1v1", obs=ts.OBS_1V1, rewards=_R_WVR, launch=ts.L_AUTO_GUN, shoot=0, hier=True, curriculum=True,
                terms=_TERMS_NO_SAFE_RETURN),
    # MultipleCombatEnv (E/envs/multiplecombat_env.py:25-66)
    "multiplecombat": dict(env="nvn", obs=ts.OBS_MULTI, rewards=_R_BASE, launch=ts.L_NONE, shoot=0, hier=False),
    "hierarchical_multiplecombat": dict(env="nvn", obs=ts.OBS_MULTI, rewards=_R_BASE, launch=ts.L_NONE, shoot=0, hier=True),
    "hierarchical_multiplecombat_shoot": dict(env="nvn", obs=ts.OBS_NV_MISSILE, rewards=_R_SHOOT, launch=ts.L_NONE, shoot=1, hier=True),       # multiplecombat_with_missile_task.py:219
    "multiplecombat_shoot": dict(env="nvn", obs=ts.OBS_NV_MISSILE, rewards=_R_SHOOT_NVN, launch=ts.L_NONE, shoot=1, hier=False),               # MultipleCombatShootMissileTask (:180)
    "hierarchical_multiplecombat_shoot_nearest": dict(env="nvn", obs=ts.OBS_MULTI_MISSILE, rewards=_R_DODGE, launch=ts.L_RL_NEAREST, shoot=1, hier=True),  # multiplecombat_task.py:201
    "scenario2": dict(env="nvn", obs=ts.OBS_NV_MISSILE, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True),
    "scenario2_curriculum": dict(env="nvn", obs=ts.OBS_NV_MISSILE, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, curriculum=True),
    "scenario2_nvn": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True),
    "scenario2_nvn_curriculum": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, curriculum=True),
    "scenario2_rwr": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, rwr=True),
    "scenario2_rwr_curriculum": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, rwr=True, curriculum=True),
    "scenario3_rwr": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, rwr=True),
    "scenario3_rwr_curriculum": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, rwr=True, curriculum=True),
    "scenario3": dict(env="nvn", obs=ts.OBS_NV_MISSILE, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True),
    "scenario3_curriculum": dict(env="nvn", obs=ts.OBS_NV_MISSILE, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, curriculum=True),
    "scenario3_nvn": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True),
    "scenario3_nvn_curriculum": dict(env="nvn", obs=ts.OBS_NVN, rewards=_SCENARIO_REWARDS, launch=ts.L_SCENARIO, shoot=4, hier=True, curriculum=True),
}


def curriculum_circle(center_lat, center_lon, radius_km, angle_deg):
    """(lat, lon, heading) on the curriculum circle: E/utils/utils.py:126-155 (calculate_coordinates_heading_by_curriculum)."""
    R = 6371
    d = radius_km / R
    clat, clon = math.radians(center_lat), math.radians(center_lon)
    theta = math.radians(180 - angle_deg)
    nlat = math.asin(math.sin(clat) * math.cos(d) + math.cos(clat) * math.sin(d) * math.cos(theta))
    nlon = clon + math.atan2(math.sin(theta) * math.sin(d) * math.cos(clat), math.cos(d) - math.sin(clat) * math.sin(nlat))
    heading = 2 * angle_deg if 0 <= angle_deg < 90 else 360 - 2 * angle_deg
    return math.degrees(nlat), math.degrees(nlon), heading


def curriculum_init_states(env_kind: str, rows, angle: int):
    """Initial conditions of the curriculum tasks' reset: ``reset_simulators_curriculum(angle)`` of SingleCombatEnv
    (E/envs/singlecombat_env.py:87-122) / MultipleCombatEnv (E/envs/multiplecombat_env.py:185-248).  ``rows`` are the yaml
    init-state rows in aircraft order; only the entries the reference ``update``s are replaced -- for the N-v-N env that
    is aircraft 0..3 whatever the team size (a 4v4 reset moves two ego aircraft onto the enemy start line; kept)."""
    rows = [list(r) for r in rows]

    def put(i, lat, lon, psi):
        rows[i][0], rows[i][1], rows[i][2], rows[i][3], rows[i][4] = lon, lat, 20000.0, float(psi), 800.0
    if env_kind == "1v1":
        lat, lon, psi = curriculum_circle(60.1, 120.0, 11.119, angle)
        put(0, lat, lon, psi)
        put(1, 60.1, 120.0, 0.0)
    elif env_kind == "nvn":
        lat, lon, psi = curriculum_circle(60.1, 120.0, 11.119, angle)
        put(0, lat, lon, psi)
        lat, lon, psi = curriculum_circle(60.1, 120.01, 11.119, angle)
        put(1, lat, lon, psi)
        put(2, 60.1, 120.0, 0.0)
        put(3, 60.1, 120.01, 0.0)
    return rows


def obs_dim_for(obs_kind: int, n_agents: int, n_aircraft_cfg: int, rwr: bool = False) -> int:
    if obs_kind == ts.OBS_HEADING:
        return 12
    if obs_kind == ts.OBS_1V1_RWR:
        return 23
    if obs_kind == ts.OBS_1V1:
        return 15
    if obs_kind in (ts.OBS_1V1_MISSILE, ts.OBS_NV_MISSILE):
        return 21
    if obs_kind == ts.OBS_MULTI:
        return 9 + (n_agents - 1) * 6           # E/tasks/multiplecombat_task.py:96-99
    if obs_kind == ts.OBS_MULTI_MISSILE:
        return 9 + n_agents * 6                 # E/tasks/multiplecombat_task.py:213-217
    if obs_kind == ts.OBS_NVN:                  # E/tasks/scenario2_task.py:246-252: len(aircraft_configs)/2 for partners AND enemies
        return int((11 if rwr else 9) + 6 * (n_aircraft_cfg / 2) + 6 * (n_aircraft_cfg / 2) + 6)   # *_rwr: num_ego_obs = 11 (:402)
    raise ValueError(obs_kind)


def build_spec(cfg: dict, substeps_override=None) -> TaskSpec:
    name = cfg.get("task")
    if name in ("scenario1_for_KAI", "scenario2_for_KAI", "scenario3_for_KAI"):
        raise NotImplementedError(f"{name}: the KAI project tasks (hard-wired Korean-peninsula way-points and a socket link, "
                                  "E/tasks/KAI_project_task.py) are outside the env-step scope (DESIGN.md section 8)")
    if name not in TASKS:
        raise NotImplementedError(f"Unknown taskname: {name}")
    t = TASKS[name]
    acs = cfg["aircraft_configs"]
    uids = list(acs.keys())
    team0 = uids[0][0]
    ego = [u for u in uids if u[0] == team0]
    enm = [u for u in uids if u[0] != team0]
    order = ego + enm                            # BaseEnv._pack: ego first, then enemies (E/envs/env_base.py:269-283)
    if order != uids:
        raise NotImplementedError("aircraft_configs must list the ego team first (dict order == pack order in every reference yaml)")
    sp = TaskSpec(name=name)
    sp.n_ego, sp.n_enm = len(ego), len(enm)
    sp.sim_freq = int(cfg.get("sim_freq", 60))
    sp.substeps = int(substeps_override if substeps_override else cfg.get("agent_interaction_steps", 12))
    sp.max_steps = int(cfg.get("max_steps", 100 if t["env"] != "control" else 100))
    sp.altitude_limit = float(cfg.get("altitude_limit", 2500))
    sp.acc_limit = (float(cfg.get("acceleration_limit_x", 10.0)), float(cfg.get("acceleration_limit_y", 10.0)),
                    float(cfg.get("acceleration_limit_z", 10.0)))
    sp.center = tuple(float(x) for x in cfg.get("battle_field_center", (120.0, 60.0, 0.0)))
    sp.obs_kind = t["obs"]
    sp.obs_dim = obs_dim_for(sp.obs_kind, sp.n_agents, len(uids), bool(t.get("rwr")))
    sp.act_kind = ts.ACT_HEADING if t["env"] == "control" else ts.ACT_COMBAT
    sp.shoot_dim = t["shoot"]
    sp.rewards = [_reward(cfg, c) for c in t["rewards"]]
    sp.use_artillery = bool(cfg.get("use_artillery", False))
    sp.use_baseline = bool(cfg.get("use_baseline", False))
    sp.baseline_type = str(cfg.get("baseline_type", "")) if sp.use_baseline else ""
    sp.launch_kind = t["launch"]
    sp.max_attack_angle = float(cfg.get("max_attack_angle", 180))
    sp.max_attack_distance = float(cfg.get("max_attack_distance", math.inf))
    sp.min_attack_interval = int(cfg.get("min_attack_interval", 125))
    sp.num_missiles = [int(acs[u].get("missile", 0)) for u in order]
    sp.init_states = [init_state_row(acs[u].get("init_state")) for u in order]
    sp.yaml_init_states = [list(r) for r in sp.init_states]
    sp.env_kind, sp.curriculum = t["env"], bool(t.get("curriculum"))
    if sp.curriculum:       # every reset of these tasks goes through reset_simulators_curriculum(curriculum_angle), stage 0 first
        sp.init_states = curriculum_init_states(t["env"], sp.yaml_init_states, 0)
        # the win-rate record of Scenario*_curriculum / WVRTask / Maneuver_curriculum, with the reference's (inert) stage rule
        sp.curriculum_rule, sp.curriculum_window = 1, 20
        sp.curriculum_threshold = 0.9 if t["env"] == "1v1" else 0.6
    if t["env"] == "control":
        sp.terminations = list(t.get("terms", _TERMS_HEADING))
        sp.dones_before_rewards, sp.team_mean, sp.share_obs, sp.reward_gate = True, False, False, ts.G_NONE
        a0 = acs[uids[0]]
        if ts.T_UNREACH_HEADING in sp.terminations:
            sp.heading_increments = (float(a0["max_heading_increment"]), float(a0["max_altitude_increment"]),
                                     float(a0["max_velocities_u_increment"]))
            sp.check_interval = float(a0["check_interval"])
    elif t["env"] == "1v1":
        if sp.n_agents != 2:
            raise AssertionError("SingleCombatEnv only supports 1v1 scenarios!")   # E/envs/singlecombat_env.py:17
        sp.terminations = list(t.get("terms", _TERMS_1V1))
        sp.dones_before_rewards, sp.team_mean, sp.share_obs, sp.reward_gate = True, False, False, ts.G_DIE_FLAG
    else:
        sp.terminations = list(_TERMS_NVN)
        sp.dones_before_rewards, sp.team_mean, sp.share_obs, sp.reward_gate = False, True, True, ts.G_ALIVE
    return sp


def load_spec(config_name: str, config_dir=None, substeps_override=None) -> TaskSpec:
    return build_spec(parse_config(config_name, config_dir), substeps_override)
