"""TaskSpec: the flat description of an env/task that BOTH the CUDA step kernels (packed into the
``AcsTaskConfig`` C struct, include/acs.h) and the CPU oracle (oracle/env_oracle.py) execute.

It is what the reference's Task classes boil down to once Python class composition is resolved
(reference envs/JSBSim/tasks/*.py): which observation packer, which reward functions in which order
with which hyper-parameters, which termination conditions in which order, which weapon-launch rule.
The host-side Task mirror (aircombat_selfplay_b200/envs/tasks.py) builds it from the yaml config.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

# observation packers
OBS_HEADING = 0        # heading_task.py:67-100 (12, clipped)
OBS_1V1 = 1            # singlecombat_task.py:88-139 (15, 2-D AO/TA, clipped)
OBS_1V1_MISSILE = 2    # singlecombat_with_missile_task.py:31-99 (21, 3-D, unclipped, enemy = enemies[0])
OBS_NV_MISSILE = 3     # multiplecombat_with_missile_task.py:32-117 (21, enemy index = own index in team)
OBS_MULTI = 4          # multiplecombat_task.py:105-135 (9+6(A-1), clipped)
OBS_MULTI_MISSILE = 5  # multiplecombat_task.py:232-267 (9+6A, clipped then missile block)
OBS_NVN = 6            # scenario2_task.py:256-316 (9 + 6*partners + 6*enemies + 6, unclipped); *_rwr: two trailing zeros more
OBS_1V1_RWR = 7        # scenario1_task.py:222-314 (23: ego 9 + nearest live enemy 6; the missile block and RWR pair stay 0)

# action normalisation
ACT_HEADING = 0        # heading_task.py:102-110
ACT_COMBAT = 1         # singlecombat_task.py:141-153

# reward kinds (reference envs/JSBSim/reward_functions/*.py)
R_ALTITUDE, R_POSTURE, R_EVENT, R_MISSILE_POSTURE, R_SHOOT_PENALTY, R_HEADING, R_RELATIVE_ALTITUDE, \
    R_COMBAT_GEOMETRY, R_GUN_BEHIT, R_GUN_TARGETTAIL, R_GUN_WEZ, R_GUN_WEZDOT = range(12)
REWARD_CLASS = {"AltitudeReward": R_ALTITUDE, "PostureReward": R_POSTURE, "EventDrivenReward": R_EVENT,
                "MissilePostureReward": R_MISSILE_POSTURE, "ShootPenaltyReward": R_SHOOT_PENALTY,
                "HeadingReward": R_HEADING, "RelativeAltitudeReward": R_RELATIVE_ALTITUDE,
                "CombatGeometryReward": R_COMBAT_GEOMETRY, "GunBEHITReward": R_GUN_BEHIT,
                "GunTargetTailReward": R_GUN_TARGETTAIL, "GunWEZReward": R_GUN_WEZ, "GunWEZDOTReward": R_GUN_WEZDOT}

# termination kinds (reference envs/JSBSim/termination_conditions/*.py)
T_UNREACH_HEADING, T_EXTREME_STATE, T_OVERLOAD, T_LOW_ALTITUDE, T_TIMEOUT, T_SAFE_RETURN = range(6)

# weapon launch rules (task.step)
L_NONE = 0       # no launches (SingleCombatTask, MultipleCombatTask, MultipleCombatShootMissileTask.step)
L_RULE_LOCK = 1  # *DodgeMissileTask.step: lock-duration deque, AIM-9L (singlecombat_with_missile_task.py:109-127)
L_RL_SINGLE = 2  # SingleCombatShootMissileTask.step (:194-204), target enemies[0], AIM-9L
L_RL_NEAREST = 3 # multiplecombat_task.py:278-299, nearest enemy, angle/distance/interval gates, AIM-9L
L_SCENARIO = 4   # scenario{1,2,3}_task.py step: gun / AIM-120B / AIM-9M / chaff
L_AUTO_GUN = 5   # WVRTask.step (WVR_task.py:67-81) / Maneuver_curriculum.step (singlecombat_task.py:290-297): the gun fires by
                 # itself whenever the farthest enemy is within 3 km and 5 degrees: -5 blood, no ammunition, no alive checks

# reward gating (who gets a reward this step)
G_NONE = 0       # BaseTask.get_reward (heading task)
G_DIE_FLAG = 1   # singlecombat_task.py:190-195
G_ALIVE = 2      # multiplecombat_task.py:147-151

MAX_AGENTS = 8
MAX_REWARDS = 12
MAX_TERMS = 6


@dataclass
class RewardSpec:
    kind: int
    scale: float = 1.0
    potential: bool = False
    p0: float = 0.0   # Altitude: safe_altitude | Posture: orientation version | RelativeAltitude: KH
    p1: float = 0.0   # Altitude: danger_altitude | Posture: range version
    p2: float = 0.0   # Altitude: Kv | Posture: target_dist


@dataclass
class TaskSpec:
    name: str = "task"
    n_ego: int = 1
    n_enm: int = 0
    sim_freq: int = 60
    substeps: int = 12                      # agent_interaction_steps (env_base.py:27)
    max_steps: int = 100
    altitude_limit: float = 2500.0
    acc_limit: Tuple[float, float, float] = (10.0, 10.0, 10.0)
    center: Tuple[float, float, float] = (120.0, 60.0, 0.0)
    obs_kind: int = OBS_HEADING
    obs_dim: int = 12
    act_kind: int = ACT_HEADING
    shoot_dim: int = 0                      # trailing shoot entries in the action: 0, 1 or 4
    rewards: List[RewardSpec] = field(default_factory=list)
    terminations: List[int] = field(default_factory=list)
    dones_before_rewards: bool = True       # BaseEnv.step order (env_base.py:159-171) vs MultipleCombatEnv (:163-180)
    team_mean: bool = False                 # multiplecombat_env.py:170-175
    share_obs: bool = False
    reward_gate: int = G_NONE
    launch_kind: int = L_NONE
    use_artillery: bool = False
    use_baseline: bool = False              # scripted red team (opponents.py); also narrows its AIM-120B cone (scenario1_task.py:137-140)
    baseline_type: str = ""                 # yaml baseline_type: pursue | maneuver (| loiter: NotImplementedError as in the reference)
    max_attack_angle: float = 180.0
    max_attack_distance: float = float("inf")
    min_attack_interval: int = 125
    num_missiles: List[int] = field(default_factory=list)       # per aircraft (yaml `missile`)
    init_states: List[List[float]] = field(default_factory=list)  # per aircraft [lon,lat,h_ft,psi,u,v,w,p,q,r,phi,theta]
    # heading task (unreach_heading.py:16-21, singlecontrol_env.py:35-46)
    heading_increments: Tuple[float, float, float] = (180.0, 7000.0, 100.0)
    check_interval: float = 30.0
    fcs_dt: float = 1.0 / 120.0
    # curriculum tasks (Scenario*_curriculum, WVRTask, Maneuver_curriculum): resets use the curriculum circle geometry
    env_kind: str = "control"
    curriculum: bool = False
    yaml_init_states: List[List[float]] = field(default_factory=list)
    # per-env win-rate record + stage rule, evaluated on the device when an episode ends (include/acs.h AcsTaskConfig):
    # window 20 episodes; threshold 0.9 for the 1v1 curriculum tasks (scenario1_task.py:169, WVR_task.py:42), 0.6 for the
    # N-v-N ones (scenario2_task.py:184); rule 1 = the reference's `len(record) > 20`, which never fires (the record is
    # capped at 20), 2 = advance when the record is full, 0 = record only
    curriculum_rule: int = 0
    curriculum_window: int = 0
    curriculum_threshold: float = 0.0

    @property
    def n_agents(self):
        return self.n_ego + self.n_enm

    @property
    def dt(self):
        return 1.0 / self.sim_freq

    @property
    def time_interval(self):
        return self.substeps / self.sim_freq

    @property
    def lock_len(self):
        return int(1 / self.time_interval)

    @property
    def n_missile_slots(self):
        """Missile slots per aircraft = the most launches one aircraft can make in an episode.  Scenario tasks keep
        separate AIM-9M and AIM-120B counters, each initialised to ``num_missiles`` (scenario2_task.py:66-69)."""
        if self.launch_kind == L_NONE:
            return 0
        n = max([0] + list(self.num_missiles))
        return 2 * n if self.launch_kind == L_SCENARIO else n
