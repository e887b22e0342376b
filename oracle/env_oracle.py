"""CPU ORACLE for the env-step layer -- TEST INFRASTRUCTURE ONLY (see oracle/f16_oracle.cpp header).

PARITY UNPINNED for the FDM (no runnable JSBSim here); this layer restates the reference's *Python*
env/task/reward/termination/missile/chaff code, which is available in full under
/root/reference/envs/JSBSim ("E/" below), function by function, in single-env scalar form.  It drives
the C++ FDM oracle (oracle/fdm.py) the way ``AircraftSimulator`` drives ``jsbsim.FGFDMExec``.

Deliberate, documented deviations from the reference (they cannot be matched by any implementation):
  * random draws (E/envs/env_base.py:153 global numpy RNG; E/envs/singlecontrol_env.py:35-37 and
    E/termination_conditions/unreach_heading.py:45-47 gymnasium PCG64) are replaced by the keyed
    counter-based generator ``u01`` below, which the CUDA kernels implement identically;
  * pymap3d (unpinned third-party dependency, E/utils/utils.py:40,54) is restated from its published
    algorithm (geodetic2ecef + ecef2enu rotation; ecef2geodetic after You (2000));
  * Scenario1's enemy shoot-action TypeError is resolved with Scenario2's rule.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np

from aircombat_selfplay_b200 import taskspec as ts
from oracle.fdm import OracleFdm

ALIVE, CRASH, SHOTDOWN = 0, 1, 2
M_INACTIVE, M_LAUNCHED, M_HIT, M_MISS = -1, 0, 1, 2

MASK64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def u01(seed, env, purpose, a=0, b=0, c=0):
    """Keyed uniform in [0,1): hash chain over (seed, env, purpose, a, b, c)."""
    h = _splitmix64(seed & MASK64)
    for v in (env, purpose, a, b, c):
        h = _splitmix64(h ^ (int(v) & MASK64))
    return (h >> 11) * (1.0 / 9007199254740992.0)


RNG_RESET, RNG_HEADING, RNG_CHAFF = 1, 2, 3

# ----------------------------------------------------------------------------- pymap3d restatement (WGS-84)
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_B = WGS84_A * (1.0 - WGS84_F)


def geodetic2ecef(lat_deg, lon_deg, alt):
    lat, lon = math.radians(lat_deg), math.radians(lon_deg)
    n = WGS84_A ** 2 / math.hypot(WGS84_A * math.cos(lat), WGS84_B * math.sin(lat))
    x = (n + alt) * math.cos(lat) * math.cos(lon)
    y = (n + alt) * math.cos(lat) * math.sin(lon)
    z = (n * (WGS84_B / WGS84_A) ** 2 + alt) * math.sin(lat)
    return x, y, z


def LLA2NEU(lon, lat, alt, lon0, lat0, alt0):
    """E/utils/utils.py:30-41 -> pymap3d.geodetic2ned, returned as (north, east, up)."""
    x1, y1, z1 = geodetic2ecef(lat, lon, alt)
    x2, y2, z2 = geodetic2ecef(lat0, lon0, alt0)
    u, v, w = x1 - x2, y1 - y2, z1 - z2
    la, lo = math.radians(lat0), math.radians(lon0)
    t = math.cos(lo) * u + math.sin(lo) * v
    east = -math.sin(lo) * u + math.cos(lo) * v
    up = math.cos(la) * t + math.sin(la) * w
    north = -math.sin(la) * t + math.cos(la) * w
    return np.array([north, east, up])


def NEU2alt(n, e, u, lon0, lat0, alt0):
    """Altitude component of E/utils/utils.py:44-55 (pymap3d.ned2geodetic): enu2ecef, then ecef2geodetic
    (You 2000).  Only the altitude is consumed on the hot path (missile air density)."""
    la, lo = math.radians(lat0), math.radians(lon0)
    x0, y0, z0 = geodetic2ecef(lat0, lon0, alt0)
    t = math.cos(la) * u - math.sin(la) * n
    w = math.sin(la) * u + math.cos(la) * n
    uu = math.cos(lo) * t - math.sin(lo) * e
    vv = math.sin(lo) * t + math.cos(lo) * e
    x, y, z = x0 + uu, y0 + vv, z0 + w
    a, b = WGS84_A, WGS84_B
    r = math.sqrt(x * x + y * y + z * z)
    E = math.sqrt(a * a - b * b)
    uq = math.sqrt(0.5 * (r * r - E * E) + 0.5 * math.hypot(r * r - E * E, 2 * E * z))
    Q = math.hypot(x, y)
    huE = math.hypot(uq, E)
    beta = math.atan(huE / uq * z / Q)
    dbeta = ((b * uq - a * huE + E * E) * math.sin(beta)) / (a * huE / math.cos(beta) - E * E * math.cos(beta))
    beta += dbeta
    alt = math.hypot(z - b * math.sin(beta), Q - a * math.cos(beta))
    inside = x * x / (a * a) + y * y / (a * a) + z * z / (b * b) < 1
    return -alt if inside else alt


def get_AO_TA_R(ego, enm, two_d=False):
    """E/utils/utils.py:58-103.  Features are (north, east, up, v_north, v_east, v_down)."""
    ex, ey, ez, evx, evy, evz = ego
    tx, ty, tz, tvx, tvy, tvz = enm
    dx, dy, dz = tx - ex, ty - ey, tz - ez
    if two_d:
        ego_v = math.sqrt(evx * evx + evy * evy)
        enm_v = math.sqrt(tvx * tvx + tvy * tvy)
        R = math.sqrt(dx * dx + dy * dy)
        p1 = dx * evx + dy * evy
        p2 = dx * tvx + dy * tvy
    else:
        ego_v = math.sqrt(evx * evx + evy * evy + evz * evz)
        enm_v = math.sqrt(tvx * tvx + tvy * tvy + tvz * tvz)
        R = math.sqrt(dx * dx + dy * dy + dz * dz)
        p1 = dx * evx + dy * evy + dz * evz
        p2 = dx * tvx + dy * tvy + dz * tvz
    AO = math.acos(min(1.0, max(-1.0, p1 / (R * ego_v + 1e-8))))
    TA = math.acos(min(1.0, max(-1.0, p2 / (R * enm_v + 1e-8))))
    cr = evx * dy - evy * dx
    side = int(cr > 0) - int(cr < 0)
    return AO, TA, R, float(side)


def _clip(v, lo, hi):
    return lo if v < lo else (hi if v > hi else v)


# ----------------------------------------------------------------------------- simulators
MISSILE_PARAMS = {
    # E/core/simulatior.py:420-433 (base class, AIM-9L) and :663-675 / :700-712 (both subclasses carry AIM-120B numbers)
    0: dict(g=9.81, t_max=60.0, t_thrust=3.0, Isp=120.0, Length=2.87, Diameter=0.127, cD=0.4, m0=84.0, dm=6.0, K=3.0,
            nyz_max=30.0, Rc=300.0, v_min=150.0),
    1: dict(g=9.81, t_max=27.22, t_thrust=1.4, Isp=1837.0, Length=3.66, Diameter=0.18, cD=0.02, m0=152.0, dm=6.0, K=5.0,
            nyz_max=50.0, Rc=5.0, v_min=150.0),
}


class Aircraft:
    """E/core/simulatior.py:89-325 AircraftSimulator over the oracle FDM."""

    def __init__(self, idx, spec: ts.TaskSpec):
        self.idx = idx
        self.spec = spec
        self.fdm = OracleFdm(spec.dt, spec.fcs_dt)
        self.num_missiles = spec.num_missiles[idx] if spec.num_missiles else 0
        self.partners: List["Aircraft"] = []
        self.enemies: List["Aircraft"] = []
        self.launch_missiles: List["Missile"] = []
        self.under_missiles: List["Missile"] = []
        self.extra = {}
        self.status = ALIVE
        self.bloods = 100.0

    is_alive = property(lambda s: s.status == ALIVE)
    is_crash = property(lambda s: s.status == CRASH)
    is_shotdown = property(lambda s: s.status == SHOTDOWN)

    def reload(self, ic):
        self.bloods = 100.0
        self.status = ALIVE
        self.launch_missiles = []
        self.under_missiles = []
        self.fdm.reset(*ic)
        self.controls = [0.0, 0.0, 0.0, 0.0]
        self._update()

    def _update(self):
        s = self.fdm.snapshot_dict()
        self.s = s
        c = self.spec.center
        # ExtraCatalog derived props are clipped on set (E/core/catalog.py:292-338, simulatior.py:308-311)
        self.h_sl_m = _clip(s["h_sl_ft"] * 0.3048, -500.0, 26000.0)
        self.geodetic = np.array([s["lon_deg"], s["lat_geod_deg"], self.h_sl_m])
        self.position = LLA2NEU(s["lon_deg"], s["lat_geod_deg"], self.h_sl_m, c[0], c[1], c[2])
        self.posture = np.array([s["roll_rad"], s["pitch_rad"], s["heading_rad"]])
        self.velocity = np.array([_clip(s["v_north_fps"] * 0.3048, -700.0, 700.0), _clip(s["v_east_fps"] * 0.3048, -700.0, 700.0),
                                  _clip(s["v_down_fps"] * 0.3048, -700.0, 700.0)])
        self.uvw_mps = [_clip(s["u_fps"] * 0.3048, -700.0, 700.0), _clip(s["v_fps"] * 0.3048, -700.0, 700.0),
                        _clip(s["w_fps"] * 0.3048, -700.0, 700.0)]
        self.vc_mps = _clip(s["vc_fps"] * 0.3048, 0.0, 1400.0)

    def set_controls(self, u):
        lo, hi = (-1.0, -1.0, -1.0, 0.0), (1.0, 1.0, 1.0, 0.9)  # E/core/catalog.py:192-197
        self.controls = [_clip(float(u[k]), lo[k], hi[k]) for k in range(4)]
        self.fdm.set_controls(*self.controls)

    def run(self):
        if self.is_alive:
            if self.bloods <= 0:
                self.status = SHOTDOWN
            self.fdm.run(1)
            self._update()

    def check_missile_warning(self):
        for m in self.under_missiles:
            if m.is_alive:
                return m
        return None


class Missile:
    """E/core/simulatior.py:393-608 MissileSimulator (param set 0) / AIM_9M, AIM_120B (param set 1)."""

    def __init__(self, parent: Aircraft, target: Aircraft, kind: int, dt: float, center):
        self.pr = MISSILE_PARAMS[kind]
        self.kind = kind
        self.dt = dt
        self.center = center
        self.parent, self.target = parent, target
        parent.launch_missiles.append(self)
        target.under_missiles.append(self)
        self.position = parent.position.copy()
        self.velocity = parent.velocity.copy()
        self.posture = parent.posture.copy()
        self.posture[0] = 0
        self.alt = parent.geodetic[2]
        self.t = 0.0
        self.m = self.pr["m0"]
        self.dtheta = self.dphi = 0.0
        self.status = M_LAUNCHED
        self.distance_pre = math.inf
        self.maxlen = int(5 / dt)
        self.consec = 0  # == sum(deque) >= maxlen  <=>  the last maxlen appended flags were all True

    is_alive = property(lambda s: s.status == M_LAUNCHED)
    is_success = property(lambda s: s.status == M_HIT)
    is_done = property(lambda s: s.status in (M_HIT, M_MISS))

    @property
    def target_distance(self):
        return float(np.linalg.norm(self.target.position - self.position))

    def run(self):
        pr = self.pr
        self.t += self.dt
        action, distance = self._guidance()
        self.consec = self.consec + 1 if distance > self.distance_pre else 0
        self.distance_pre = distance
        if distance < pr["Rc"] and self.target.is_alive and self.status != M_MISS:
            self.status = M_HIT
            self.target.status = SHOTDOWN
        elif (self.t > pr["t_max"]) or (float(np.linalg.norm(self.velocity)) < pr["v_min"]) or self.consec >= self.maxlen \
                or not self.target.is_alive:
            self.status = M_MISS
        else:
            self._state_trans(action)

    def _guidance(self):
        pr = self.pr
        x_m, y_m, z_m = self.position
        dx_m, dy_m, dz_m = self.velocity
        v_m = math.sqrt(dx_m * dx_m + dy_m * dy_m + dz_m * dz_m)
        theta_m = math.asin(dz_m / v_m)
        x_t, y_t, z_t = self.target.position
        dx_t, dy_t, dz_t = self.target.velocity
        Rxy = math.sqrt((x_m - x_t) ** 2 + (y_m - y_t) ** 2)
        Rxyz = math.sqrt((x_m - x_t) ** 2 + (y_m - y_t) ** 2 + (z_t - z_m) ** 2)
        dbeta = ((dy_t - dy_m) * (x_t - x_m) - (dx_t - dx_m) * (y_t - y_m)) / Rxy ** 2
        deps = ((dz_t - dz_m) * Rxy ** 2 - (z_t - z_m) * ((x_t - x_m) * (dx_t - dx_m) + (y_t - y_m) * (dy_t - dy_m))) / (Rxyz ** 2 * Rxy)
        K = max(pr["K"] * (pr["t_max"] - self.t) / pr["t_max"], 0)
        ny = K * v_m / pr["g"] * math.cos(theta_m) * dbeta
        nz = K * v_m / pr["g"] * deps + math.cos(theta_m)
        return (_clip(ny, -pr["nyz_max"], pr["nyz_max"]), _clip(nz, -pr["nyz_max"], pr["nyz_max"])), Rxyz

    def _state_trans(self, action):
        pr = self.pr
        self.position = self.position + self.dt * self.velocity
        self.alt = NEU2alt(self.position[0], self.position[1], self.position[2], *self.center)
        v = float(np.linalg.norm(self.velocity))
        theta, phi = self.posture[1], self.posture[2]
        Isp = pr["Isp"] if self.t < pr["t_thrust"] else 0
        T = pr["g"] * Isp * pr["dm"]
        S0 = math.pi * (pr["Diameter"] / 2) ** 2
        S0 += math.sqrt(math.sin(self.dtheta) ** 2 + math.sin(self.dphi) ** 2) * pr["Diameter"] * pr["Length"]
        rho = 1.225 * math.exp(-self.alt / 9300)
        D = 0.5 * pr["cD"] * S0 * rho * v ** 2
        nx = (T - D) / (self.m * pr["g"])
        ny, nz = action
        dv = pr["g"] * (nx - math.sin(theta))
        self.dphi = pr["g"] / v * (ny / math.cos(theta))
        self.dtheta = pr["g"] / v * (nz - math.cos(theta))
        v += self.dt * dv
        phi += self.dt * self.dphi
        theta += self.dt * self.dtheta
        self.velocity = np.array([v * math.cos(theta) * math.cos(phi), v * math.cos(theta) * math.sin(phi), v * math.sin(theta)])
        self.posture = np.array([0, theta, phi])
        if self.t < pr["t_thrust"]:
            self.m = self.m - self.dt * pr["dm"]


class Chaff:
    """E/core/simulatior.py:327-391.  ``count`` chaffs created in the same task.step share position and clock."""

    def __init__(self, parent: Aircraft, dt, count):
        self.position = parent.position.copy()
        self.dt = dt
        self.t = 0.0
        self.count = count
        self.done = False
        self.parent = parent

    def run(self):
        self.t += self.dt
        if self.t > 20:
            self.done = True


# ----------------------------------------------------------------------------- the environment
class OracleEnv:
    """One environment.  reset()/step() follow E/envs/env_base.py:98-173 (dones before rewards) or
    E/envs/multiplecombat_env.py:66-182 (rewards, team mean, then dones) as TaskSpec selects."""

    def __init__(self, spec: ts.TaskSpec, seed: int = 0, env_index: int = 0):
        self.spec = spec
        self.seed = seed
        self.env_index = env_index
        A = spec.n_agents
        self.sims = [Aircraft(i, spec) for i in range(A)]
        for i, s in enumerate(self.sims):
            for j, o in enumerate(self.sims):
                if i == j:
                    continue
                same = (i < spec.n_ego) == (j < spec.n_ego)
                (s.partners if same else s.enemies).append(o)
        self.episode = -1
        self.init_states = [list(x) for x in spec.init_states]

    # ------------------------------------------------------------------ reset
    def reset(self):
        sp = self.spec
        self.episode += 1
        self.current_step = 0
        self.substep_count = 0
        ics = [list(x) for x in self.init_states]
        if sp.obs_kind == ts.OBS_HEADING:
            # E/envs/singlecontrol_env.py:32-49
            self.heading_turn_counts = 0
            h = 0.0 + 180.0 * u01(self.seed, self.env_index, RNG_RESET, self.episode, 0)
            alt = 14000.0 + 16000.0 * u01(self.seed, self.env_index, RNG_RESET, self.episode, 1)
            vu = 400.0 + 800.0 * u01(self.seed, self.env_index, RNG_RESET, self.episode, 2)
            ics[0][3], ics[0][2], ics[0][4] = h, alt, vu
            self.target_heading_deg = _clip(h, 0.0, 360.0)
            self.target_altitude_ft = _clip(alt, -1400.0, 85000.0)
            self.target_velocities_u_mps = _clip(vu * 0.3048, -700.0, 700.0)
            self.heading_check_time = 0.0
        for s, ic in zip(self.sims, ics):
            s.reload(ic)
        self.missiles = {}      # key (agent, n) -> Missile, python-dict insertion order == run order
        self.chaffs = []        # Chaff groups in creation order
        self.detached = []
        # task.reset
        A = sp.n_agents
        self.die_flag = [False] * A
        self.shoot_action = [[0, 0, 0, 0] for _ in range(A)]
        self.remaining_missiles = [s.num_missiles for s in self.sims]
        self.remaining_9m = [s.num_missiles for s in self.sims]
        self.remaining_120b = [s.num_missiles for s in self.sims]
        self.remaining_gun = [s.num_missiles for s in self.sims]
        self.remaining_chaff = [s.num_missiles for s in self.sims]
        self.last_shoot_time = [-sp.min_attack_interval] * A
        self.last_shot_missile = [None] * A
        self.last_shot_chaff = [None] * A
        self.lock = [[] for _ in range(A)]
        # reward_function.reset (E/reward_functions/reward_function_base.py:20-32 + subclasses)
        self.pre_rewards = [[0.0] * A for _ in sp.rewards]
        self.prev_missile_v = None
        self.cg_prev = None
        self.tt_prev = None
        self.wd_prev = None
        self.pre_remaining = [s.num_missiles for s in self.sims]
        self.hr_last = [None] * A
        for ri, r in enumerate(sp.rewards):
            if r.kind == ts.R_MISSILE_POSTURE:
                self.prev_missile_v = None
            if r.kind == ts.R_COMBAT_GEOMETRY:
                self.cg_prev = None
            if r.kind == ts.R_GUN_TARGETTAIL:
                self.tt_prev = None
            if r.kind == ts.R_GUN_WEZDOT:
                self.wd_prev = None
            if r.potential:
                for a in range(A):
                    self.pre_rewards[ri][a] = self._reward_one(ri, a)
        obs = [self.get_obs(a) for a in range(A)]
        return np.array(obs), self._share(obs)

    def _share(self, obs):
        if not self.spec.share_obs:
            return None
        st = np.hstack(obs)
        return np.array([st.copy() for _ in range(self.spec.n_agents)])

    # ------------------------------------------------------------------ step
    def step(self, action):
        """action: int array [A, 4 + shoot_dim] of LOW-LEVEL discrete actions (the hierarchical GRU controller
        sits above this boundary)."""
        sp = self.spec
        A = sp.n_agents
        self.current_step += 1
        info = {"current_step": self.current_step, "done_cause": [-1] * A}
        for a in range(A):
            act = action[a]
            if sp.act_kind == ts.ACT_HEADING:
                u = [act[0] * 2. / (41 - 1.) - 1., act[1] * 2. / (41 - 1.) - 1., act[2] * 2. / (41 - 1.) - 1., act[3] * 0.5 / (30 - 1.) + 0.4]
            else:
                u = [act[0] / 20 - 1., act[1] / 20 - 1., act[2] / 20 - 1., act[3] / 58 + 0.4]
            if sp.shoot_dim == 1:
                self.shoot_action[a] = [int(act[4]), 0, 0, 0]
            elif sp.shoot_dim == 4:
                self.shoot_action[a] = [int(x) for x in act[4:8]]
            self.sims[a].set_controls(u)
        for _ in range(sp.substeps):
            for s in self.sims:
                s.run()
            for m in list(self.missiles.values()):
                m.run()
            for c in self.chaffs:
                c.run()
            for key, m in self.missiles.items():
                if m.is_done:
                    continue
                for c in self.chaffs:
                    if c.done:
                        continue
                    if float(np.linalg.norm(c.position - m.position)) <= 300:
                        for j in range(c.count):
                            # the episode is part of the key: substep counters restart at every reset
                            if u01(self.seed, self.env_index, RNG_CHAFF, (self.episode << 20) + self.substep_count, key[0] * 64 + key[1],
                                   c.parent.idx * 64 + j) < 0.85:
                                m.status = M_MISS
            self.substep_count += 1
        self._task_step()
        obs = [self.get_obs(a) for a in range(A)]
        share = self._share(obs)
        dones = [False] * A
        rewards = [0.0] * A
        if sp.dones_before_rewards:
            for a in range(A):
                dones[a] = self._termination(a, info)
            for a in range(A):
                rewards[a] = self._get_reward(a)
        else:
            for a in range(A):
                rewards[a] = self._get_reward(a)
            if sp.team_mean:
                ego = float(np.mean([rewards[a] for a in range(sp.n_ego)]))
                enm = float(np.mean([rewards[a] for a in range(sp.n_ego, A)]))
                rewards = [ego] * sp.n_ego + [enm] * sp.n_enm
            for a in range(A):
                dones[a] = self._termination(a, info)
        if sp.obs_kind == ts.OBS_HEADING:
            info["heading_turn_counts"] = self.heading_turn_counts
        return np.array(obs), share, np.array(rewards), np.array(dones), info

    # ------------------------------------------------------------------ task.step: weapons
    def _attack_geometry(self, agent, enemy):
        target = enemy.position - agent.position
        heading = agent.velocity
        distance = float(np.linalg.norm(target))
        ang = math.degrees(math.acos(_clip(float(np.sum(target * heading)) / (distance * float(np.linalg.norm(heading)) + 1e-8), -1, 1)))
        return distance, ang

    def _launch(self, a, target, kind, n):
        agent = self.sims[a]
        m = Missile(agent, target, kind, self.spec.dt, self.spec.center)
        key = (a, n)
        if key in self.missiles:
            self.detached.append(self.missiles[key])
        self.missiles[key] = m  # an existing key keeps its position in the dict (E/envs/env_base.py:90-92)
        return m

    def _get_target(self, agent):  # scenario tasks: argmax distance (E/tasks/scenario2_task.py:150-156)
        d = [float(np.linalg.norm(e.position - agent.position)) for e in agent.enemies]
        return agent.enemies[int(np.argmax(d))]

    def _a2a_available(self, a):  # E/tasks/scenario2_task.py:116-148
        agent = self.sims[a]
        ret = [False, False, False]
        enemy = self._get_target(agent)
        if not enemy.is_alive:
            return ret, enemy
        distance, ang = self._attack_geometry(agent, enemy)
        if distance / 1000 < 3 and ang < 5:
            ret[0] = True
        if distance / 1000 < 37 and ang < 90:
            ret[1] = True
        if distance / 1000 < 7 and ang < 90:
            ret[2] = True
        if self.spec.use_baseline and a >= self.spec.n_ego:
            ret[1] = False
            if distance / 1000 < 37 and ang < 90 / 2:
                ret[1] = True
        return ret, enemy

    def _task_step(self):
        sp = self.spec
        A = sp.n_agents
        if sp.use_artillery:  # E/tasks/singlecombat_task.py:163-188
            for a in range(A):
                agent = self.sims[a]
                ego_f = list(agent.position) + list(agent.velocity)
                for enm in agent.enemies:
                    if enm.is_alive:
                        AO, _, R, _ = get_AO_TA_R(ego_f, list(enm.position) + list(enm.velocity))
                        if 0 <= AO <= 0.5236:
                            o = 1 - AO / 0.5236
                        elif -0.5236 <= AO <= 0:
                            o = 1 + AO / 0.5236
                        else:
                            o = 0
                        Rk = R / 1000
                        dfn = 1 if Rk <= 1 else ((3 - Rk) / 2. if Rk <= 3 else 0)
                        enm.bloods -= o * dfn
        if sp.launch_kind == ts.L_NONE:
            return
        for a in range(A):
            agent = self.sims[a]
            if sp.launch_kind == ts.L_RULE_LOCK:  # E/tasks/singlecombat_with_missile_task.py:109-127
                distance, ang = self._attack_geometry(agent, agent.enemies[0])
                self.lock[a].append(ang < sp.max_attack_angle)
                self.lock[a] = self.lock[a][-sp.lock_len:] if sp.lock_len > 0 else []
                interval = self.current_step - self.last_shoot_time[a]
                flag = agent.is_alive and sum(self.lock[a]) >= sp.lock_len and distance <= sp.max_attack_distance \
                    and self.remaining_missiles[a] > 0 and interval >= sp.min_attack_interval
                if flag:
                    self._launch(a, agent.enemies[0], 0, self.remaining_missiles[a])
                    self.remaining_missiles[a] -= 1
                    self.last_shoot_time[a] = self.current_step
            elif sp.launch_kind == ts.L_RL_SINGLE:  # :194-204
                flag = agent.is_alive and self.shoot_action[a][0] and self.remaining_missiles[a] > 0
                if flag and (self.last_shot_missile[a] is None or self.last_shot_missile[a].is_done):
                    self.last_shot_missile[a] = self._launch(a, agent.enemies[0], 0, self.remaining_missiles[a])
                    self.remaining_missiles[a] -= 1
            elif sp.launch_kind == ts.L_RL_NEAREST:  # E/tasks/multiplecombat_task.py:278-299
                dists = [float(np.linalg.norm(e.position - agent.position)) for e in agent.enemies]
                ti = int(np.argmin(dists))
                distance, ang = self._attack_geometry(agent, agent.enemies[ti])
                interval = self.current_step - self.last_shoot_time[a]
                flag = agent.is_alive and self.shoot_action[a][0] and self.remaining_missiles[a] > 0 and ang <= sp.max_attack_angle \
                    and distance <= sp.max_attack_distance and interval >= sp.min_attack_interval
                if flag:
                    self._launch(a, agent.enemies[ti], 0, self.remaining_missiles[a])
                    self.remaining_missiles[a] -= 1
                    self.last_shoot_time[a] = self.current_step
            elif sp.launch_kind == ts.L_AUTO_GUN:  # E/tasks/WVR_task.py:67-81, singlecombat_task.py:290-297
                enemy = self._get_target(agent)
                distance, ang = self._attack_geometry(agent, enemy)
                if distance / 1000 < 3 and ang < 5:
                    enemy.bloods -= 5
            elif sp.launch_kind == ts.L_SCENARIO:  # E/tasks/scenario2_task.py:73-114
                sa = self.shoot_action[a]
                f_gun = agent.is_alive and sa[0] and self.remaining_gun[a] > 0
                f_9m = agent.is_alive and sa[1] and self.remaining_9m[a] > 0
                f_120 = agent.is_alive and sa[2] and self.remaining_120b[a] > 0
                f_chaff = agent.is_alive and sa[3] and self.remaining_chaff[a] > 0
                free = lambda: self.last_shot_missile[a] is None or self.last_shot_missile[a].is_done
                if f_gun and free():
                    avail, enemy = self._a2a_available(a)
                    if avail[0]:
                        enemy.bloods -= 5
                        self.remaining_gun[a] -= 1
                if f_120 and free():
                    avail, _ = self._a2a_available(a)
                    if avail[1]:
                        self.last_shot_missile[a] = self._launch(a, self._get_target(agent), 1, self.remaining_120b[a])
                        self.remaining_120b[a] -= 1
                if f_9m and free():
                    avail, _ = self._a2a_available(a)
                    if avail[2]:
                        self.last_shot_missile[a] = self._launch(a, self._get_target(agent), 1, self.remaining_9m[a])
                        self.remaining_9m[a] -= 1
                if f_chaff and (self.last_shot_chaff[a] is None or self.last_shot_chaff[a].done):
                    cnt = 0
                    for m in self.missiles.values():
                        if m.target is agent and m.target_distance < 1000:
                            cnt += 1
                    if cnt > 0:
                        c = Chaff(agent, sp.dt, cnt)
                        self.chaffs.append(c)
                        self.last_shot_chaff[a] = c
                        self.remaining_chaff[a] -= cnt

    # ------------------------------------------------------------------ observations
    def _ego9(self, s: Aircraft):
        return [s.h_sl_m / 5000, math.sin(s.posture[0]), math.cos(s.posture[0]), math.sin(s.posture[1]), math.cos(s.posture[1]),
                s.uvw_mps[0] / 340, s.uvw_mps[1] / 340, s.uvw_mps[2] / 340, s.vc_mps / 340]

    def _rel6(self, ego: Aircraft, other: Aircraft, two_d=False):
        AO, TA, R, side = get_AO_TA_R(list(ego.position) + list(ego.velocity), list(other.position) + list(other.velocity), two_d)
        return [(other.uvw_mps[0] - ego.uvw_mps[0]) / 340, (other.h_sl_m - ego.h_sl_m) / 1000, AO, TA, R / 10000, side]

    def _missile6(self, ego: Aircraft):
        m = ego.check_missile_warning()
        if m is None:
            return None
        mf = list(m.position) + list(m.velocity)
        AO, TA, R, side = get_AO_TA_R(list(ego.position) + list(ego.velocity), mf)
        return [(float(np.linalg.norm(m.velocity)) - ego.uvw_mps[0]) / 340, (mf[2] - ego.h_sl_m) / 1000, AO, TA, R / 10000, side]

    def get_obs(self, a):
        sp = self.spec
        s = self.sims[a]
        o = np.zeros(sp.obs_dim)
        k = sp.obs_kind
        if k == ts.OBS_HEADING:  # E/tasks/heading_task.py:67-100
            psi_deg = s.s["psi_deg"]
            d_alt = _clip((self.target_altitude_ft - s.s["h_sl_ft"]) * 0.3048, -40000.0, 40000.0)
            ang = (self.target_heading_deg - psi_deg) % 360
            if ang > 180:
                ang -= 360
            d_head = _clip(ang, -180.0, 180.0)
            d_vel = _clip(self.target_velocities_u_mps - s.uvw_mps[0], -1400.0, 1400.0)
            o[0] = d_alt / 1000
            o[1] = d_head / 180 * math.pi
            o[2] = d_vel / 340
            o[3:12] = self._ego9(s)
            return np.clip(o, -10, 10)
        o[0:9] = self._ego9(s)
        if k == ts.OBS_1V1:
            o[9:15] = self._rel6(s, s.enemies[0], two_d=True)
            return np.clip(o, -10, 10)
        if k == ts.OBS_1V1_RWR:  # E/tasks/scenario1_task.py:222-314
            dist = sorted(((float(np.linalg.norm(e.position - s.position)), i) for i, e in enumerate(s.enemies)), key=lambda x: x[0])
            target = 0
            for _, i in dist:
                if s.enemies[i].is_alive:
                    target = i
                    break
            o[9:15] = self._rel6(s, s.enemies[target])
            return o
        if k in (ts.OBS_1V1_MISSILE, ts.OBS_NV_MISSILE):
            ti = 0 if k == ts.OBS_1V1_MISSILE else (a if a < sp.n_ego else a - sp.n_ego)
            o[9:15] = self._rel6(s, s.enemies[ti])
            m6 = self._missile6(s)
            if m6 is not None:
                o[15:21] = m6
            return o
        if k in (ts.OBS_MULTI, ts.OBS_MULTI_MISSILE):
            off = 9
            for other in s.partners + s.enemies:
                o[off:off + 6] = self._rel6(s, other)
                off += 6
            o = np.clip(o, -10, 10)
            if k == ts.OBS_MULTI_MISSILE:
                m6 = self._missile6(s)
                if m6 is not None:
                    o[off:off + 6] = m6
            return o
        if k == ts.OBS_NVN:
            off = 9
            for other in s.partners:
                o[off:off + 6] = self._rel6(s, other)
                off += 6
            for other in s.enemies:
                o[off:off + 6] = self._rel6(s, other)
                off += 6
            m6 = self._missile6(s)
            if m6 is not None:
                o[off:off + 6] = m6
            return o
        raise NotImplementedError(k)

    # ------------------------------------------------------------------ rewards
    def _get_reward(self, a):
        sp = self.spec
        if sp.reward_gate == ts.G_DIE_FLAG:
            if self.die_flag[a]:
                return 0.0
            self.die_flag[a] = not self.sims[a].is_alive
        elif sp.reward_gate == ts.G_ALIVE:
            if not self.sims[a].is_alive:
                return 0.0
        tot = 0.0
        for ri in range(len(sp.rewards)):
            tot += self._reward_one(ri, a)
        return tot

    def _process(self, ri, a, new_reward):
        r = self.spec.rewards[ri]
        reward = new_reward * r.scale
        if r.potential:
            reward, self.pre_rewards[ri][a] = reward - self.pre_rewards[ri][a], reward
        return reward

    def _reward_one(self, ri, a):
        sp = self.spec
        r = sp.rewards[ri]
        s = self.sims[a]
        k = r.kind
        ego_f = list(s.position) + list(s.velocity)
        FT = 1 / 3.28084
        if k == ts.R_ALTITUDE:  # altitude_reward.py:20-40
            ego_z = s.position[2] / 1000
            ego_vz = s.velocity[2] / 340
            Pv = 0.
            if ego_z <= r.p0:
                Pv = -_clip(ego_vz / r.p2 * (r.p0 - ego_z) / r.p0, 0., 1.)
            PH = 0.
            if ego_z <= r.p1:
                PH = _clip(ego_z / r.p1, 0., 1.) - 1. - 1.
            return self._process(ri, a, Pv + PH)
        if k == ts.R_POSTURE:  # posture_reward.py:26-75
            new = 0
            for e in s.enemies:
                AO, TA, R, _ = get_AO_TA_R(ego_f, list(e.position) + list(e.velocity))
                new += posture_orientation(int(r.p0), AO, TA) * posture_range(int(r.p1), R / 1000, r.p2)
            return self._process(ri, a, new)
        if k == ts.R_EVENT:  # event_driven_reward.py:15-34
            rew = 0
            if s.is_shotdown:
                rew -= 200
            elif s.is_crash:
                rew -= 200
            for m in s.launch_missiles:
                if m.is_success:
                    rew += 200
            return self._process(ri, a, rew)
        if k == ts.R_MISSILE_POSTURE:  # missile_posture_reward.py:18-46 (bypasses _process; aliasing of previous_missile_v kept)
            rew = 0
            m = s.check_missile_warning()
            if m is not None:
                mv = m.velocity
                av = s.velocity
                if self.prev_missile_v is None:
                    self.prev_missile_v = m  # the reference stores a reference to the missile's live velocity array
                pv = self.prev_missile_v.velocity
                v_dec = (float(np.linalg.norm(pv)) - float(np.linalg.norm(mv))) / 340 * r.scale
                ang = float(np.dot(mv, av)) / (float(np.linalg.norm(mv)) * float(np.linalg.norm(av)))
                if ang < 0:
                    rew = ang / (max(v_dec, 0) + 1)
                else:
                    rew = ang * max(v_dec, 0)
            else:
                self.prev_missile_v = None
                rew = 0
            return rew
        if k == ts.R_SHOOT_PENALTY:  # shoot_penalty_reward.py:17-32
            rew = 0
            if self.remaining_missiles[a] == self.pre_remaining[a] - 1:
                rew -= 30
            self.pre_remaining[a] = self.remaining_missiles[a]
            return self._process(ri, a, rew)
        if k == ts.R_HEADING:  # heading_reward.py:18-71
            roll, p, q = s.s["roll_rad"], s.s["p_rad_sec"], s.s["q_rad_sec"]
            psi_deg = s.s["psi_deg"]
            ang = (self.target_heading_deg - psi_deg) % 360
            if ang > 180:
                ang -= 360
            d_head = _clip(ang, -180.0, 180.0)
            d_alt = _clip((self.target_altitude_ft - s.s["h_sl_ft"]) * 0.3048, -40000.0, 40000.0)
            d_vel = _clip(self.target_velocities_u_mps - s.uvw_mps[0], -1400.0, 1400.0)
            heading_r = math.exp(-((d_head / 5.0) ** 2))
            alt_r = math.exp(-((d_alt / 15.24) ** 2))
            roll_r = math.exp(-((roll / 0.35) ** 2))
            speed_r = math.exp(-((d_vel / 24) ** 2))
            rew = (heading_r * alt_r * roll_r * speed_r) ** (1 / 4)
            if self.current_step > 1:
                rew = rew + (-abs(p - self.hr_last[a][1]) * 1.0) + (-abs(q - self.hr_last[a][2]) * 1.0)
            self.hr_last[a] = (roll, p, q)
            return self._process(ri, a, rew)
        if k == ts.R_RELATIVE_ALTITUDE:  # relative_altitude_reward.py:18-32
            ego_z = s.position[2] / 1000
            enm_z = s.enemies[0].position[2] / 1000
            return self._process(ri, a, min(r.p0 - abs(ego_z - enm_z), 0))
        # per-enemy geometry rewards
        geo = [get_AO_TA_R(ego_f, list(e.position) + list(e.velocity)) for e in s.enemies]
        n = len(geo)
        if k == ts.R_COMBAT_GEOMETRY:  # combat_geometry_reward.py:28-68 (i never increments; prev lists only grow)
            if self.cg_prev is None:
                self.cg_prev = (geo[0][0], geo[0][1])
            new = 0
            for _ in range(n):
                new += -(geo[0][0] - self.cg_prev[0]) - (geo[0][1] - self.cg_prev[1])
            return self._process(ri, a, new)
        if k == ts.R_GUN_BEHIT:  # gun_behit_reward.py:27-54
            new = 0
            for AO, TA, R, _ in geo:
                if (R >= 500 * FT) and (R <= 3000 * FT) and (AO >= 179 * math.pi / 180):
                    new += -5
            return self._process(ri, a, new)
        if k == ts.R_GUN_WEZ:  # gun_WEZ_reward.py:28-55
            new = 0
            for AO, TA, R, _ in geo:
                if (R >= 500 * FT) and (R <= 3000 * FT) and (AO <= 1 * math.pi / 180):
                    new += 5 + 5 * (3000 * FT - R) / (2500 * FT)
            return self._process(ri, a, new)
        if k == ts.R_GUN_TARGETTAIL:  # gun_targettail_reward.py:28-78
            d = []
            for AO, TA, R, _ in geo:
                if (R >= 3000 * FT) and (R <= 5000 * FT):
                    d.append(R * math.sin(TA))
                elif R <= 3000 * FT:
                    d.append(math.sqrt(R ** 2 + (3000 * FT) ** 2 - 2 * R * (3000 * FT) * math.cos(TA)))
                else:
                    d.append(math.sqrt(R ** 2 + (5000 * FT) ** 2 - 2 * R * (5000 * FT) * math.cos(TA)))
            if self.tt_prev is None:  # first call after reset records prev[i] = d[0], d[0], d[1], ...
                self.tt_prev = [d[0]] + [d[i - 1] for i in range(1, n)]
            new = 0
            for i in range(n):
                new += -1 / 60 * math.tanh((d[i] - self.tt_prev[i]) / math.sqrt(geo[i][2]))
            return self._process(ri, a, new)
        if k == ts.R_GUN_WEZDOT:  # gun_WEZDOT_reward.py:29-77
            d = []
            for AO, TA, R, _ in geo:
                if (R >= 500 * FT) and (R <= 3000 * FT):
                    d.append(R * math.sin(AO))
                else:
                    d.append(math.sqrt(R ** 2 + (3000 * FT) ** 2 - 2 * R * (3000 * FT) * math.cos(AO)))
            if self.wd_prev is None:
                self.wd_prev = [d[0]] + [d[i - 1] for i in range(1, n)]
            new = 0
            for i in range(n):
                new += -1 / 60 * math.tanh((d[i] - self.wd_prev[i]) / math.sqrt(geo[i][2]))
            return self._process(ri, a, new)
        raise NotImplementedError(k)

    # ------------------------------------------------------------------ terminations (first done short-circuits)
    def _termination(self, a, info):
        sp = self.spec
        s = self.sims[a]
        for t in sp.terminations:
            done = False
            if t == ts.T_UNREACH_HEADING:  # unreach_heading.py:22-65
                if s.s["sim_time"] >= self.heading_check_time:
                    psi_deg = s.s["psi_deg"]
                    ang = (self.target_heading_deg - psi_deg) % 360
                    if ang > 180:
                        ang -= 360
                    d_head = _clip(ang, -180.0, 180.0)
                    if abs(d_head) > 10:
                        done = True
                    else:
                        inc = ([0.2, 0.4, 0.6, 0.8, 1.0] + [1.0] * 10)[self.heading_turn_counts]
                        draws = [u01(self.seed, self.env_index, RNG_HEADING, self.episode, self.heading_turn_counts, j) for j in range(3)]
                        dh = (-inc + 2 * inc * draws[0]) * sp.heading_increments[0]
                        da = (-inc + 2 * inc * draws[1]) * sp.heading_increments[1]
                        dv = (-inc + 2 * inc * draws[2]) * sp.heading_increments[2]
                        nh = (self.target_heading_deg + dh + 360) % 360
                        self.target_heading_deg = _clip(nh, 0.0, 360.0)
                        self.target_altitude_ft = _clip(self.target_altitude_ft + da, -1400.0, 85000.0)
                        self.target_velocities_u_mps = _clip(self.target_velocities_u_mps + dv, -700.0, 700.0)
                        self.heading_check_time = _clip(self.heading_check_time + sp.check_interval, 0.0, 1000000.0)
                        self.heading_turn_counts += 1
            elif t == ts.T_EXTREME_STATE:  # extreme_state.py + catalog.py:386-416
                ss = s.s
                ev = ss["eci_velocity_mag_fps"] >= 1e10
                er = math.sqrt(ss["p_rad_sec"] ** 2 + ss["q_rad_sec"] ** 2 + ss["r_rad_sec"] ** 2) >= 1000
                ea = ss["h_sl_ft"] >= 1e10
                eacc = max(abs(ss["n_pilot_x"]), abs(ss["n_pilot_y"]), abs(ss["n_pilot_z"])) > 1e1
                done = bool(ea or er or ev or eacc)
                if done:
                    s.status = CRASH
            elif t == ts.T_OVERLOAD:  # overload.py:18-46
                ss = s.s
                if ss["sim_time"] > 10:
                    if abs(ss["n_pilot_x"]) > sp.acc_limit[0] or abs(ss["n_pilot_y"]) > sp.acc_limit[1] or abs(ss["n_pilot_z"] + 1) > sp.acc_limit[2]:
                        done = True
                if done:
                    s.status = CRASH
            elif t == ts.T_LOW_ALTITUDE:  # low_altitude.py:15-34
                done = s.h_sl_m <= sp.altitude_limit
                if done:
                    s.status = CRASH
            elif t == ts.T_TIMEOUT:
                done = self.current_step >= sp.max_steps
            elif t == ts.T_SAFE_RETURN:  # safe_return.py:15-50
                if s.is_shotdown or s.is_crash:
                    done = True
                elif all(not e.is_alive for e in s.enemies) and all(not m.is_alive for m in s.under_missiles):
                    done = True
            if done:
                info["done_cause"][a] = t
                return True
        return False


def posture_orientation(version, AO, TA):  # posture_reward.py:51-63
    if version == 0:
        return (1. - math.tanh(9 * (AO - math.pi / 9))) / 3. + 1 / 3. + min((math.atanh(1. - max(2 * TA / math.pi, 1e-4))) / (2 * math.pi), 0.) + 0.5
    if version == 1:
        return (1. - math.tanh(2 * (AO - math.pi / 2))) / 2. * (math.atanh(1. - max(2 * TA / math.pi, 1e-4))) / (2 * math.pi) + 0.5
    return 1 / (50 * AO / math.pi + 2) + 1 / 2 + min((math.atanh(1. - max(2 * TA / math.pi, 1e-4))) / (2 * math.pi), 0.) + 0.5


def posture_range(version, R, target_dist):  # posture_reward.py:65-75
    if version == 0:
        return math.exp(-(R - target_dist) ** 2 * 0.004) / (1. + math.exp(-(R - target_dist + 2) * 2))
    if version == 1:
        return _clip(1.2 * min(math.exp(-(R - target_dist) * 0.21), 1) / (1. + math.exp(-(R - target_dist + 1) * 0.8)), 0.3, 1)
    if version == 2:
        sg = int(7 - R > 0) - int(7 - R < 0)
        return max(_clip(1.2 * min(math.exp(-(R - target_dist) * 0.21), 1) / (1. + math.exp(-(R - target_dist + 1) * 0.8)), 0.3, 1), sg)
    return 1 * (R < 5) + (R >= 5) * _clip(-0.032 * R ** 2 + 0.284 * R + 0.38, 0, 1) + _clip(math.exp(-0.16 * R), 0, 0.2)
