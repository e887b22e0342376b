"""Shared helpers: drive the CUDA env batch (through the C ABI) and per-env CPU oracles with the same seeded action
sequences and compare observations, rewards, done flags and aircraft status step by step."""
import numpy as np

CONFIGS = ["singlecontrol/heading", "1v1/NoWeapon/Selfplay", "1v1/ShootMissile/Selfplay", "1v1/DodgeMissile/Selfplay",
           "2v2/NoWeapon/Selfplay", "2v2/ShootMissile/HierarchySelfplay", "scenario1/scenario1", "scenario2/scenario2",
           "scenario2/scenario2_nvn", "scenario3/scenario3", "scenario3/scenario3_nvn", "singlecontrol/approach",
           "scenario1/scenario1_rwr", "scenario2/scenario2_rwr", "scenario3/scenario3_rwr_curriculum", "scenario1/WVR_selfplay",
           "scenario1/Maneuver_curriculum_selfplay", "scenario2/scenario2_curriculum"]


def random_actions(rng, spec, n_envs, shoot_p=0.3, mode="random"):
    """int32 [n_envs, A, 4 + shoot_dim] low-level actions.  mode 'straight' is the reference tests' level-flight action
    (R/tests/test_jsbsim.py:341), 'dive' pushes the nose down to provoke LowAltitude / Overload terminations."""
    A = spec.n_agents
    act = np.zeros((n_envs, A, 4 + spec.shoot_dim), dtype=np.int32)
    if mode == "random":
        act[..., 0:3] = rng.integers(0, 41, (n_envs, A, 3))
        act[..., 3] = rng.integers(0, 30, (n_envs, A))
    elif mode == "straight":
        act[..., 0:4] = np.array([20, 19, 20, 0])
    elif mode == "dive":
        act[..., 0:4] = np.array([20, 40, 20, 29])
    elif mode == "smooth":      # small excursions around level flight: keeps the two teams closing on each other
        act[..., 0:3] = 20 + rng.integers(-3, 4, (n_envs, A, 3))
        act[..., 3] = rng.integers(10, 30, (n_envs, A))
    if spec.shoot_dim:
        act[..., 4:] = (rng.random((n_envs, A, spec.shoot_dim)) < shoot_p).astype(np.int32)
    return act


class Pair:
    """n_envs CUDA envs + n_envs oracle envs of one TaskSpec, same seed."""

    def __init__(self, spec, n_envs, seed=0, init_states=None, sample=None):
        """``sample``: env indices that get an oracle (default: all of them).  The GPU batch always has ``n_envs`` envs
        and is stepped with one distinct action row per env; the comparison then covers the sampled rows only -- this is
        how BASELINE-size batches are checked against the oracle (row-indexed action reads, per-env RNG keys, block and
        warp boundaries) without stepping tens of thousands of Python oracles."""
        import torch
        from aircombat_selfplay_b200.capi import EnvBatch
        from oracle.env_oracle import OracleEnv
        self.spec, self.n = spec, n_envs
        self.sample = list(range(n_envs)) if sample is None else [int(i) for i in sample]
        # device_share_obs: share_obs is what the kernel wrote, not the host layer's stride-0 view of obs
        self.gpu = EnvBatch(spec, n_envs, seed=seed, device_share_obs=True)
        self.cpu = [OracleEnv(spec, seed=seed, env_index=i) for i in self.sample]
        if init_states is not None:
            self.gpu.set_init_states(init_states)
            for e in self.cpu:
                e.init_states = [list(r) for r in init_states]
        self.torch = torch
        self._idx = torch.tensor(self.sample, device="cuda", dtype=torch.long)

    def _rows(self, t):
        return None if t is None else t.index_select(0, self._idx).cpu().numpy()

    def reset(self):
        obs, share = self.gpu.reset()
        g_obs = self._rows(obs)
        g_share = self._rows(share)
        c = [e.reset() for e in self.cpu]
        c_obs = np.stack([x[0] for x in c])
        c_share = None if g_share is None else np.stack([x[1] for x in c])
        return (g_obs, g_share), (c_obs, c_share)

    def step(self, act):
        t = self.torch
        obs, share, rew, done, info = self.gpu.step(act if isinstance(act, t.Tensor) else t.tensor(act, device="cuda"))
        g = dict(obs=self._rows(obs), share=self._rows(share), rew=self._rows(rew), done=self._rows(done).astype(bool),
                 info=self._rows(info))
        if isinstance(act, t.Tensor):
            act = act.index_select(0, self._idx).cpu().numpy()
        else:
            act = act[self.sample]
        c = [e.step(act[i]) for i, e in enumerate(self.cpu)]
        cc = dict(obs=np.stack([x[0] for x in c]), share=None if g["share"] is None else np.stack([x[1] for x in c]),
                  rew=np.stack([x[2] for x in c]), done=np.stack([x[3] for x in c]).astype(bool),
                  cause=np.array([x[4]["done_cause"] for x in c]),
                  status=np.array([[s.status for s in e.sims] for e in self.cpu]),
                  turn=np.array([x[4].get("heading_turn_counts", 0) for x in c]))
        return g, cc


def side_columns(obs_dim):
    """Columns holding the `side` flag of get_AO_TA_R (E/utils/utils.py:75-76): the last entry of every 6-wide block."""
    return [k for k in range(9, obs_dim) if (k - 9) % 6 == 5]


def degenerate_side(envs, tol=1e-9):
    """[n_envs, A] bool: the agent flies exactly along the line of sight to some other entity, so `side` is the sign of
    rounding noise (the reference's yaml geometry -- same longitude, headings 0/180 -- starts every episode this way)."""
    out = np.zeros((len(envs), len(envs[0].sims)), dtype=bool)
    for i, e in enumerate(envs):
        for a, s in enumerate(e.sims):
            others = [o for o in e.sims if o is not s] + list(s.under_missiles)
            for o in others:
                d = np.asarray(o.position)[:2] - s.position[:2]
                v = s.velocity[:2]
                cr = abs(v[0] * d[1] - v[1] * d[0])
                if cr <= tol * (np.linalg.norm(d) * np.linalg.norm(v) + 1e-300):
                    out[i, a] = True
    return out


def mask_degenerate_sides(g_obs, c_obs, envs):
    """Copies the oracle's side flags over the GPU's where the geometry is degenerate (both are noise there)."""
    g_obs = g_obs.copy()
    deg = degenerate_side(envs)
    cols = side_columns(g_obs.shape[-1])
    for i, a in zip(*np.nonzero(deg)):
        g_obs[i, a, cols] = c_obs[i, a, cols]
    return g_obs


def obs_tolerance(spec, c_obs, base=1e-9, missile_tol=1e-5):
    """Per-element tolerance (relative to max(1, |x|)) for observations.

    * own-state and relative blocks: ``base`` (north_star: <= 1e-9 on positions, attitudes, velocities);
    * AO / TA columns: ``acos`` is ill-conditioned near 0 and pi -- one ulp of its argument moves the angle by
      eps / sin(angle) -- so the bound there is base + 4e-15 / sin(angle) (the reference's yaml geometry starts every
      episode at AO ~ 1e-7, TA ~ pi);
    * missile block: stated drift bound ``missile_tol`` -- proportional navigation divides by R_xy^2, a fly-by within
      metres of the target amplifies 1e-12 input differences by ~1e5 (measured 9e-8 after a near miss).
    """
    from aircombat_selfplay_b200 import taskspec as ts
    D = c_obs.shape[-1]
    tol = np.full(c_obs.shape, base)
    if spec.obs_kind == ts.OBS_HEADING:
        return tol
    for k in range(9, D):
        if (k - 9) % 6 in (2, 3):
            s = np.maximum(np.abs(np.sin(c_obs[..., k])), 1e-12)
            tol[..., k] = base + 4e-15 / s
    mblock = {ts.OBS_1V1_MISSILE: 15, ts.OBS_NV_MISSILE: 15, ts.OBS_MULTI_MISSILE: D - 6,
              ts.OBS_NVN: 9 + 6 * (spec.n_agents - 1)}.get(spec.obs_kind)
    if mblock is not None:
        tol[..., mblock:mblock + 6] = np.maximum(tol[..., mblock:mblock + 6], missile_tol)
    return tol


def compare_step(g, c, spec, rew_tol=1e-6, envs=None, missile_tol=1e-5):
    """Returns a list of human-readable mismatches (empty = parity).  ``g["share"]`` -- when the batch materialises
    share_obs on the device (``device_share_obs=True``) it is what the KERNEL wrote, and is compared with the oracle's
    get_state() row by row; otherwise it is the stride-0 view of ``obs`` the host layer exposes."""
    bad = []
    if envs is not None:
        g = dict(g)
        g["obs"] = mask_degenerate_sides(g["obs"], c["obs"], envs)
        if g["share"] is not None:
            # the same masking on every agent's copy of the concatenated observations (noise columns only)
            B, A, D = c["obs"].shape
            deg = degenerate_side(envs)
            cols = side_columns(D)
            sh = g["share"].copy().reshape(B, A, A, D)
            csh = c["share"].reshape(B, A, A, D)
            for i, a in zip(*np.nonzero(deg)):
                sh[i, :, a, cols] = csh[i, :, a, cols]
            g["share"] = sh.reshape(B, A, A * D)
    tol = obs_tolerance(spec, c["obs"], missile_tol=missile_tol)
    e = np.abs(g["obs"] - c["obs"]) / np.maximum(1.0, np.abs(c["obs"]))
    if not np.all(e <= tol):
        k = np.unravel_index(np.nanargmax(np.where(np.isnan(e), np.inf, e / tol)), e.shape)
        bad.append(f"obs{tuple(int(x) for x in k)}: gpu={g['obs'][k]!r} oracle={c['obs'][k]!r} tol={tol[k]:.1e}")
    if g["share"] is not None:
        A = c["obs"].shape[1]
        stol = np.repeat(tol.reshape(tol.shape[0], 1, -1), A, axis=1)
        e = np.abs(g["share"] - c["share"]) / np.maximum(1.0, np.abs(c["share"]))
        if not np.all(e <= stol):
            bad.append(f"share_obs max err {np.nanmax(e):.3e}")
    e = np.abs(g["rew"] - c["rew"])
    if not np.all(e <= rew_tol):
        k = np.unravel_index(np.nanargmax(np.where(np.isnan(e), np.inf, e)), e.shape)
        bad.append(f"reward{tuple(int(x) for x in k)}: gpu={g['rew'][k]!r} oracle={c['rew'][k]!r}")
    if not np.array_equal(g["done"], c["done"]):
        bad.append(f"done: gpu={g['done'].tolist()} oracle={c['done'].tolist()}")
    if not np.array_equal(g["info"][..., 0], c["cause"]):
        bad.append(f"done cause: gpu={g['info'][..., 0].tolist()} oracle={c['cause'].tolist()}")
    if not np.array_equal(g["info"][..., 1], c["status"]):
        bad.append(f"status: gpu={g['info'][..., 1].tolist()} oracle={c['status'].tolist()}")
    return bad


def compare_reset(g_obs, c_obs, spec, envs):
    g_obs = mask_degenerate_sides(g_obs, c_obs, envs)
    e = np.abs(g_obs - c_obs) / np.maximum(1.0, np.abs(c_obs))
    tol = obs_tolerance(spec, c_obs)
    if np.all(e <= tol):
        return []
    k = np.unravel_index(np.argmax(e / tol), e.shape)
    return [f"reset obs{tuple(int(x) for x in k)}: gpu={g_obs[k]!r} oracle={c_obs[k]!r} tol={tol[k]:.1e}"]


def close_init_states(spec, rng, dist_km=(4.0, 12.0)):
    """Head-on geometry at short range so weapons, fuzes and chaff come into play within a few steps."""
    rows = [list(r) for r in spec.init_states]
    d = rng.uniform(*dist_km)
    for a, r in enumerate(rows):
        ego = a < spec.n_ego
        k = a if ego else a - spec.n_ego
        r[0] = 120.0 + 0.01 * k
        r[1] = 60.0 if ego else 60.0 + d / 111.2
        r[2] = 20000.0 + (0 if ego else 300.0)
        r[3] = 0.0 if ego else 180.0
        r[4] = 800.0
    return rows


def low_init_states(spec, h_ft=9200.0):
    """Start just above the LowAltitude limit (2500 m = 8202 ft) so a dive crashes within a few steps."""
    rows = [list(r) for r in spec.init_states]
    for r in rows:
        r[2] = h_ft
    return rows


def inject_oracle_state(pair):
    """Overwrites the continuous state of the CUDA batch with the oracles' (every env of ``pair`` must have an oracle), so
    that the next step starts from IDENTICAL states on both sides and the comparison measures single-step deltas
    (north_star: <= 1e-9 on positions, attitudes, velocities from identical states).  Integer state (status, counters,
    missile bookkeeping) is compared bit-exactly every step and therefore already identical."""
    import math
    import torch
    from tests.fdm_parity import oracle_named_state
    gpu, envs = pair.gpu, pair.cpu
    assert len(envs) == gpu.n_envs
    A, S = gpu.n_agents, None

    def put(arena, fill):
        names, t = gpu.arena(arena)
        h = t.cpu().numpy()
        fill({n: k for k, n in enumerate(names)}, h)
        gpu.set_arena(arena, torch.tensor(h, device="cuda"))

    states = [[oracle_named_state(s.fdm) for s in e.sims] for e in envs]

    def fill_named(ix, h):
        for i, e in enumerate(envs):
            for a in range(A):
                d = states[i][a]
                for n, k in ix.items():
                    if n in d:
                        h[k, i * A + a] = d[n]
    put("fdm", fill_named)
    put("out", fill_named)

    def fill_ac(ix, h):
        for i, e in enumerate(envs):
            for a, s in enumerate(e.sims):
                r = i * A + a
                h[ix["pos_n"], r], h[ix["pos_e"], r], h[ix["pos_u"], r] = s.position
                h[ix["vel_n"], r], h[ix["vel_e"], r], h[ix["vel_d"], r] = s.velocity
                h[ix["h_sl_m"], r] = s.h_sl_m
                h[ix["u_mps"], r], h[ix["v_mps"], r], h[ix["w_mps"], r] = s.uvw_mps
                h[ix["vc_mps"], r] = s.vc_mps
                h[ix["bloods"], r] = s.bloods
                if e.hr_last[a] is not None:
                    h[ix["hr_roll"], r], h[ix["hr_p"], r], h[ix["hr_q"], r] = e.hr_last[a]
                c = e.last_shot_chaff[a]
                if c is not None:
                    h[ix["chaff_n"], r], h[ix["chaff_e"], r], h[ix["chaff_u"], r] = c.position
                    h[ix["chaff_t"], r] = c.t
                for ri in range(len(e.spec.rewards)):
                    h[ix[f"pre_reward{ri}"], r] = e.pre_rewards[ri][a]
    put("ac_d", fill_ac)

    def fill_env(ix, h):
        for i, e in enumerate(envs):
            if e.spec.obs_kind == 0:
                h[ix["tgt_heading_deg"], i] = e.target_heading_deg
                h[ix["tgt_altitude_ft"], i] = e.target_altitude_ft
                h[ix["tgt_velocity_mps"], i] = e.target_velocities_u_mps
                h[ix["check_time"], i] = e.heading_check_time
            if e.cg_prev is not None:
                h[ix["cg_prev_ao"], i], h[ix["cg_prev_ta"], i] = e.cg_prev
            for nm, prev in (("tt_prev", e.tt_prev), ("wd_prev", e.wd_prev)):
                if prev is not None:
                    for j, x in enumerate(prev):
                        h[ix[f"{nm}{j}"], i] = x
    put("env_d", fill_env)

    def fill_ms(ix, h):
        S = h.shape[1] // (len(envs) * A)
        for i, e in enumerate(envs):
            for a, s in enumerate(e.sims):
                for k, m in enumerate(s.launch_missiles):
                    c = (i * A + a) * S + k
                    h[ix["pos_n"], c], h[ix["pos_e"], c], h[ix["pos_u"], c] = m.position
                    h[ix["vel_n"], c], h[ix["vel_e"], c], h[ix["vel_u"], c] = m.velocity
                    h[ix["theta"], c], h[ix["phi"], c] = m.posture[1], m.posture[2]
                    h[ix["alt"], c], h[ix["t"], c], h[ix["m"], c] = m.alt, m.t, m.m
                    h[ix["dtheta"], c], h[ix["dphi"], c], h[ix["d_prev"], c] = m.dtheta, m.dphi, m.distance_pre
                    h[ix["sin_theta"], c], h[ix["cos_theta"], c] = math.sin(m.posture[1]), math.cos(m.posture[1])
    put("ms_d", fill_ms)
