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

    def __init__(self, spec, n_envs, seed=0, init_states=None):
        import torch
        from aircombat_selfplay_b200.capi import EnvBatch
        from oracle.env_oracle import OracleEnv
        self.spec, self.n = spec, n_envs
        self.gpu = EnvBatch(spec, n_envs, seed=seed)
        self.cpu = [OracleEnv(spec, seed=seed, env_index=i) for i in range(n_envs)]
        if init_states is not None:
            self.gpu.set_init_states(init_states)
            for e in self.cpu:
                e.init_states = [list(r) for r in init_states]
        self.torch = torch
        self.alive_oracle = [True] * n_envs   # envs whose oracle is still being stepped (no auto reset in this harness)

    def reset(self):
        obs, share = self.gpu.reset()
        g_obs = obs.cpu().numpy()
        g_share = None if share is None else share.cpu().numpy()
        c = [e.reset() for e in self.cpu]
        c_obs = np.stack([x[0] for x in c])
        c_share = None if g_share is None else np.stack([x[1] for x in c])
        return (g_obs, g_share), (c_obs, c_share)

    def step(self, act):
        t = self.torch
        obs, share, rew, done, info = self.gpu.step(t.tensor(act, device="cuda"))
        g = dict(obs=obs.cpu().numpy(), share=None if share is None else share.cpu().numpy(), rew=rew.cpu().numpy(),
                 done=done.cpu().numpy().astype(bool), info=info.cpu().numpy())
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


def compare_step(g, c, spec, rew_tol=1e-6, envs=None):
    """Returns a list of human-readable mismatches (empty = parity)."""
    bad = []
    if envs is not None:
        g = dict(g)
        g["obs"] = mask_degenerate_sides(g["obs"], c["obs"], envs)
        if g["share"] is not None:
            A = g["obs"].shape[1]
            g["share"] = np.repeat(g["obs"].reshape(g["obs"].shape[0], 1, -1), A, axis=1)
    tol = obs_tolerance(spec, c["obs"])
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
