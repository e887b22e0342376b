"""GPU parity of the env-step layer: the CUDA kernels (through the C ABI, acs_env_*) against the CPU oracle on the same
seeded action sequences, for every task family.

Tolerances (BASELINE.json north_star): observations <= 1e-9 relative (they are functions of positions, attitudes and
velocities), rewards <= 1e-6 absolute, done flags / done causes / aircraft status bit-exact."""
import numpy as np
import pytest
import torch

from aircombat_selfplay_b200.tasks import load_spec
from tests.env_parity import (CONFIGS, Pair, close_init_states, compare_reset, compare_step, inject_oracle_state, low_init_states,
                              random_actions)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 2, 3], ids=["one-thread-frame", "two-warp-frame", "three-warp-frame", "four-warp-frame"], autouse=True)
def frame_split(request, monkeypatch):
    """Every parity case runs against both substep kernels (include/acs.h, acs_env_set_option "frame_split"); the
    variable is read at acs_env_create."""
    monkeypatch.setenv("ACS_FRAME_SPLIT", str(request.param))
    return request.param


def _run(name, n_envs, steps, mode, init=None, substeps=None, seed=3):
    spec = load_spec(name, substeps_override=substeps)
    rng = np.random.default_rng(seed)
    init_states = {"close": close_init_states(spec, rng), "low": low_init_states(spec), None: None}[init]
    p = Pair(spec, n_envs, seed=seed, init_states=init_states)
    (g_obs, _), (c_obs, _) = p.reset()
    assert not compare_reset(g_obs, c_obs, spec, p.cpu)
    events = set()
    for t in range(steps):
        g, c = p.step(random_actions(rng, spec, n_envs, mode=mode))
        bad = compare_step(g, c, spec, envs=p.cpu)
        assert not bad, (name, mode, t, bad)
        for e in p.cpu:
            events.update({0: "launched", 1: "hit", 2: "miss"}[m.status] for m in e.missiles.values())
            if e.chaffs:
                events.add("chaff")
        events.update(f"term{z}" for z in np.unique(c["cause"]) if z >= 0)
    names, faults = p.gpu.arena("env_i")
    assert int(faults[names.index("faults")].sum()) == 0
    return events


@pytest.mark.parametrize("name", CONFIGS)
def test_random_actions(name):
    _run(name, n_envs=4, steps=25, mode="random")


@pytest.mark.parametrize("name", [c for c in CONFIGS if not c.startswith("singlecontrol/")])
def test_close_engagement(name):
    """Head-on at 4-12 km: launches, fuze hits, misses, chaff, shot-down aircraft, SafeReturn terminations."""
    ev = _run(name, n_envs=6, steps=60, mode="smooth", init="close")
    spec = load_spec(name)
    if spec.launch_kind in (1, 2, 3, 4):      # the missile-carrying launch rules
        assert "launched" in ev and "miss" in ev


@pytest.mark.parametrize("name", ["1v1/NoWeapon/Selfplay", "2v2/NoWeapon/Selfplay", "1v1/ShootMissile/Selfplay",
                                  "scenario2/scenario2"])
def test_crash_terminations(name):
    ev = _run(name, n_envs=3, steps=60, mode="dive", init="low")
    assert "term3" in ev      # LowAltitude


def test_substeps_6_and_12():
    _run("1v1/NoWeapon/Selfplay", n_envs=4, steps=20, mode="random", substeps=6)
    _run("scenario2/scenario2", n_envs=4, steps=20, mode="random", substeps=12)


def test_heading_task_retargets():
    """UnreachHeading re-targets with keyed draws at sim_time >= check_time (first at step 1: check_time starts at 0)."""
    spec = load_spec("singlecontrol/heading")
    p = Pair(spec, 8, seed=11)
    p.reset()
    rng = np.random.default_rng(0)
    turns = 0
    for t in range(12):
        g, c = p.step(random_actions(rng, spec, 8, mode="straight"))
        assert not compare_step(g, c, spec, envs=p.cpu)
        assert np.array_equal(g["info"][:, 0, 3], c["turn"])
        turns = max(turns, int(c["turn"].max()))
    assert turns >= 1


def test_auto_reset_matches_a_fresh_reset():
    """acs_env_step(auto_reset=1): an env whose agents are all done comes back with its reset observation while the
    rewards / dones of the terminal step are kept (reference envs/env_wrappers.py:191-204)."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec("1v1/NoWeapon/Selfplay")
    spec.max_steps = 3
    n = 5
    b = EnvBatch(spec, n, seed=1)
    obs0 = b.reset()[0].clone()
    rng = np.random.default_rng(0)
    for t in range(3):
        obs, _, rew, done, info = b.step(torch.tensor(random_actions(rng, spec, n), device="cuda"), auto_reset=True)
    assert bool(done.all()) and bool(b.env_done.all())
    assert torch.allclose(obs, obs0, rtol=0, atol=0)          # the 1v1 reset is deterministic: bit-identical
    names, ei = b.arena("env_i")
    assert int(ei[names.index("current_step")].abs().max()) == 0
    assert int(ei[names.index("episode")].min()) == 1
    # the next step runs from the reset state
    obs, _, rew, done, info = b.step(torch.tensor(random_actions(rng, spec, n), device="cuda"), auto_reset=True)
    assert not bool(done.any())
    assert int(info[:, 0, 2].min()) == 1


def test_determinism_bitwise():
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec("scenario2/scenario2")
    res = []
    for rep in range(2):
        rng = np.random.default_rng(5)
        b = EnvBatch(spec, 64, seed=9)
        b.set_init_states(close_init_states(spec, np.random.default_rng(1)))
        b.reset()
        for t in range(30):
            b.step(torch.tensor(random_actions(rng, spec, 64, mode="smooth"), device="cuda"), auto_reset=True)
        res.append(b.out_buf.clone())
    assert torch.equal(res[0], res[1])


def test_env_offset_keeps_rng_streams_per_env():
    """Sharding: a handle that owns envs [8, 16) of a 16-env job draws the same numbers as the last 8 envs of one
    16-env handle (multi-GPU slices are bit-identical to the single-GPU run)."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec("singlecontrol/heading")
    rng = np.random.default_rng(2)
    acts = [random_actions(rng, spec, 16) for _ in range(8)]
    full = EnvBatch(spec, 16, seed=4)
    half = EnvBatch(spec, 8, seed=4, env_offset=8)
    o_full, o_half = full.reset()[0].clone(), half.reset()[0].clone()
    assert torch.equal(o_full[8:], o_half)
    for a in acts:
        of = full.step(torch.tensor(a, device="cuda"), auto_reset=True)[0]
        oh = half.step(torch.tensor(a[8:], device="cuda"), auto_reset=True)[0]
        assert torch.equal(of[8:], oh)


@pytest.mark.parametrize("name", ["scenario2/scenario2", "2v2/NoWeapon/Selfplay"])
def test_device_share_obs_equals_the_broadcast_view(name):
    """share_obs written by the kernel (device_share_obs=True) == the stride-0 view of obs used by default (the no-weapon
    config takes the observation-warp path of k_env_post and, with short episodes, the fused template reset)."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec(name)
    spec.max_steps = 3
    rng = np.random.default_rng(8)
    a = EnvBatch(spec, 16, seed=2, device_share_obs=True)
    b = EnvBatch(spec, 16, seed=2)
    oa, sa = a.reset()
    ob, sb = b.reset()
    assert sa.shape == sb.shape == (16, 4, 4 * spec.obs_dim) and torch.equal(sa, sb) and torch.equal(oa, ob)
    for _ in range(5):
        act = torch.tensor(random_actions(rng, spec, 16), device="cuda")
        _, sa, *_ = a.step(act, auto_reset=True)
        _, sb, *_ = b.step(act, auto_reset=True)
        assert torch.equal(sa, sb)


def test_multi_warp_frames_match_one_thread_frame():
    """The three substep kernels evaluate the same expressions on different threads: states agree to rounding of the
    compiler's FMA contraction choices, flags exactly (ragged batch: 1000 envs x 4 lanes is not a multiple of a block)."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec("scenario2/scenario2")
    n = 1000
    bs = []
    for split in (0, 1, 2, 3):
        b = EnvBatch(spec, n, seed=9)
        b.set_option("frame_split", split)
        b.set_init_states(close_init_states(spec, np.random.default_rng(1)))
        b.reset()
        bs.append(b)
    rng = np.random.default_rng(5)
    for t in range(20):
        act = torch.tensor(random_actions(rng, spec, n, mode="smooth"), device="cuda")
        ra = [x.clone() for x in bs[0].step(act, auto_reset=True)]
        for b in bs[1:]:
            rb = b.step(act, auto_reset=True)
            assert torch.equal(ra[3], rb[3]) and torch.equal(ra[4], rb[4]), t           # dones, info
            assert torch.allclose(ra[2], rb[2], rtol=0, atol=1e-6), t                   # rewards
    sa = bs[0].arena("fdm")[1]
    scale = sa.abs().amax(dim=1, keepdim=True).clamp_min(1.0)
    for b in bs[1:]:
        assert float(((sa - b.arena("fdm")[1]).abs() / scale).max()) < 1e-8


@pytest.mark.parametrize("name", ["scenario2/scenario2", "1v1/ShootMissile/Selfplay", "2v2/NoWeapon/Selfplay", "scenario3/scenario3"])
def test_reset_template_is_bit_identical_to_a_recomputed_reset(name, monkeypatch):
    """Fixed per-lane initial conditions: reset() always produces the same env state, so it is computed once per handle and
    the (auto-)reset copies it -- mode 1 the FDM reload only, mode 2 every word reset() writes plus the reset observation,
    fused into k_env_post or in its own launch.  All give the bits of the recomputed reset (mode 0), in every arena, also
    after set_init_states."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec(name)
    spec.max_steps = 4
    n = 70
    bs = []
    for tpl, fused in (("0", "0"), ("1", "0"), ("2", "0"), ("2", "1")):
        monkeypatch.setenv("ACS_RESET_TEMPLATE", tpl)
        monkeypatch.setenv("ACS_FUSED_RESET", fused)
        b = EnvBatch(spec, n, seed=3)
        assert b.get_option("reset_template") == int(tpl)
        b.reset()
        bs.append(b)
    rng = np.random.default_rng(1)
    for t in range(11):
        if t == 5:
            init = close_init_states(spec, np.random.default_rng(7))
            for b in bs:
                b.set_init_states(init)
        if t == 8:          # an explicit masked reset goes through the same template
            mask = torch.tensor(np.arange(n) % 3 == 0, device="cuda").to(torch.uint8)
            for b in bs:
                b.reset(mask)
        act = torch.tensor(random_actions(rng, spec, n), device="cuda")
        for b in bs:
            b.step(act, auto_reset=True)
        for b in bs[1:]:
            assert torch.equal(bs[0].out_buf, b.out_buf), t
    for name_ in ("fdm", "out", "ac_d", "ac_i", "env_d", "env_i", "ms_d", "ms_i"):
        for b in bs[1:]:
            assert torch.equal(bs[0].arena(name_)[1], b.arena(name_)[1]), name_
    names, ei = bs[0].arena("env_i")
    assert int(ei[names.index("episode")].min()) >= 2


@pytest.mark.parametrize("name", ["1v1/NoWeapon/Selfplay", "2v2/NoWeapon/Selfplay"])
def test_observation_warps_give_the_same_bits(name, monkeypatch):
    """k_env_post with get_obs on its own warps (no-weapon tasks) == the single-pass kernel, bit for bit, resets included."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec(name)
    spec.max_steps = 5
    n = 333
    bs = []
    for on in ("0", "1"):
        monkeypatch.setenv("ACS_POST_SPLIT", on)
        b = EnvBatch(spec, n, seed=2)
        b.set_init_states(low_init_states(spec))
        b.reset()
        bs.append(b)
    rng = np.random.default_rng(4)
    for t in range(12):
        act = torch.tensor(random_actions(rng, spec, n, mode="dive" if t % 2 else "random"), device="cuda")
        for b in bs:
            b.step(act, auto_reset=True)
        assert torch.equal(bs[0].out_buf, bs[1].out_buf), t
    for arena in ("fdm", "out", "ac_d", "ac_i", "env_d", "env_i"):
        assert torch.equal(bs[0].arena(arena)[1], bs[1].arena(arena)[1]), arena


def test_single_env_single_substep():
    """Smallest launch: one env, one substep per step (a block of mostly idle lanes; K = 1 means every frame is both the
    first and the last of its step, the four-warp frame computes no look-ahead inside the loop)."""
    _run("1v1/NoWeapon/Selfplay", n_envs=1, steps=12, mode="random", substeps=1)
    _run("scenario2/scenario2", n_envs=1, steps=12, mode="smooth", init="close", substeps=1)


@pytest.mark.parametrize("name,n", [("1v1/NoWeapon/Selfplay", 65536), ("2v2/ShootMissile/HierarchySelfplay", 16384)])
def test_full_size_batches_are_replicas(name, n, frame_split):
    """BASELINE.json sizes, through a size-independent property: the combat tasks start every env from the same state, so
    with the same action rows every env of the batch must stay bit-identical to env 0 (grid / lane indexing, reset
    template scatter and auto-reset at full size)."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec(name, substeps_override=12)
    spec.max_steps = 4
    b = EnvBatch(spec, n, seed=1)
    obs0 = b.reset()[0].clone()
    assert torch.equal(obs0, obs0[:1].expand_as(obs0))
    rng = np.random.default_rng(3)
    for t in range(6):
        row = torch.tensor(random_actions(rng, spec, 1), device="cuda")
        obs, _, rew, done, info = b.step(row.expand(n, -1, -1).contiguous(), auto_reset=True)
        assert torch.equal(obs, obs[:1].expand_as(obs)), t
        assert torch.equal(rew, rew[:1].expand_as(rew)) and torch.equal(done, done[:1].expand_as(done)), t
    names, ei = b.arena("env_i")
    assert int(ei[names.index("episode")].min()) == int(ei[names.index("episode")].max()) == 1
    b.close()


# ---------------------------------------------------------------------------------------------- BASELINE sizes vs the oracle
FULL_SIZE = [("1v1/NoWeapon/Selfplay", 4096), ("1v1/NoWeapon/Selfplay", 65536), ("1v1/ShootMissile/Selfplay", 16384),
             ("2v2/ShootMissile/HierarchySelfplay", 8192), ("scenario3/scenario3", 4096)]


@pytest.mark.parametrize("name,n", FULL_SIZE)
def test_full_size_batches_match_sampled_oracles(name, n, frame_split):
    """Every BASELINE.json batch size with DISTINCT random action rows per env (close-engagement geometry so the weapon
    configs launch): 20 sampled envs -- first, last, warp / block boundaries, random -- are compared step by step with
    their own oracle (env index keys the RNG streams, the action read is row-indexed), on every substep kernel."""
    spec = load_spec(name, substeps_override=12)
    rng = np.random.default_rng(17)
    init = close_init_states(spec, rng)
    edge = [0, 1, 15, 16, 31, 32, 63, 64, 127, 128, n // 2 - 1, n // 2, n - 129, n - 128, n - 2, n - 1]
    sample = sorted(set(edge) | set(int(x) for x in rng.integers(0, n, 6)))
    p = Pair(spec, n, seed=5, init_states=init, sample=sample)
    (g_obs, _), (c_obs, _) = p.reset()
    assert not compare_reset(g_obs, c_obs, spec, p.cpu)
    launched = False
    for t in range(24):
        g, c = p.step(random_actions(rng, spec, n, mode="smooth", shoot_p=0.3))
        bad = compare_step(g, c, spec, envs=p.cpu)
        assert not bad, (name, n, t, bad)
        launched = launched or any(e.missiles for e in p.cpu)
    if spec.launch_kind in (1, 2, 3, 4):
        assert launched
    names, faults = p.gpu.arena("env_i")
    assert int(faults[names.index("faults")].sum()) == 0
    p.gpu.close()


@pytest.mark.parametrize("name", ["1v1/NoWeapon/Selfplay", "1v1/ShootMissile/Selfplay", "1v1/DodgeMissile/Selfplay",
                                  "2v2/ShootMissile/HierarchySelfplay", "scenario2/scenario2", "scenario3/scenario3_nvn",
                                  "singlecontrol/heading"])
def test_single_step_deltas_from_identical_states(name):
    """north_star: single-step deltas from IDENTICAL states.  After every step the oracles' continuous state (FDM, derived
    aircraft properties, missiles, chaff, reward memories) is copied into the CUDA arenas (acs_env_set_arena), so each
    comparison measures one interaction step (K substeps) from a common state: the missile block then holds the same
    1e-9 as everything else -- the 1e-5 drift bound of the free-running tests is proportional-navigation chaos
    accumulated over a fly-by, not a per-step error."""
    spec = load_spec(name)
    rng = np.random.default_rng(23)
    init = None if name.startswith("singlecontrol/") else close_init_states(spec, rng)
    n = 6
    p = Pair(spec, n, seed=4, init_states=init)
    (g_obs, _), (c_obs, _) = p.reset()
    assert not compare_reset(g_obs, c_obs, spec, p.cpu)
    inject_oracle_state(p)
    ev = set()
    for t in range(70):
        g, c = p.step(random_actions(rng, spec, n, mode="smooth", shoot_p=0.3))
        bad = compare_step(g, c, spec, envs=p.cpu, missile_tol=1e-9)
        assert not bad, (name, t, bad)
        inject_oracle_state(p)
        for e in p.cpu:
            ev.update({0: "launched", 1: "hit", 2: "miss"}[m.status] for m in e.missiles.values())
    if spec.launch_kind in (1, 2, 3, 4):
        assert "launched" in ev and "miss" in ev
    p.gpu.close()


@pytest.mark.parametrize("name", ["1v1/ShootMissile/Selfplay", "2v2/ShootMissile/HierarchySelfplay", "scenario2/scenario2",
                                  "1v1/DodgeMissile/Selfplay"])
def test_missiles_after_the_loop_equal_the_in_kernel_missile_phase(name, monkeypatch):
    """One-thread frame: the aircraft are integrated without stopping and k_env_missiles does the missile / chaff work of the
    step afterwards -- missiles that cannot score this step on their own (K substeps in registers against the recorded target
    trajectory), envs with a missile in reach in per-substep lockstep, a hit restoring the target's state of that substep.
    The two-warp frame keeps the missile phase between the frames.  Same flags bit for bit and the same numbers to
    rounding (same expressions, other threads), through launches, fly-bys, hits, double hits, chaff and auto-resets."""
    from aircombat_selfplay_b200.capi import EnvBatch
    spec = load_spec(name)
    spec.max_steps = 60
    n = 300
    bs = []
    for split in (0, 1):
        monkeypatch.setenv("ACS_FRAME_SPLIT", str(split))
        b = EnvBatch(spec, n, seed=3)
        assert b.get_option("launches_per_step") == (3 if split == 0 else 2)
        b.set_init_states(close_init_states(spec, np.random.default_rng(2)))
        b.reset()
        bs.append(b)
    rng = np.random.default_rng(9)
    fast = lockstep = hits = 0
    for t in range(120):
        act = torch.tensor(random_actions(rng, spec, n, mode="smooth", shoot_p=0.3), device="cuda")
        ra = [None if x is None else x.clone() for x in bs[0].step(act, auto_reset=True)]
        rb = bs[1].step(act, auto_reset=True)
        assert torch.equal(ra[3], rb[3]) and torch.equal(ra[4], rb[4]), t                         # dones, info (cause, status, step)
        # two FDM kernels: 1e-13 per frame grows over a 60-step episode; the potential-based rewards scale it by 15 / step
        assert torch.allclose(ra[0], rb[0], rtol=0, atol=1e-6) and torch.allclose(ra[2], rb[2], rtol=0, atol=1e-4), t
        names, ei = bs[0].arena("env_i")
        d = ei[names.index("deferred")]
        fast += int((d == 1).sum()); lockstep += int((d == 2).sum())
        mn, mi = bs[0].arena("ms_i")
        mj = bs[1].arena("ms_i")[1]
        for f, fname in enumerate(mn):      # missile status / bookkeeping; `consec` counts dist > d_prev comparisons, which the
            if fname != "consec":           # 1e-13 differences between the two FDM kernels may flip at closest approach
                assert torch.equal(mi[f], mj[f]), (t, fname)
        hits += int((mi[mn.index("status")] == 1).sum())
        assert int(ei[names.index("faults")].sum()) == 0
    assert fast > 0 and lockstep > 0
    if spec.launch_kind != 4:            # the AIM-9L tasks score hits in this geometry (300 m fuze)
        assert hits > 0


@pytest.mark.parametrize("name", ["singlecontrol/heading", "1v1/NoWeapon/Selfplay", "1v1/ShootMissile/Selfplay", "1v1/DodgeMissile/Selfplay",
                                  "2v2/NoWeapon/Selfplay", "2v2/ShootMissile/HierarchySelfplay", "scenario1/scenario1", "scenario2/scenario2",
                                  "scenario3/scenario3_nvn", "scenario1/WVR_selfplay"])
def test_specialised_post_kernel_equals_the_generic_one(name, monkeypatch):
    """k_env_post is compiled per task family (observation packer, launch rule, reward classes as template constants); the
    instantiation acs_env_create picks returns the bits of the generic kernel that reads all three from the config."""
    from aircombat_selfplay_b200.capi import EnvBatch
    monkeypatch.setenv("ACS_FRAME_SPLIT", "0")
    spec = load_spec(name)
    spec.max_steps = 40
    n = 200
    bs = []
    for generic in ("0", "1"):
        monkeypatch.setenv("ACS_POST_GENERIC", generic)
        b = EnvBatch(spec, n, seed=6, device_share_obs=True)
        assert (b.get_option("post_family") >= 0) == (generic == "0")
        if not name.startswith("singlecontrol/"):
            b.set_init_states(close_init_states(spec, np.random.default_rng(3)))
        b.reset()
        bs.append(b)
    rng = np.random.default_rng(10)
    for t in range(90):
        act = torch.tensor(random_actions(rng, spec, n, mode="smooth" if t % 3 else "random", shoot_p=0.3), device="cuda")
        for b in bs:
            b.step(act, auto_reset=True)
        assert torch.equal(bs[0].out_buf, bs[1].out_buf), t
    for arena in ("ac_d", "ac_i", "env_d", "env_i", "ms_d", "ms_i"):
        assert torch.equal(bs[0].arena(arena)[1], bs[1].arena(arena)[1]), arena
