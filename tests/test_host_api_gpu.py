"""GPU tests of the reference-facing Python API, modelled on the reference's own tests (reference tests/test_jsbsim.py):
env classes (shapes, same seed => same trajectory, crash semantics) and the VecEnv contract (shapes, infos, auto reset)."""
import numpy as np
from pathlib import Path
import pytest
import torch

from aircombat_selfplay_b200.env_wrappers import (BatchedVecEnv, DummyVecEnv, ShareBatchedVecEnv, ShareDummyVecEnv,
                                                  ShareSubprocVecEnv, SubprocVecEnv)
from aircombat_selfplay_b200.envs import BatchedEnv, MultipleCombatEnv, SingleCombatEnv, SingleControlEnv

pytestmark = pytest.mark.gpu


def _sample(env, n=None):
    return [env.action_space.sample() for _ in range(n or env.num_agents)]


class TestSingleControlEnv:
    def test_env(self):                                   # reference tests/test_jsbsim.py:18-64
        env = SingleControlEnv("singlecontrol/heading")
        assert env.num_agents == 1
        for agent in env.agents.values():
            assert len(agent.partners) == 0 and len(agent.enemies) == 0
        obs_shape = (env.num_agents, *env.observation_space.shape)
        env.seed(0)
        env.action_space.seed(0)
        obs = env.reset()
        assert obs.shape == obs_shape
        obs_buf, act_buf, rew_buf, done_buf = [obs], [], [], []
        for _ in range(40):
            actions = np.array(_sample(env))
            obs, reward, done, info = env.step(actions)
            assert obs.shape == obs_shape and reward.shape == (1, 1) and done.shape == (1, 1)
            assert info["current_step"] == env.current_step
            assert ("heading_turn_counts" in info) == (bool(done) and info["done_condition"][0] == "unreach_heading")
            act_buf.append(actions); obs_buf.append(obs); rew_buf.append(reward); done_buf.append(done)
            if done:
                assert env.current_step <= env.max_steps
                break
        env.seed(0)                                        # repetition: same seed => same data
        obs = env.reset()
        assert np.linalg.norm(obs - obs_buf[0]) < 1e-8
        for t in range(len(done_buf)):
            obs, reward, done, info = env.step(act_buf[t])
            assert np.linalg.norm(obs - obs_buf[t + 1]) < 1e-8 and np.all(reward == rew_buf[t]) and np.all(done == done_buf[t])

    @pytest.mark.parametrize("vecenv", [DummyVecEnv, SubprocVecEnv])
    def test_vec_env(self, vecenv):                       # reference tests/test_jsbsim.py:66-89
        n = 4
        envs = vecenv([lambda: SingleControlEnv("singlecontrol/heading") for _ in range(n)])
        obs_shape = (n, envs.num_agents, *envs.observation_space.shape)
        obss = envs.reset()
        assert obss.shape == obs_shape
        actions = np.array([[envs.action_space.sample() for _ in range(envs.num_agents)] for _ in range(n)])
        for _ in range(1500):     # the first heading check that can fail is at sim_time 30 s = step 300 (K = 6)
            obss, rewards, dones, infos = envs.step(actions)
            assert obss.shape == obs_shape and rewards.shape == (n, 1, 1) and dones.shape == (n, 1, 1) \
                and infos.shape[0] == n and isinstance(infos[0], dict) and "current_step" in infos[0]
            if np.any(dones):
                break
        assert np.any(dones)
        envs.close()


class TestSingleCombatEnv:
    @pytest.mark.parametrize("config", ["1v1/NoWeapon/Selfplay", "1v1/DodgeMissile/Selfplay", "1v1/ShootMissile/Selfplay",
                                        "1v1/NoWeapon/HierarchySelfplay", "scenario1/scenario1"])
    def test_env(self, config):                           # reference tests/test_jsbsim.py:94-145
        env = SingleCombatEnv(config)
        for agent in env.agents.values():
            assert len(agent.partners) == 0 and len(agent.enemies) == 1
        obs_shape = (env.num_agents, *env.observation_space.shape)
        env.seed(0)
        env.action_space.seed(0)
        obs = env.reset()
        assert obs.shape == obs_shape
        obs_buf, act_buf, rew_buf, done_buf = [obs], [], [], []
        for _ in range(30):
            actions = _sample(env) if env.current_step % 2 == 0 else np.array(
                [np.concatenate([np.atleast_1d(x).ravel() for x in a]) if isinstance(a, tuple) else a for a in _sample(env)])
            obs, rewards, dones, info = env.step(actions)
            assert obs.shape == obs_shape and rewards.shape == (2, 1) and dones.shape == (2, 1)
            act_buf.append(actions); obs_buf.append(obs); rew_buf.append(rewards); done_buf.append(dones)
            if np.all(dones):
                break
        env.seed(0)
        obs = env.reset()
        assert np.linalg.norm(obs - obs_buf[0]) < 1e-8
        for t in range(len(done_buf)):
            obs, rewards, dones, info = env.step(act_buf[t])
            assert np.linalg.norm(obs - obs_buf[t + 1]) < 1e-8 and np.all(rewards == rew_buf[t]) and np.all(dones == done_buf[t])

    def test_agent_crash(self):                           # reference tests/test_jsbsim.py:147-157
        env = SingleCombatEnv("1v1/NoWeapon/Selfplay")
        env.seed(0)
        env.reset()
        env.agents[env.ego_ids[0]].crash()
        obs, rewards, dones, info = env.step(np.array(_sample(env)))
        assert np.min(rewards) < -100      # crash reward
        assert np.all(dones)               # no weapons: once one side is gone, the env terminates
        assert "done_condition" in info

    def test_wrong_env_class(self):
        with pytest.raises(NotImplementedError):
            SingleCombatEnv("scenario2/scenario2")


class TestMultipleCombatEnv:
    @pytest.mark.parametrize("config", ["2v2/NoWeapon/Selfplay", "2v2/ShootMissile/HierarchySelfplay", "scenario2/scenario2",
                                        "scenario3/scenario3_nvn"])
    def test_env(self, config):                           # reference tests/test_jsbsim.py:279-330
        env = MultipleCombatEnv(config)
        A = env.num_agents
        assert A == len(env.agents)
        for agent in env.agents.values():
            assert len(agent.partners) == A // 2 - 1 and len(agent.enemies) == A // 2
        obs_shape = (A, *env.observation_space.shape)
        share_shape = (A, *env.share_observation_space.shape)
        env.seed(0)
        env.action_space.seed(0)
        obs, share = env.reset()
        assert obs.shape == obs_shape and share.shape == share_shape
        assert np.array_equal(share[0], obs.reshape(-1))
        bufs = []
        for _ in range(20):
            actions = _sample(env)
            obs, share, rewards, dones, info = env.step(actions)
            assert obs.shape == obs_shape and share.shape == share_shape and rewards.shape == (A, 1) and dones.shape == (A, 1)
            # team reward: every member of a team gets the team mean (reference multiplecombat_env.py:170-175)
            assert np.all(rewards[:A // 2] == rewards[0]) and np.all(rewards[A // 2:] == rewards[A // 2])
            bufs.append((actions, obs, rewards, dones))
        env.seed(0)
        env.reset()
        for actions, o, r, d in bufs:
            obs, share, rewards, dones, info = env.step(actions)
            assert np.linalg.norm(obs - o) < 1e-8 and np.all(rewards == r) and np.all(dones == d)

    def test_dead_agent_semantics(self):                  # reference tests/test_jsbsim.py:332-382
        env = MultipleCombatEnv("2v2/NoWeapon/Selfplay")
        env.seed(0)
        env.reset()
        env.agents[env.ego_ids[0]].crash()
        straight = np.array([[20, 19, 20, 0]] * 4)
        obs, share, rewards, dones, info = env.step(straight)
        assert dones[0][0] and not dones[1][0] and not np.all(dones)
        frozen = obs[0, :9].copy()
        for _ in range(3):
            obs, share, rewards, dones, info = env.step(straight)
            assert dones[0][0] and np.linalg.norm(obs[0, :9] - frozen) < 1e-8     # a dead aircraft's own state is frozen

    @pytest.mark.parametrize("vecenv", [ShareDummyVecEnv, ShareSubprocVecEnv])
    def test_vec_env(self, vecenv):                       # reference tests/test_jsbsim.py:384-404
        n = 4
        envs = vecenv([lambda: MultipleCombatEnv("scenario2/scenario2") for _ in range(n)])
        A = envs.num_agents
        obs_shape = (n, A, *envs.observation_space.shape)
        share_shape = (n, A, *envs.share_observation_space.shape)
        obss, share = envs.reset()
        assert obss.shape == obs_shape and share.shape == share_shape
        actions = [[envs.action_space.sample() for _ in range(A)] for _ in range(n)]
        for _ in range(10):
            obss, share, rewards, dones, infos = envs.step(actions)
            assert obss.shape == obs_shape and share.shape == share_shape and rewards.shape == (n, A, 1) \
                and dones.shape == (n, A, 1) and infos.shape[0] == n and isinstance(infos[0], dict)
            assert infos[0]["current_step"] >= 1
        envs.close()


def test_vecenv_matches_device_api_and_auto_resets():
    """The numpy-facing VecEnv returns exactly what the device API computes, and an env whose agents are all done
    restarts inside the same call (reset obs in place, terminal rewards/dones kept)."""
    n = 16
    ve = BatchedVecEnv("1v1/NoWeapon/Selfplay", n, seed=3)
    ve.core.spec.max_steps  # noqa: B018
    ref = BatchedEnv("1v1/NoWeapon/Selfplay", n, seed=3)
    o_ve = ve.reset()
    o_ref = ref.reset()[0]
    assert np.array_equal(o_ve, o_ref.cpu().numpy())
    rng = np.random.default_rng(0)
    for t in range(5):
        a = rng.integers(0, 30, (n, 2, 4))
        obs, rew, done, infos = ve.step(a)
        o2, _, r2, d2, i2 = ref.step(torch.tensor(a, dtype=torch.int32, device="cuda"))
        assert np.array_equal(obs, o2.cpu().numpy()) and np.array_equal(rew[..., 0], r2.cpu().numpy())
        assert np.array_equal(done[..., 0], d2.cpu().numpy().astype(bool))
        assert infos[3]["current_step"] == t + 1
    ve.close(); ref.close()


def test_step_outputs_stay_valid_while_held_without_host_copies():
    """copy=True (default): the arrays a step returns are views of a pinned output slot that is never rewritten while the
    caller (or a view it derived) holds them -- the reference's "fresh arrays" contract without copying them out; a slot
    whose arrays were dropped is reused; with every slot held the arrays are copied out instead; copy=False keeps one
    buffer that the next step overwrites."""
    n = 8
    ve = ShareBatchedVecEnv("2v2/NoWeapon/Selfplay", n, seed=2)
    ref = BatchedEnv("2v2/NoWeapon/Selfplay", n, seed=2)
    rng = np.random.default_rng(0)
    acts = [rng.integers(0, 30, (n, ve.num_agents, 4)) for _ in range(14)]
    truth = []                                              # what each step returned, from the device API
    ref.reset()
    for a in acts:
        o, _, r, d, _ = ref.step(torch.tensor(a, dtype=torch.int32, device="cuda"))
        truth.append((o.cpu().numpy().copy(), r.cpu().numpy().copy(), d.cpu().numpy().astype(bool)))
    ve.reset()
    held = []
    for t, a in enumerate(acts[:6]):                        # hold everything: each step must land in its own buffer
        held.append(ve.step(a))
    assert len(ve._slots) >= 6
    for t, (obs, share, rew, done, infos) in enumerate(held):
        assert np.array_equal(obs, truth[t][0]) and np.array_equal(rew[..., 0], truth[t][1]) and np.array_equal(done[..., 0], truth[t][2])
        assert np.array_equal(share[:, 1], obs.reshape(n, -1)) and infos[0]["current_step"] == t + 1
        assert not obs.flags.owndata                        # a view of the pinned slot, not a host copy
    row = held[2][0][3]                                     # a derived view alone keeps its slot alive
    keep_t2 = truth[2][0][3].copy()
    held = None
    n_slots = len(ve._slots)
    for t in range(6, 10):                                  # outputs dropped every step: the pool does not grow
        out = ve.step(acts[t])
        assert np.array_equal(out[0], truth[t][0])
        out = None
    assert len(ve._slots) == n_slots and np.array_equal(row, keep_t2)
    ve.MAX_SLOTS = len(ve._slots)                           # exhaust the pool: copies, still correct and independent
    held = [ve.step(acts[t]) for t in range(10, 14)] + [row]
    extra = [sl["out"]["obs"] for sl in ve._slots]          # every slot referenced from outside
    o1 = ve.step(acts[0])[0]
    assert o1.flags.owndata and len(ve._slots) == ve.MAX_SLOTS
    for t in range(10, 14):
        assert np.array_equal(held[t - 10][0], truth[t][0])
    del extra, held, o1
    ve.copy = False                                         # one buffer, overwritten by the next step
    a0 = ve.step(acts[1])[0]
    before = a0.copy()
    a1 = ve.step(acts[2])[0]
    assert a1 is a0 and not np.array_equal(before, a1)
    ve.close(); ref.close()


@pytest.mark.parametrize("config", ["1v1/NoWeapon/Selfplay", "scenario2/scenario2"])
def test_whole_step_graph_equals_eager_launches(config, monkeypatch):
    """VecEnv.step replays one CUDA graph (pinned H2D of the actions, controller + env kernels, D2H of the packed outputs);
    it returns the same bits as the eager launch sequence, auto-resets included (episodes capped at 6 steps)."""
    import aircombat_selfplay_b200.envs as envs_mod
    orig = envs_mod.load_spec

    def short(*args, **kw):
        spec = orig(*args, **kw)
        spec.max_steps = 6
        return spec
    monkeypatch.setattr(envs_mod, "load_spec", short)
    n = 48
    cls = ShareBatchedVecEnv if config.startswith("scenario") else BatchedVecEnv
    a, b = cls(config, n, seed=5), cls(config, n, seed=5)
    b.use_cuda_graph = False
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    A, D = a.num_agents, a.core.act_dim
    for t in range(14):
        if a.core.hier:
            act = np.concatenate([rng.integers(0, 3, (n, A, 1)), rng.integers(0, 5, (n, A, 1)), rng.integers(0, 3, (n, A, 1)),
                                  rng.integers(0, 2, (n, A, D - 3))], axis=-1)
        else:
            act = np.concatenate([rng.integers(0, 41, (n, A, 3)), rng.integers(0, 30, (n, A, 1))], axis=-1)
        oa, ob = a.step(act), b.step(act)
        for x, y in zip(oa[:-1], ob[:-1]):
            assert np.array_equal(np.asarray(x), np.asarray(y)), t
        assert oa[-1][0]["current_step"] == ob[-1][0]["current_step"] == t % 6 + 1
    assert a._cur["graph"] is not None and all(sl["graph"] is None for sl in b._slots)
    a.close(); b.close()


def test_hierarchical_controller_runs_batched_on_device():
    ve = ShareBatchedVecEnv("scenario2/scenario2", 32, seed=1)
    assert ve.action_space.__class__.__name__ == "Tuple" and ve.core.act_dim == 7
    obs, share = ve.reset()
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = np.concatenate([rng.integers(0, 3, (32, 4, 1)), rng.integers(0, 5, (32, 4, 1)), rng.integers(0, 3, (32, 4, 1)),
                            rng.integers(0, 2, (32, 4, 4))], axis=-1)
        obs, share, rew, done, infos = ve.step(a)
        assert np.isfinite(obs).all() and np.isfinite(rew).all()
    assert ve.core.rnn.abs().sum() > 0
    ve.close()


@pytest.mark.parametrize("config,kind", [("scenario2/scenario2", "pursue"), ("scenario1/scenario1", "maneuver"),
                                         ("scenario1/WVR_selfplay", "pursue")])
def test_scripted_opponents_on_device(config, kind, tmp_path):
    """use_baseline yamls: the red team is flown by the batched PursueAgent / ManeuverAgent (opponents.py); the enemy rows
    of the action array are ignored, as in the reference."""
    import yaml
    from aircombat_selfplay_b200.tasks import parse_config
    cfg = parse_config(config)
    cfg.update({"use_baseline": True, "baseline_type": kind, "use_artillery": True})
    (tmp_path / "vs").mkdir()
    (tmp_path / "vs" / "cfg.yaml").write_text(yaml.safe_dump(cfg, sort_keys=False))
    n = 64
    envs = [BatchedEnv("vs/cfg", n, seed=1, config_dir=str(tmp_path)) for _ in range(2)]
    for e in envs:
        e.reset()
    assert envs[0].opponents is not None and envs[0].opponents.kind == kind
    rng = np.random.default_rng(0)
    A, D = envs[0].n_agents, envs[0].act_dim
    for t in range(15):
        a = rng.integers(0, 2, (n, A, D))
        b = a.copy()
        b[:, A // 2:] = rng.integers(0, 2, (n, A - A // 2, D))      # different garbage in the enemy rows
        o0 = envs[0].step(torch.tensor(a, dtype=torch.int32, device="cuda"))[0].clone()
        o1 = envs[1].step(torch.tensor(b, dtype=torch.int32, device="cuda"))[0].clone()
        assert torch.equal(o0, o1)                                   # enemy action rows do not matter
        assert torch.isfinite(o0).all()
    assert envs[0].rnn.view(n, A, 128)[:, A // 2:].abs().sum() > 0   # the scripted agents' recurrent state is live
    if kind == "maneuver":
        assert int(envs[0].opponents.step.min()) == 15


def test_tacview_render(tmp_path):
    env = SingleCombatEnv("1v1/ShootMissile/Selfplay")
    env.reset()
    path = tmp_path / "rec.txt.acmi"
    for t in range(8):
        env.step(np.array([[20, 19, 20, 10, 1], [20, 19, 20, 10, 1]]))
        env.render(mode="txt", filepath=str(path))
    text = path.read_text(encoding="utf-8-sig").splitlines()
    assert text[0] == "FileType=text/acmi/tacview" and text[2].startswith("0,ReferenceTime=")
    assert sum(l.startswith("#") for l in text) == 8
    a = [l for l in text if l.startswith("A0100,T=")]
    assert len(a) == 8 and "Name=F16" in a[0] and "Color=Blue" in a[0]
    lon, lat, alt = (float(x) for x in a[0].split("T=")[1].split(",")[0].split("|")[:3])
    assert abs(lon - 120.0) < 0.01 and abs(lat - 60.0) < 0.02 and 5000 < alt < 7000
    assert any(l.startswith("A01004,T=") for l in text)          # the missile launched with uid = agent id + remaining count
    with pytest.raises(NotImplementedError):
        env.render(mode="human")


@pytest.mark.parametrize("config", ["1v1/ShootMissile/Selfplay", "scenario2/scenario2"])
def test_env_state_checkpoint_resumes_bit_exactly(config):
    n = 32
    env = BatchedEnv(config, n, seed=5)
    env.reset()
    rng = np.random.default_rng(1)
    A, D = env.n_agents, env.act_dim
    hi = 2 if env.hier else 30
    acts = [torch.tensor(rng.integers(0, hi, (n, A, D)), dtype=torch.int32, device="cuda") for _ in range(24)]
    for a in acts[:12]:
        env.step(a)
    sd = env.state_dict()
    first = [env.step(a)[0].clone() for a in acts[12:]]
    other = BatchedEnv(config, n, seed=5)       # a fresh process would do the same
    other.load_state_dict(sd)
    second = [other.step(a)[0].clone() for a in acts[12:]]
    for x, y in zip(first, second):
        assert torch.equal(x, y)


def test_task_plugin_surface():
    """env.task keeps the reference's Task surface (reference envs/JSBSim/tasks/task_base.py:8-122) as a view of what the
    device computed: reward / termination descriptors in evaluation order, get_obs, normalize_action, get_reward,
    get_termination, _check_missile_warning -- the calls the reference's render scripts make (render_vs_pursue.py:65)."""
    env = SingleCombatEnv("1v1/ShootMissile/Selfplay")
    task = env.task
    assert task.num_agents == 2 and task.observation_space.shape == (21,)
    assert [r.name for r in task.reward_functions] == ["PostureReward", "AltitudeReward", "EventDrivenReward", "ShootPenaltyReward"]
    assert [t.name for t in task.termination_conditions] == ["LowAltitude", "ExtremeState", "Overload", "SafeReturn", "Timeout"]
    alt = task.reward_functions[1]
    assert alt.params == {"safe_altitude": 4.0, "danger_altitude": 3.5, "Kv": 0.2} and not alt.is_potential
    assert task.reward_functions[0].is_potential and task.reward_functions[0].params["orientation_version"] == "v2"
    obs = env.reset()
    assert np.array_equal(task.get_obs(env, "A0100"), obs[0]) and np.array_equal(task.get_obs(env, 1), obs[1])
    np.testing.assert_allclose(task.normalize_action(env, "A0100", [20, 19, 20, 0, 1]), [0.0, -0.05, 0.0, 0.4], atol=1e-15)
    np.testing.assert_allclose(task.normalize_action(env, "B0100", [0, 40, 41, 29, 0]), [-1.0, 1.0, 1.0, 0.9], atol=1e-15)
    assert task._check_missile_warning(env, "B0100") is None
    obs, rew, done, info = env.step(np.array([[20, 19, 20, 10, 1], [20, 19, 20, 10, 0]]))      # A0100 shoots
    w = task._check_missile_warning(env, "B0100")
    assert w is not None and w["shooter"] == 0 and task._check_missile_warning(env, "A0100") is None
    np.testing.assert_allclose(w["position"], env.agents["A0100"].get_position(), atol=300.0)     # one step after launch
    r, _ = task.get_reward(env, "A0100", {})
    assert r == float(rew[0, 0]) and task.get_termination(env, "A0100", {}) == (False, {})
    assert float(task.reward_functions[0].pre_rewards()[0, 0]) != 0.0                            # PostureReward is potential-based
    env.close()
    # a crash: the LowAltitude condition reports it, the aggregate carries the done_condition
    env = SingleCombatEnv("1v1/NoWeapon/Selfplay")
    env.core.set_init_states([[120.0, 60.0, 8400.0, 0.0, 800.0] + [0.0] * 7, [120.0, 60.1, 20000.0, 180.0, 800.0] + [0.0] * 7])
    env.reset()
    for _ in range(40):
        obs, rew, done, info = env.step(np.array([[20, 40, 20, 29], [20, 19, 20, 10]]))
        if done[0]:
            break
    assert done[0] and done[1]                 # A0100 flew into the ground; B0100 is left without a live enemy: SafeReturn
    d, inf = env.task.get_termination(env, "A0100", {})
    assert d and inf["done_condition"] == "low_altitude" == info["done_condition"][0]
    low = env.task.termination_conditions[0]
    assert low.get_termination(env.task, env, "A0100", {})[:2] == (True, False)
    assert env.task.termination_conditions[3].get_termination(env.task, env, "A0100", {})[:2] == (False, False)
    assert env.task.termination_conditions[3].get_termination(env.task, env, "B0100", {})[:2] == (True, True)     # the win
    assert info["done_condition"][1] == "safe_return"
    env.close()
    # hierarchical task: normalize_action previews the controller without advancing its recurrent state
    env = MultipleCombatEnv("scenario2/scenario2")
    env.reset()
    env.step([(np.array([1, 2, 1]), np.array([0, 0, 0, 0]))] * 4)
    h0 = env.core.rnn.clone()
    u = env.task.normalize_action(env, "A0200", np.array([0, 4, 2, 0, 0, 0, 0]))
    assert u.shape == (4,) and (-1 <= u[:3]).all() and (u[:3] <= 1).all() and 0.4 <= u[3] <= 0.9
    assert torch.equal(h0, env.core.rnn) and env.task["hier"] and env.task.descriptor["env"] == "nvn"
    assert len(env.task.reward_functions) == 11 and env.task.reward_functions[-1].name == "ShootPenaltyReward"
    env.close()


def test_render_vs_pursue_style_episode():
    """The loop of the reference's render_vs_pursue.py:47-80 against this package: a scripted PursueAgent red team
    (use_baseline yaml), the ego policy's actions from outside, bloods and infos read back per step, until all done."""
    import yaml
    from aircombat_selfplay_b200.tasks import parse_config
    import tempfile
    cfg = parse_config("scenario1/scenario1_curriculum_vs_pursue")
    cfg["max_steps"] = 60
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "vs.yaml").write_text(yaml.safe_dump(cfg, sort_keys=False))
        env = SingleCombatEnv("vs", config_dir=d)
    num_agents = env.num_agents
    for angle in (0, 60):
        env.core.set_curriculum_angle(angle)          # env.reset_simulators_curriculum(i) of the reference script
        obs = env.reset()
        steps = 0
        while True:
            ego_obs = obs[:num_agents // 2]
            assert np.array_equal(ego_obs[0], env.task.get_obs(env, env.ego_ids[0]))
            ego_actions = np.array([[1, 2, 1, 0, 0, 0, 0]])
            enm_actions = np.zeros((1, 7), dtype=np.int64)           # ignored: the scripted agent flies the red aircraft
            obs, rewards, dones, infos = env.step(np.concatenate([ego_actions, enm_actions], axis=0))
            steps += 1
            bloods = [env.agents[a].bloods for a in env.agents]
            assert len(bloods) == 2 and infos["current_step"] == steps
            if dones.all():
                assert "done_condition" in infos
                break
        assert steps <= 60
    env.close()


def test_device_rollout_matches_the_numpy_runner_path():
    """collect -> step -> insert on CUDA tensors (rollout.DeviceRollout) fills the same buffer as the reference runner's numpy
    path (runner/share_jsbsim_runner.py:157-223 + algorithms/utils/buffer.py:312-350, restated here in numpy over the
    VecEnv API): same observations, rewards, masks, active masks and recurrent-state resets, episode ends included."""
    import aircombat_selfplay_b200.envs as envs_mod
    from aircombat_selfplay_b200.rollout import DeviceRollout
    orig = envs_mod.load_spec

    def short(*args, **kw):
        spec = orig(*args, **kw)
        spec.max_steps = 8
        return spec
    import pytest as _pt
    mp = _pt.MonkeyPatch()
    mp.setattr(envs_mod, "load_spec", short)
    try:
        n, T, cfg = 64, 12, "2v2/NoWeapon/Selfplay"
        env = BatchedEnv(cfg, n, seed=3)
        ve = ShareBatchedVecEnv(cfg, n, seed=3)
    finally:
        mp.undo()
    A, D = env.n_agents, env.spec.obs_dim
    # the first aircraft starts 4 m above the LowAltitude limit: in most envs it is done steps before the env is (active mask 0)
    init = [list(r) for r in env.spec.init_states]
    init[0][2] = 8215.0
    env.set_init_states(init); ve.core.set_init_states(init)

    def policy(share, obs, ha, hc, masks):                     # deterministic toy actor-critic with a recurrent state; element-wise
        x = obs.to(torch.float32)                              # IEEE ops only, so the CPU and the GPU evaluation agree bit for bit
        acts = torch.stack([(x[:, 9 + k].abs() * 997.0).long() % m for k, m in enumerate((41, 41, 41, 30))], dim=-1)
        ha2 = ha * masks.view(-1, 1, 1) + x[:, :1].unsqueeze(-1)
        return x[:, :1], acts, -x[:, 1:2], ha2, hc + 1
    roll = DeviceRollout(env, T)
    roll.warmup()
    done_steps = roll.run(policy)
    assert done_steps == T * n * A
    b = roll.buffer
    # ---- the reference path in numpy
    obs, share = ve.reset()
    Bo = np.zeros((T + 1, n, A, D), np.float32); Bs = np.zeros((T + 1, n, A, A * D), np.float32)
    Br = np.zeros((T, n, A, 1), np.float32); Bm = np.ones((T + 1, n, A, 1), np.float32); Bam = np.ones((T + 1, n, A, 1), np.float32)
    Bha = np.zeros((T + 1, n, A, 1, 128), np.float32); Bact = np.zeros((T, n, A, 4), np.float32)
    Bo[0], Bs[0] = obs, share
    for t in range(T):
        v, acts, lp, ha, hc = policy(torch.from_numpy(np.concatenate(Bs[t])), torch.from_numpy(np.concatenate(Bo[t])),
                                     torch.from_numpy(np.concatenate(Bha[t])), torch.zeros(n * A, 1, 128), torch.from_numpy(np.concatenate(Bm[t])))
        actions = np.array(np.split(acts.numpy(), n))
        ha = np.array(np.split(ha.numpy(), n))
        obs, share, rewards, dones, infos = ve.step(actions)
        d = dones.squeeze(axis=-1)
        de = np.all(d, axis=-1)
        ha[de] = 0
        masks = np.ones((n, A, 1), np.float32); masks[de] = 0
        am = np.ones((n, A, 1), np.float32); am[d] = 0; am[de] = 1
        Bo[t + 1], Bs[t + 1], Br[t], Bm[t + 1], Bam[t + 1], Bha[t + 1], Bact[t] = obs, share, rewards, masks, am, ha, actions
    assert (Bm == 0).any()                                    # episodes ended inside the window
    print('agents done before their env:', int((Bam == 0).sum()))
    for name, ref, got in (("obs", Bo, b.obs), ("share_obs", Bs, b.share_obs), ("rewards", Br, b.rewards), ("masks", Bm, b.masks),
                           ("active_masks", Bam, b.active_masks), ("rnn_states_actor", Bha, b.rnn_states_actor), ("actions", Bact, b.actions)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-6, err_msg=name)
    env.close(); ve.close()


def _poke_status(env, rows_mask, value):
    names, ai = env.batch.arena("ac_i")
    st = ai[names.index("status")].view(env.n_envs, env.n_agents)
    st[rows_mask] = value
    env.batch.set_arena("ac_i", ai)


@pytest.mark.parametrize("rule", [2, 1])
def test_per_env_curriculum_stages_and_win_rate_record(rule):
    """Every env keeps its own curriculum stage and win-rate record on the device (reference: per env process,
    envs/JSBSim/tasks/scenario2_task.py:172-223).  Stage k resets exactly like reset_simulators_curriculum(angles[k]); the
    record is updated when an episode ends and the rule runs before the auto-reset that follows.  rule 2 advances a full
    record above the threshold; rule 1 is the reference's `len(record) > window`, which never fires."""
    cfg, n, angles = "scenario2/scenario2_curriculum", 24, [0, 60, 120]
    env = BatchedEnv(cfg, n, seed=1, curriculum_rule=rule, curriculum_window=3, curriculum_threshold=0.6)
    env.set_curriculum_stages(angles)
    stages = torch.arange(n, device="cuda") % 3
    env.set_env_stages(stages)
    obs = env.reset()[0].clone()
    ref_obs = []
    for a in angles:                                   # the batch-wide stage API gives the reference reset of that angle
        r = BatchedEnv(cfg, 1, seed=1)
        r.set_curriculum_angle(a)
        ref_obs.append(r.reset()[0][0].clone())
        r.close()
    for k in range(3):
        assert torch.equal(obs[stages == k], ref_obs[k].unsqueeze(0).expand(int((stages == k).sum()), -1, -1)), k
    assert not torch.equal(ref_obs[0], ref_obs[1])
    # episodes: envs 0..7 win (both enemies down -> the ego team ends alive through SafeReturn), 8..15 lose, the rest fly on
    env.set_env_stages(torch.zeros(n, dtype=torch.int32, device="cuda"))
    env.reset()
    act = torch.zeros((n, 4, 7), dtype=torch.int32, device="cuda")
    act[..., 0:3] = 1
    win = torch.zeros((n, 4), dtype=torch.bool, device="cuda"); win[0:8, 2:] = True
    lose = torch.zeros((n, 4), dtype=torch.bool, device="cuda"); lose[8:16, :2] = True
    for episode in range(1, 5):
        _poke_status(env, win | lose, 2)
        obs, _, rew, done, info = env.step(act)
        assert bool(env.batch.env_done[:16].all()) and not bool(env.batch.env_done[16:].any()), episode
        stage, wins, count = env.curriculum_state()
        if rule == 2 and episode >= 3:                 # third win: record full, rate 1.0 >= 0.6 -> stage 1, record cleared
            assert stage[:8].tolist() == [1] * 8 and count[:8].tolist() == [episode - 3] * 8
            if episode == 3:                           # the reset inside that very step already used the new stage
                assert torch.equal(obs[:8], ref_obs[1].unsqueeze(0).expand(8, -1, -1))
        else:
            assert stage[:8].tolist() == [0] * 8 and wins[:8].tolist() == [min(episode, 3)] * 8
        assert stage[8:16].tolist() == [0] * 8 and wins[8:16].tolist() == [0] * 8 and count[8:16].tolist() == [min(episode, 3)] * 8
        assert stage[16:].tolist() == [0] * 8 and count[16:].tolist() == [0] * 8
    env.close()
