"""CPU tests of the host layer: yaml -> TaskSpec resolution, spaces, lazy infos, and the batched low-level controller
against golden outputs of the reference's BaselineActor (tests/golden/controller.npz, tools/make_golden_controller.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from aircombat_selfplay_b200 import spaces, taskspec as ts
from aircombat_selfplay_b200.controller import LowLevelController, hierarchical_input
from aircombat_selfplay_b200.tasks import TASKS, load_spec, parse_config

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_controller_matches_reference_baseline_actor():
    g = np.load(GOLDEN / "controller.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    ctl = LowLevelController()
    ctl.load_reference_state_dict(sd)
    ctl.eval()
    h = torch.zeros(g["x"].shape[1], 128)
    for t in range(g["x"].shape[0]):
        a, h = ctl(torch.from_numpy(g["x"][t]), h)
        assert np.array_equal(a.numpy(), g["actions"][t]), t          # arg-max classes: exact
        np.testing.assert_allclose(h.numpy(), g["h"][t], rtol=0, atol=2e-6)   # fp32 GRU state


def test_hierarchical_input_layout():
    obs = torch.arange(2 * 15, dtype=torch.float64).reshape(2, 15) / 100
    obs[0, 0] = 0.5      # 2500 m  -> forced climb in the 1v1 task
    obs[1, 0] = 1.2      # 6000 m
    high = torch.tensor([[2, 0, 1], [2, 4, 0]])
    x = hierarchical_input(high, obs, force_climb_below_m=3500.0)
    assert x.dtype == torch.float32 and x.shape == (2, 12)
    assert x[0, 0] == pytest.approx(0.1) and x[1, 0] == pytest.approx(-0.1)        # singlecombat_task.py:235-239
    assert x[0, 1] == pytest.approx(-np.pi / 6) and x[1, 1] == pytest.approx(np.pi / 6)
    assert x[0, 2] == pytest.approx(0.0) and x[1, 2] == pytest.approx(0.05)
    assert torch.allclose(x[:, 3:], obs[:, :9].float())
    x2 = hierarchical_input(high, obs, None)
    assert x2[0, 0] == pytest.approx(-0.1)                                          # multiplecombat_task.py:173


@pytest.mark.parametrize("name,A,D,act,share", [
    ("singlecontrol/heading", 1, 12, 4, False), ("1v1/NoWeapon/Selfplay", 2, 15, 4, False),
    ("1v1/ShootMissile/Selfplay", 2, 21, 5, False), ("2v2/NoWeapon/Selfplay", 4, 27, 4, True),
    ("2v2/ShootMissile/HierarchySelfplay", 4, 33, 5, True), ("scenario1/scenario1", 2, 21, 8, False),
    ("scenario2/scenario2", 4, 21, 8, True), ("scenario2/scenario2_nvn", 4, 39, 8, True),
    ("scenario3/scenario3", 8, 21, 8, True), ("scenario3/scenario3_nvn", 8, 63, 8, True)])
def test_task_resolution(name, A, D, act, share):
    sp = load_spec(name)
    assert sp.n_agents == A and sp.obs_dim == D and 4 + sp.shoot_dim == act and sp.share_obs == share
    assert len(sp.init_states) == A and all(len(r) == 12 for r in sp.init_states)
    assert sp.substeps == parse_config(name)["agent_interaction_steps"]
    if sp.share_obs:      # MultipleCombatEnv order: rewards, team mean, then dones (multiplecombat_env.py:163-180)
        assert not sp.dones_before_rewards and sp.team_mean and sp.terminations[0] == ts.T_SAFE_RETURN
    else:
        assert sp.dones_before_rewards and not sp.team_mean


def test_missile_slots_cover_every_launch_of_an_episode():
    sp = load_spec("scenario2/scenario2")
    assert sp.n_missile_slots == 2 * max(sp.num_missiles)      # separate AIM-9M and AIM-120B counters
    assert load_spec("1v1/ShootMissile/Selfplay").n_missile_slots == 4
    assert load_spec("1v1/NoWeapon/Selfplay").n_missile_slots == 0


def test_unknown_task_raises_like_the_reference():
    from aircombat_selfplay_b200.tasks import build_spec
    with pytest.raises(NotImplementedError):
        build_spec({"task": "nope", "aircraft_configs": {}})
    with pytest.raises(FileNotFoundError):
        load_spec("no/such/config")


def test_spaces_and_action_layout():
    from aircombat_selfplay_b200.envs import action_space_for
    s = action_space_for("scenario2")
    assert s.__class__.__name__ == "Tuple" and spaces.flat_action_dim(s) == 7
    assert spaces.flat_action_dim(action_space_for("heading")) == 4
    assert action_space_for("hierarchical_multiplecombat_shoot_nearest").nvec.tolist() == [3, 5, 3, 2]
    assert spaces.flat_action_dim(action_space_for("singlecombat_shoot")) == 5
    box = spaces.Box(low=-10, high=10.0, shape=(21,))
    assert box.shape == (21,)
    for name in TASKS:
        assert spaces.flat_action_dim(action_space_for(name)) in (3, 4, 5, 7, 8)


def test_lazy_info_is_a_dict_view():
    from aircombat_selfplay_b200.envs import LazyInfo
    info = np.zeros((3, 2, 4), dtype=np.int32)
    info[..., 0] = -1
    src = {"info": info, "heading": False}
    infos = np.empty(3, dtype=object)
    for i in range(3):
        infos[i] = LazyInfo(src, i)
    info[1, :, 2] = 7
    info[1, 0, 0] = ts.T_LOW_ALTITUDE
    assert isinstance(infos[0], dict) and infos.shape[0] == 3
    assert infos[1]["current_step"] == 7 and "done_condition" in infos[1] and infos[1]["done_condition"][0] == "low_altitude"
    assert "done_condition" not in infos[0] and "heading_turn_counts" not in infos[0]
    assert dict(infos[1].items())["current_step"] == 7
    # copies and serialisation see the content (a dict subclass with empty storage would yield {})
    import json
    import pickle
    full = {"current_step": 7, "done_condition": ["low_altitude", ""]}
    assert dict(infos[1]) == full and infos[1].copy() == full and {**infos[1]} == full and (lambda **kw: kw)(**infos[1]) == full
    assert json.loads(json.dumps(infos[1])) == full and pickle.loads(pickle.dumps(infos[1])) == full and len(infos[1]) == 2
    # heading task: the key the runner collects appears only when UnreachHeading ended the episode (unreach_heading.py:60-62)
    hsrc = {"info": info, "heading": True}
    info[2, 0, 3] = 4
    assert "heading_turn_counts" not in LazyInfo(hsrc, 2)
    info[2, 0, 0] = ts.T_UNREACH_HEADING
    assert LazyInfo(hsrc, 2)["heading_turn_counts"] == 4
    info[2, 0, 0] = ts.T_TIMEOUT
    assert "heading_turn_counts" not in LazyInfo(hsrc, 2)


def test_shipped_configs_match_the_reference_yamls():
    """Every yaml of the reference resolves to the same TaskSpec as the shipped file of the same name (only where the
    reference tree is mounted: the build container)."""
    import subprocess
    import sys
    ref = Path("/root/reference/envs/JSBSim/configs")
    if not ref.exists():
        pytest.skip("reference tree not mounted")
    r = subprocess.run([sys.executable, str(Path(__file__).resolve().parents[1] / "tools" / "check_configs.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count("[ok]") >= 39


def test_every_shipped_config_resolves():
    from aircombat_selfplay_b200.tasks import CONFIG_DIR
    names = [f"{f.parent.relative_to(CONFIG_DIR)}/{f.stem}" for f in sorted(CONFIG_DIR.glob("**/*.yaml"))]
    assert len(names) >= 45
    for n in names:
        sp = load_spec(n)
        assert sp.obs_dim > 0 and len(sp.rewards) > 0 and len(sp.terminations) >= 4


def test_missing_controller_checkpoint_is_loud(monkeypatch, tmp_path):
    """No baseline_model.pt -> AcsError (the reference's torch.load raises too); a named but missing file -> FileNotFoundError;
    random-init weights only behind the explicit opt-in."""
    import pytest
    from aircombat_selfplay_b200 import controller
    from aircombat_selfplay_b200.capi import AcsError
    monkeypatch.delenv("ACS_ALLOW_RANDOM_CONTROLLER", raising=False)
    monkeypatch.delenv("ACS_BASELINE_MODEL", raising=False)
    if controller.find_checkpoint() is None:
        with pytest.raises(AcsError):
            controller.make_controller("cpu")
        assert controller.make_controller("cpu", allow_random=True).checkpoint is None
    with pytest.raises(FileNotFoundError):
        controller.make_controller("cpu", path=str(tmp_path / "nope.pt"))
    monkeypatch.setenv("ACS_BASELINE_MODEL", str(tmp_path / "nope.pt"))
    with pytest.raises(FileNotFoundError):
        controller.make_controller("cpu")
    # a reference checkout next to the yaml files is found through config_dir
    g = np.load(GOLDEN / "controller.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    (tmp_path / "model").mkdir()
    (tmp_path / "configs").mkdir()
    torch.save(sd, tmp_path / "model" / "baseline_model.pt")
    monkeypatch.delenv("ACS_BASELINE_MODEL")
    ctl = controller.make_controller("cpu", config_dir=str(tmp_path / "configs"))
    assert ctl.checkpoint == str(tmp_path / "model" / "baseline_model.pt")
