"""CPU: the env-layer oracle (oracle/env_oracle.py) against golden trajectories of the REFERENCE's own Python env layer
(tests/golden/env_*.npz, written by tools/make_golden.py by importing /root/reference with its missing third-party
modules shimmed).  This pins the restatement of the reference's tasks, observation packing, rewards (quirks included),
terminations, missile, chaff, gun and artillery code; the FDM below both sides is the same CPU oracle FDM, so these
vectors do NOT pin the FDM arithmetic.

Tolerance: both sides run the same fp64 arithmetic in a different order of Python operations -> 1e-9 on observations,
1e-9 on rewards, flags bit-exact."""
import json
from pathlib import Path

import numpy as np
import pytest

from aircombat_selfplay_b200.tasks import build_spec
from oracle import env_oracle as eo

# the *_vs_* files drive the red team with scripted agents: tests/test_opponents_golden.py
GOLDEN = sorted(p for p in (Path(__file__).resolve().parent / "golden").glob("env_*.npz") if "_vs_" not in p.name)


@pytest.fixture()
def chaff_draws_pinned(monkeypatch):
    """The generator replaced np.random.rand() in the chaff test (reference env_base.py:153) by a recorded sequence -- 0.5
    everywhere in the older files, values on both sides of 0.85 in the *_chaff_mixed files (``chaff_draws``).  The
    returned function pins the oracle's keyed chaff draw to replay that sequence in call order; a different number or
    order of draws than the reference made shows up as a trajectory mismatch or as unconsumed draws."""
    real = eo.u01
    state = {"seq": None, "k": 0}

    def u01(seed, env, purpose, a=0, b=0, c=0):
        if purpose != eo.RNG_CHAFF:
            return real(seed, env, purpose, a, b, c)
        k = state["k"]
        state["k"] += 1
        if state["seq"] is None:
            return 0.5
        assert k < len(state["seq"]), "the oracle draws more chaff numbers than the reference did"
        return float(state["seq"][k])
    monkeypatch.setattr(eo, "u01", u01)
    return state


def test_golden_files_exist():
    assert len(GOLDEN) >= 12


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_reproduces_reference_trajectory(path, chaff_draws_pinned):
    g = np.load(path, allow_pickle=False)
    cfg = json.loads(str(g["config"]))
    spec = build_spec(cfg)
    if "chaff_draws" in g.files:
        chaff_draws_pinned["seq"] = g["chaff_draws"]
    env = eo.OracleEnv(spec, seed=int(g["seed"]), env_index=0)
    obs, share = env.reset()
    assert obs.shape == g["obs"][0].shape
    np.testing.assert_allclose(obs, g["obs"][0], rtol=0, atol=1e-9)
    T = g["actions"].shape[0]
    for t in range(T):
        obs, share, rew, done, info = env.step(g["actions"][t])
        np.testing.assert_allclose(obs, g["obs"][t + 1], rtol=1e-9, atol=1e-9, err_msg=f"{path.stem} obs step {t}")
        np.testing.assert_allclose(rew, g["rewards"][t], rtol=1e-9, atol=1e-9, err_msg=f"{path.stem} reward step {t}")
        assert np.array_equal(done, g["dones"][t]), (path.stem, t, done, g["dones"][t])
        assert [s.status for s in env.sims] == g["status"][t].tolist(), (path.stem, t)
        np.testing.assert_allclose([s.bloods for s in env.sims], g["bloods"][t], rtol=0, atol=1e-9)
        if "share_obs" in g.files:
            np.testing.assert_allclose(share, g["share_obs"][t], rtol=1e-9, atol=1e-9)
        assert len(env.missiles) == int(g["n_missiles"][t]), (path.stem, t)
        assert sum(c.count for c in env.chaffs) == int(g["n_chaffs"][t]), (path.stem, t)
    if "turn_counts" in g.files:
        assert env.heading_turn_counts == int(g["turn_counts"])
    if "chaff_draws" in g.files:
        assert chaff_draws_pinned["k"] == len(g["chaff_draws"]), (path.stem, chaff_draws_pinned["k"], len(g["chaff_draws"]))


def test_chaff_golden_covers_both_sides_of_the_decoy_threshold():
    """At least one reference trajectory has chaff draws below AND at / above 0.85 (a decoyed missile and a missile that
    flies through the cloud), so the `< 0.85` comparison and the repeated per-substep draws are pinned."""
    seen = False
    for p in GOLDEN:
        g = np.load(p, allow_pickle=False)
        if "chaff_draws" in g.files and len(g["chaff_draws"]):
            d = g["chaff_draws"]
            seen = seen or ((d < 0.85).any() and (d >= 0.85).any() and (d == 0.85).any())
    assert seen
