"""CPU: the batched scripted opponents (aircombat_selfplay_b200/opponents.py) + low-level controller against golden
trajectories of the reference's own PursueAgent / ManeuverAgent driving its BaselineActor (use_baseline yamls;
tests/golden/env_*_vs_*.npz from tools/make_golden.py).  The simulator below both is the CPU oracle env, so this pins the
host-side logic: request computation, controller wiring, recurrent-state handling, the red team's shoot bits."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from aircombat_selfplay_b200.controller import LowLevelController
from aircombat_selfplay_b200.opponents import RuleOpponents
from aircombat_selfplay_b200.tasks import build_spec
from oracle import env_oracle as eo

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
CASES = sorted(GOLDEN_DIR.glob("env_*_vs_*.npz"))


class OracleState:
    """DeviceState's accessors over one oracle env (n_envs = 1)."""

    def __init__(self, env):
        self.env = env

    def _t(self, x):
        return torch.tensor(np.asarray(x, dtype=np.float64)).reshape(1, -1)

    def pos(self, i):
        return self._t(self.env.sims[i].position)

    def vel(self, i):
        return self._t(self.env.sims[i].velocity)

    def u(self, i):
        return self._t([self.env.sims[i].uvw_mps[0]]).reshape(1)

    def heading(self, i):
        return self._t([self.env.sims[i].posture[2]]).reshape(1)

    def altitude(self, i):
        return self._t([self.env.sims[i].h_sl_m]).reshape(1)

    def ego9(self, i):
        return self._t(self.env._ego9(self.env.sims[i]))


def test_cases_exist():
    assert len(CASES) >= 4


@pytest.mark.parametrize("path", CASES, ids=[p.stem for p in CASES])
def test_scripted_opponents_reproduce_reference(path, monkeypatch):
    real = eo.u01
    monkeypatch.setattr(eo, "u01", lambda seed, env, purpose, a=0, b=0, c=0: 0.5 if purpose == eo.RNG_CHAFF else real(seed, env, purpose, a, b, c))
    g = np.load(path, allow_pickle=False)
    cfg = json.loads(str(g["config"]))
    spec = build_spec(cfg)
    assert spec.use_baseline and spec.baseline_type in ("pursue", "maneuver")
    c = np.load(GOLDEN_DIR / "controller.npz")
    ctl = LowLevelController()
    ctl.load_reference_state_dict({k[3:]: torch.from_numpy(c[k]) for k in c.files if k.startswith("sd:")})
    ctl.eval()
    env = eo.OracleEnv(spec, seed=int(g["seed"]), env_index=0)
    obs, _ = env.reset()
    np.testing.assert_allclose(obs, g["obs"][0], rtol=0, atol=1e-9)
    opp = RuleOpponents(spec.baseline_type, spec.env_kind, spec.n_ego, spec.n_enm, spec.substeps / spec.sim_freq,
                        OracleState(env), 1, torch.device("cpu"))
    rnn = torch.zeros(spec.n_enm, 128)
    for t in range(g["actions"].shape[0]):
        act = g["actions"][t].copy()
        x = opp.inputs().view(-1, 12)
        low, rnn = ctl(x, rnn)
        act[spec.n_ego:, :4] = low.numpy()
        if spec.shoot_dim:
            act[spec.n_ego:, 4:] = 1 if spec.use_artillery else 0
        obs, share, rew, done, info = env.step(act)
        np.testing.assert_allclose(obs, g["obs"][t + 1], rtol=1e-9, atol=1e-9, err_msg=f"{path.stem} obs step {t}")
        np.testing.assert_allclose(rew, g["rewards"][t], rtol=1e-9, atol=1e-9, err_msg=f"{path.stem} reward step {t}")
        assert np.array_equal(done, g["dones"][t]), (path.stem, t)
        assert [s.status for s in env.sims] == g["status"][t].tolist(), (path.stem, t)
        assert len(env.missiles) == int(g["n_missiles"][t]), (path.stem, t)
