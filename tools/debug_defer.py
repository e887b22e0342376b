import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from aircombat_selfplay_b200.capi import EnvBatch
from aircombat_selfplay_b200.tasks import load_spec
from tests.env_parity import close_init_states, random_actions
name = sys.argv[1] if len(sys.argv) > 1 else "1v1/ShootMissile/Selfplay"
os.environ["ACS_FRAME_SPLIT"] = "0"
spec = load_spec(name); spec.max_steps = 120
n = 300
bs = []
for on in ("1", "0"):
    os.environ["ACS_DEFER_MISSILES"] = on
    b = EnvBatch(spec, n, seed=3)
    b.set_init_states(close_init_states(spec, np.random.default_rng(2)))
    b.reset(); bs.append(b)
rng = np.random.default_rng(9)
acts_log = []
for t in range(150):
    act_np = random_actions(rng, spec, n, mode="smooth", shoot_p=0.3)
    acts_log.append(act_np)
    act = torch.tensor(act_np, device="cuda")
    pre = [{k: b.arena(k)[1].clone() for k in ("ms_i", "ms_d", "ac_i", "env_i")} for b in bs]
    ra = [None if x is None else x.clone() for x in bs[0].step(act, auto_reset=True)]
    rb = bs[1].step(act, auto_reset=True)
    bad = (ra[4] != rb[4]).any(-1).any(-1) | (ra[3] != rb[3]).any(-1) | ((ra[0] - rb[0]).abs() > 1e-9).any(-1).any(-1)
    if bool(bad.any()):
        e = int(torch.nonzero(bad)[0])
        A, S = spec.n_agents, max(1, spec.n_missile_slots)
        print("step", t, "first bad env", e, "n bad", int(bad.sum()))
        print("info defer ", ra[4][e].tolist(), "\ninfo lockst", rb[4][e].tolist())
        print("dones", ra[3][e].tolist(), rb[3][e].tolist())
        names, ei = bs[0].arena("env_i")
        print("mode (deferred flag) of env:", int(ei[names.index("deferred")][e]), "faults", int(ei[names.index("faults")][e]))
        mn = bs[0].arena("ms_i")[0]
        for tag, b, pr in (("defer", bs[0], pre[0]), ("lockstep", bs[1], pre[1])):
            mi = b.arena("ms_i")[1].view(len(mn), n, A, S)[:, e]
            md = b.arena("ms_d")[1].view(-1, n, A, S)[:, e]
            print(tag, "ms_i post", {k: mi[i].tolist() for i, k in enumerate(mn)})
            print(tag, "ms_i pre ", {k: pr["ms_i"].view(len(mn), n, A, S)[i, e].tolist() for i, k in enumerate(mn)})
            print(tag, "ms pos_n t d_prev post", md[0].tolist(), md[9].tolist(), md[13].tolist())
            print(tag, "status pre", pr["ac_i"].view(-1, n, A)[0, e].tolist(), "post", b.arena("ac_i")[1].view(-1, n, A)[0, e].tolist())
        print("obs diff max", float((ra[0][e] - rb[0][e]).abs().max()))
        from oracle.env_oracle import OracleEnv
        o = OracleEnv(spec, seed=3, env_index=e)
        o.init_states = [list(r) for r in close_init_states(spec, np.random.default_rng(2))]
        o.reset()
        for tt in range(t + 1):
            out = o.step(acts_log[tt][e])
            if tt >= t - 1:
                print("oracle step", tt, "status", [s_.status for s_ in o.sims], "done", out[3].tolist(), "cause", out[4]["done_cause"],
                      "missiles", [(k, m.status, round(m.distance_pre, 3)) for k, m in o.missiles.items()])
            if out[3].all():
                print("oracle episode ended at", tt); break
        break
else:
    print("no mismatch in 150 steps")
names, ei = bs[0].arena("env_i")
print("faults total", int(ei[names.index("faults")].sum()))
