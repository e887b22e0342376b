"""Ad-hoc GPU env-layer parity report (the pytest -m gpu tests assert the same quantities).  Usage:
python tests/run_env_parity.py [config ...]"""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from aircombat_selfplay_b200.tasks import load_spec
from tests.env_parity import CONFIGS, Pair, mask_degenerate_sides, close_init_states, compare_step, low_init_states, random_actions


def run(name, n_envs=6, steps=40, mode="random", close=False, substeps=None, seed=3, low=False):
    spec = load_spec(name, substeps_override=substeps)
    rng = np.random.default_rng(seed)
    init = close_init_states(spec, rng) if close else (low_init_states(spec) if low else None)
    p = Pair(spec, n_envs, seed=seed, init_states=init)
    (g_obs, g_share), (c_obs, c_share) = p.reset()
    g_obs = mask_degenerate_sides(g_obs, c_obs, p.cpu)
    e0 = np.max(np.abs(g_obs - c_obs) / np.maximum(1, np.abs(c_obs)))
    nbad, first, events = 0, None, set()
    worst_obs = worst_rew = 0.0
    for t in range(steps):
        act = random_actions(rng, spec, n_envs, mode=mode)
        g, c = p.step(act)
        bad = compare_step(g, c, spec, envs=p.cpu)
        worst_obs = max(worst_obs, float(np.nanmax(np.abs(g["obs"] - c["obs"]) / np.maximum(1, np.abs(c["obs"])))))
        worst_rew = max(worst_rew, float(np.nanmax(np.abs(g["rew"] - c["rew"]))))
        for e in p.cpu:
            for m in e.missiles.values():
                events.add({0: "launched", 1: "HIT", 2: "MISS"}[m.status])
            if e.chaffs:
                events.add("chaff")
        for cz in np.unique(c["cause"]):
            if cz >= 0:
                events.add(f"term{cz}")
        if bad:
            nbad += 1
            if first is None:
                first = (t, bad)
    tag = f"{name} mode={mode} close={close} K={spec.substeps}"
    print(f"[{'OK ' if nbad == 0 else 'BAD'}] {tag}: reset err {e0:.2e}, worst obs {worst_obs:.2e}, worst rew {worst_rew:.2e}, "
          f"bad steps {nbad}/{steps}, events {sorted(events)}")
    if first:
        print("      first mismatch at step", first[0])
        for b in first[1]:
            print("        ", b)
    return nbad == 0


if __name__ == "__main__":
    names = sys.argv[1:] or CONFIGS
    ok = True
    t0 = time.time()
    for n in names:
        ok &= run(n)
        if n != "singlecontrol/heading":
            ok &= run(n, close=True, mode="smooth", steps=60)
        ok &= run(n, mode="dive", steps=60, n_envs=3, low=True)
    ok &= run("1v1/NoWeapon/Selfplay", substeps=6)
    print("ALL OK" if ok else "MISMATCHES", f"{time.time() - t0:.1f}s")
    sys.exit(0 if ok else 1)
