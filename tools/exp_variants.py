"""Tuning experiment: per-kernel times of several library builds (ACS_LIB) on the BASELINE workloads.

usage: python tools/exp_variants.py [--splits 0,-1] [--cases 1v1:4096,1v1:65536,...] lib1.so lib2.so ...
Each (lib, case, split) runs in its own subprocess (a library is dlopen'ed once per process); rounds are interleaved so
that slow drift of the box hits every variant alike.  "default" = the in-tree libacs.so.
"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CONFIGS = {"1v1": "1v1/NoWeapon/Selfplay", "1v1_shoot": "1v1/ShootMissile/Selfplay", "2v2": "2v2/NoWeapon/Selfplay",
           "2v2_shoot": "2v2/ShootMissile/HierarchySelfplay", "4v4": "scenario3/scenario3", "scenario2": "scenario2/scenario2"}


def child(case: str, n_envs: int, split: int, steps: int, warm: int) -> None:
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import torch
    from aircombat_selfplay_b200.capi import EnvBatch
    from aircombat_selfplay_b200.tasks import load_spec
    spec = load_spec(CONFIGS[case], substeps_override=12)
    b = EnvBatch(spec, n_envs, seed=0)
    if split >= 0:
        b.set_option("frame_split", split)
    b.reset()
    rng = np.random.default_rng(0)
    A = spec.n_agents
    n = steps + warm
    acts = torch.tensor(np.concatenate([rng.integers(0, 41, (n, n_envs, A, 3)), rng.integers(0, 30, (n, n_envs, A, 1)),
                                        (rng.random((n, n_envs, A, spec.shoot_dim)) < 0.05).astype(np.int64)], axis=-1).astype(np.int32),
                        device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for t in range(warm):
        b.step(acts[t], auto_reset=True)
    torch.cuda.synchronize()
    b.set_timing(True)
    for t in range(steps):
        if not os.environ.get("EXP_NO_FLUSH"):
            flush.fill_(t & 0xff)      # cold L2, as bench.py times it
        b.step(acts[warm + t], auto_reset=True)
    ms, k = b.get_timing()
    b.close()
    print(json.dumps({"sub": ms["substeps"] / k, "mis": ms.get("missiles", 0.0) / k, "post": ms["post"] / k, "reset": ms["reset"] / k, "A": A}))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--splits", default="0")
    ap.add_argument("--cases", default="1v1:4096,1v1:16384,1v1:65536,2v2_shoot:8192,1v1_shoot:16384,4v4:4096")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=30)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--child", nargs=3)
    a = ap.parse_args()
    if a.child:
        child(a.child[0], int(a.child[1]), int(a.child[2]), a.steps, a.warm)
        return
    libs = ["default"] + a.libs
    cases = [(c.split(":")[0], int(c.split(":")[1])) for c in a.cases.split(",")]
    splits = [int(s) for s in a.splits.split(",")]
    best = {}
    for rnd in range(a.rounds):
        for lib in libs:
            for case, n in cases:
                for split in splits:
                    env = dict(os.environ)
                    path, *kv = lib.split("@")          # "lib.so@KEY=VAL@KEY2=VAL2": environment of that variant
                    for x in kv:
                        env[x.split("=")[0]] = x.split("=", 1)[1]
                    if path != "default":
                        env["ACS_LIB"] = str(Path(path).resolve())
                    r = subprocess.run([sys.executable, __file__, "--steps", str(a.steps), "--warm", str(a.warm), "--child", case, str(n), str(split)],
                                       env=env, capture_output=True, text=True, timeout=600)
                    key = (lib.split("/")[-1], case, n, split)
                    if r.returncode != 0:
                        best[key] = {"error": r.stderr.strip().splitlines()[-1:]}
                        continue
                    d = json.loads(r.stdout.strip().splitlines()[-1])
                    if key not in best or d["sub"] < best[key]["sub"]:
                        best[key] = d
    for (lib, case, n, split), d in sorted(best.items(), key=lambda kv: (kv[0][1], kv[0][2], kv[0][3], kv[0][0])):
        if "error" in d:
            print(f"{case:>10} x{n:<6} split={split:>2} {lib:<28} ERROR {d['error']}")
            continue
        tot = d["sub"] + d.get("mis", 0.0) + d["post"] + d["reset"]
        print(f"{case:>10} x{n:<6} split={split:>2} {lib:<28} sub {d['sub']:.4f} missiles {d.get('mis', 0.0):.4f} post {d['post']:.4f} reset {d['reset']:.4f} ms"
              f" -> {n * d['A'] / tot / 1e3:7.1f} M agent-steps/s", flush=True)


if __name__ == "__main__":
    main()
