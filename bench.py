#!/usr/bin/env python
"""bench.py -- agent-steps/s of the env-step hot path (60 Hz sim, 12 substeps per step) on N B200s.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line on rank 0.
  value         whole-job agent-steps/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e           the same metric through the reference-facing VecEnv API with HOST numpy actions in / obs,reward,done out
  roofline      the dominant kernel (k_env_substeps) against the measured HBM peak, plus the fp64-pipe view
  cpu_baseline  the CPU oracle port of the same workload on the host cores (rank 0, N=1 only; bounded sample)
`--impl reference` times the reference-shaped CPU path instead (the oracle port: Python env layer over the C++ FDM,
one env per worker process on all host cores) -- see DESIGN.md section 6 for why the real JSBSim cannot run here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "agent-steps/sec (60 Hz sim, 12 substeps/step)"
UNIT = "agent-steps/s"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}



def _dist_init(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if (args.impl == "b200") else "gloo", rank=rank, world_size=world)
    return world, rank, local


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _max_over_ranks(x: float, world, device) -> float:
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------- workloads
WORKLOADS = {
    # name: (description, envs per GPU, agents per env)
    "fdm_only": ("F-16 FDM substep loop only (no task layer) -- bring-up workload, NOT the headline metric", 4096, 2),
}


def cpu_fdm_baseline(n_aircraft: int, steps: int, substeps: int, seconds_budget: float = 15.0):
    """CPU oracle port on all host cores: one OracleFdm per aircraft, threads release the GIL inside the C call."""
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    from oracle.fdm import OracleFdm
    cores = os.cpu_count() or 1
    n = min(n_aircraft, 8 * cores)
    fdms = [OracleFdm() for _ in range(n)]
    rng = np.random.default_rng(0)
    for f in fdms:
        f.reset(psi_deg=float(rng.uniform(0, 360)), h_sl_ft=float(rng.uniform(15000, 28000)))
        f.set_controls(0.1, -0.1, 0.0, 0.7)
    chunks = [fdms[k::cores] for k in range(cores)]

    def work(chunk):
        for f in chunk:
            f.run(substeps)
    done, t0 = 0, time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        while True:
            list(ex.map(work, chunks))
            done += 1
            el = time.perf_counter() - t0
            if done >= steps or el > seconds_budget:
                break
    return {"value": n * done / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} aircraft x {done} steps x {substeps} substeps, CPU oracle FDM (restated JSBSim F-16 path), {cores} threads"}


def run_b200(args):
    import numpy as np
    import torch
    world, rank, local = _dist_init(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the simulator has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from aircombat_selfplay_b200.capi import FdmBatch
    desc, n_envs, n_agents = WORKLOADS[args.workload]
    if args.envs:
        n_envs = args.envs
    K = 12
    rows = n_envs * n_agents
    rng = np.random.default_rng(rank)
    ic = np.zeros((rows, 12))
    ic[:, 0] = 120.0; ic[:, 1] = 60.0 + 0.1 * (np.arange(rows) % 2)
    ic[:, 2] = 20000.0; ic[:, 3] = 180.0 * (np.arange(rows) % 2); ic[:, 4] = 800.0
    fb = FdmBatch(n_envs, n_agents, device=local)
    fb.reset(torch.tensor(ic, device=dev))
    total = args.warmup + args.steps
    acts_host = np.column_stack([rng.integers(0, 41, (total * rows, 3)) / 20.0 - 1.0, rng.integers(0, 30, total * rows) / 58.0 + 0.4]) \
        .reshape(total, rows, 4)
    acts_dev = torch.tensor(acts_host, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_dev(t):
        fb.set_controls(acts_dev[t])
        fb.run(K)

    for t in range(args.warmup):
        step_dev(t)
    torch.cuda.synchronize(); _barrier(world)
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(); step_dev(args.warmup + k); ev[k][1].record()
    torch.cuda.synchronize(); _barrier(world)
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    ms = _max_over_ranks(ms, world, dev)
    value = world * rows * args.steps / (ms * 1e-3)

    # e2e: host numpy actions in, host outputs out, through the public host API
    pin_in = torch.empty((rows, 4), dtype=torch.float64).pin_memory()
    n_out = len(fb.output_names)
    pin_out = torch.empty((n_out, rows), dtype=torch.float64).pin_memory()
    u_dev = torch.empty((rows, 4), dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); _barrier(world)
    t0 = time.perf_counter()
    for k in range(args.steps):
        pin_in.copy_(torch.from_numpy(acts_host[args.warmup + k]))
        u_dev.copy_(pin_in, non_blocking=True)
        fb.set_controls(u_dev); fb.run(K)
        pin_out.copy_(fb.get_outputs(), non_blocking=True)
        torch.cuda.synchronize()
    _barrier(world)
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e = {"value": world * rows * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": rows * 4 * 8, "d2h_bytes_per_step": n_out * rows * 8}

    peak, peak_src = _peaks()
    n_state = len(fb.state_names)
    bytes_per_agent_step = (2 * n_state + n_out) * 8 + 4 * 8
    kernel_ms = ms / args.steps
    ach = rows * bytes_per_agent_step / (kernel_ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "envs_per_gpu": n_envs, "agents_per_env": n_agents, "substeps": K, "sim_freq": 60,
                       "l2": "flushed between timed steps (256 MB fill)"},
            "e2e": e2e, "gpu_launches": 2 * args.steps, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_agent_step": bytes_per_agent_step,
                         "note": "fp64-pipe/latency bound, see DESIGN.md"}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_fdm_baseline(rows, 5, K)
        print(json.dumps(line), flush=True)


def run_reference(args):
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, n_envs, n_agents = WORKLOADS[args.workload]
    K = 12
    t0 = time.perf_counter()
    b = cpu_fdm_baseline(n_envs * n_agents, args.steps + args.warmup, K, seconds_budget=60.0)
    line = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": desc, "substeps": K, "sim_freq": 60},
            "cpu_baseline": b, "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's own size)")
    ap.add_argument("--workload", default="fdm_only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
