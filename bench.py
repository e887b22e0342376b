#!/usr/bin/env python
"""bench.py -- agent-steps/s of the env-step hot path (60 Hz sim, 12 substeps per step) on N B200s.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE JSON line on rank 0.
  value         whole-job agent-steps/s with inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e           the same metric through the reference-facing VecEnv API: HOST numpy actions in, host obs / rewards /
                dones out, both copies inside the timed region (default copy=True API; `views`: copy=False;
                `device_rollout`: the zero-copy runner loop on CUDA tensors, no PCIe)
  roofline      the dominant kernel (k_env_substeps / _split / _split3, by batch size) against the fp64 peak measured live
                (primary: the kernel is fp64-pipe / issue bound), with the measured-HBM-peak view under `hbm`
  workloads     short samples of the other BASELINE.json configs (65 536-env 1v1, 1v1 ShootMissile, 2v2 ShootMissile =
                north-star config, 4v4) with their per-kernel times and roofline fractions
  cpu_baseline  the CPU oracle port of the same workload on the host cores (rank 0, N=1 only; bounded sample)
`--impl reference` times the reference-shaped CPU path instead: the restated Python env layer over the restated C++
FDM, one worker process per host core -- DESIGN.md section 3 explains why the real JSBSim cannot run here.
Envs shard across GPUs as contiguous slices with no data-path collective ("scaling": "weak").
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
SUBSTEPS = 12

# name -> (yaml config, envs per GPU, what BASELINE.json calls it)
WORKLOADS = {
    "1v1_noweapon": ("1v1/NoWeapon/Selfplay", 4096, "SingleCombat 1v1/NoWeapon/Selfplay, 4096 envs on 1xB200 (BASELINE.json configs[1])"),
    "1v1_shoot": ("1v1/ShootMissile/Selfplay", 16384, "SingleCombat 1v1/ShootMissile/Selfplay, 16384 envs (configs[2])"),
    "2v2_shoot": ("2v2/ShootMissile/HierarchySelfplay", 8192, "MultipleCombat 2v2/ShootMissile, 65536 envs over 8 GPUs = 8192 per GPU (configs[3])"),
    "4v4": ("scenario3/scenario3", 4096, "MultipleCombat 4v4 with missiles, 8 agents/env (configs[4])"),
    "heading": ("singlecontrol/heading", 32, "SingleControl heading, 32 envs (configs[0])"),
}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel: str, config: str, n_envs: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this exact workload (profiles/traffic.json, written next to the ncu summaries); null when there is none."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text()).get(f"{kernel}|{config}|{n_envs}")
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a
    thread every 2 ms (the timed region of the small configs lasts tens of milliseconds, shorter than one `nvidia-smi
    -lms` period); `nvidia-smi` is the fallback when NVML cannot be loaded."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.h = index, [], None, None, None
        self.sm, self.mask, self.mx, self._stop = [], 0, None, threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except TypeError:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [x for x in vis.split(",") if x.strip().isdigit()]
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def poll_once(self):
        """One sample now (called by the bench while the timed launches are still in flight, so that even a timed region
        shorter than the polling period is sampled under load)."""
        n = self.nvml
        if n is None:
            return
        try:
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
            self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            pass

    def _poll(self):
        while not self._stop.is_set():
            self.poll_once()
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1)
            n, names = self.nvml, []
            for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
                bit = getattr(n, attr, None) or getattr(n, attr.replace("ClocksEventReason", "ClocksThrottleReason"), 0)
                if bit and (self.mask & bit):
                    names.append(name)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": names, "samples": len(sm),
                    "source": "nvml, 2 ms period"}
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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


def _dist_init(backend):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend, rank=rank, world_size=world)
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


def synthetic_actions(rng, spec, high_dim, n_envs, n_steps):
    """i.i.d. uniform ints over each MultiDiscrete dim, shoot bits Bernoulli(0.05)."""
    import numpy as np
    A = spec.n_agents
    a = np.zeros((n_steps, n_envs, A, high_dim + spec.shoot_dim), dtype=np.int32)
    nvec = [41, 41, 41, 30] if high_dim == 4 else [3, 5, 3]
    for k, n in enumerate(nvec):
        a[..., k] = rng.integers(0, n, (n_steps, n_envs, A))
    if spec.shoot_dim:
        a[..., high_dim:] = rng.random((n_steps, n_envs, A, spec.shoot_dim)) < 0.05
    return a


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    config, n_envs, warm_rounds, n_rounds, seed, budget_s, min_s = args
    import numpy as np
    from aircombat_selfplay_b200.tasks import load_spec
    from oracle.env_oracle import OracleEnv
    spec = load_spec(config, substeps_override=SUBSTEPS)
    envs = [OracleEnv(spec, seed=seed, env_index=i) for i in range(n_envs)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(seed)
    acts = synthetic_actions(rng, spec, 4, n_envs, 64)    # low-level rows: the oracle sits below the GRU controller

    def one_round(t):
        for i, e in enumerate(envs):
            _, _, _, d, _ = e.step(acts[t % 64, i])
            if d.all():
                e.reset()
    for t in range(warm_rounds):
        one_round(t)
    t0 = time.perf_counter()
    done = 0
    while True:
        one_round(warm_rounds + done)
        done += 1
        el = time.perf_counter() - t0
        if el > budget_s or (done >= n_rounds and el >= min_s):
            break
    return n_envs * spec.n_agents * done, time.perf_counter() - t0


def cpu_env_baseline(config: str, rounds: int, warm_rounds: int = 2, budget_s: float = 20.0, envs_per_core: int = 2, min_s: float = 0.0):
    """The oracle port of the same workload on every host core: one process per core, ``envs_per_core`` envs each (the
    reference's SubprocVecEnv layout), synthetic actions of the same distribution, K = 12.  Each worker times its own
    step loop (env construction and the first reset are outside) for at least ``min_s`` seconds and ``rounds`` env-steps,
    at most ``budget_s`` seconds; the rate is total agent-steps / slowest worker."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(config, envs_per_core, warm_rounds, rounds, 100 + k, budget_s, min_s) for k in range(cores)])
    agent_steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    rates = sorted(r[0] / r[1] for r in res)
    return {"value": agent_steps / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "per_worker_agent_steps_per_s": {"min": rates[0], "median": rates[len(rates) // 2], "max": rates[-1]},
            "sample": f"{cores} worker processes x {envs_per_core} envs of {config}, {agent_steps} agent-steps in {wall:.1f} s, K={SUBSTEPS}; "
                      "restated Python env layer over the restated C++ F-16 FDM (not JSBSim itself, which cannot run here)"}


# ----------------------------------------------------------------------------- B200 arm
KERNELS = ("k_env_substeps", "k_env_substeps_split", "k_env_substeps_split3", "k_env_substeps_split4")
# fp64 work of one aircraft substep: 2*DFMA + DMUL + DADD thread instructions the substep kernel EXECUTES per aircraft, from
# the committed ncu capture of the kernels as they are now (profiles/r2b_k_env_substeps_65536envs.txt: 23 487 per aircraft per
# 12-substep step; it was 24 770 / 25 081 before the guard-free division / sqrt sequences of csrc/fmath.cuh removed Newton
# steps).  Re-counted whenever the frame's arithmetic changes, so that `achieved` never credits work the kernel no longer
# does; missile / chaff arithmetic is NOT counted (conservative).
FLOPS_PER_SUBSTEP = 1957.0
MISSILE_SLOT_BYTES = 2 * 16 * 8 + 8 * 4 + 2 * 4     # 16 doubles read + written, 8 ints read, 2 ints written per live slot


def workload_config(desc, config, n_envs, A, hier):
    return {"workload": desc, "scenario": config, "envs_per_gpu": n_envs, "agents_per_env": A, "substeps": SUBSTEPS,
            "sim_freq": 60, "auto_reset": True, "l2": "flushed between timed steps (256 MB fill)",
            "controller": ("batched GRU low-level controller in PyTorch, random-init weights of the reference architecture"
                           if hier else "none (direct stick/throttle classes)")}


def measure(config, n_envs, steps, warmup, seed, world, rank, local, dev, flush, fp64_peak, e2e_views=True):
    """One workload on this rank's GPU: device-resident rate, per-kernel times, end-to-end rate through the VecEnv API
    (host numpy in / out), both rooflines of the substep kernel.  Every timing is the max over ranks."""
    import numpy as np
    import torch
    from aircombat_selfplay_b200 import capi
    from aircombat_selfplay_b200.env_wrappers import BatchedVecEnv, ShareBatchedVecEnv
    from aircombat_selfplay_b200.tasks import load_spec
    cls = ShareBatchedVecEnv if load_spec(config).share_obs else BatchedVecEnv
    # one env slice per GPU: rank r owns envs [r*n_envs, (r+1)*n_envs) of the job (RNG streams keyed by global env index)
    ve = cls(config, n_envs, device=local, seed=seed, env_offset=rank * n_envs, substeps=SUBSTEPS,
             allow_random_controller=True)     # synthetic weights of the reference's architecture (no checkpoint travels)
    core, batch, spec = ve.core, ve.core.batch, ve.core.spec
    A = spec.n_agents
    rng = np.random.default_rng(seed + rank)
    total = warmup + steps
    acts_host = synthetic_actions(rng, spec, core.high_dim, n_envs, total)
    acts_dev = torch.from_numpy(acts_host).to(dev)

    # ---- device-resident arm: the step (controller + env kernels) replays as one CUDA graph
    core.reset()
    for t in range(warmup):
        core.step(acts_dev[t])
    torch.cuda.synchronize(); _barrier(world)
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record(); core.step(acts_dev[warmup + k]); ev[k][1].record()
    sampler.poll_once()                     # the launches above run ahead of the device: this sample is under load
    torch.cuda.synchronize(); _barrier(world)
    clocks = sampler.stop()
    ms = _max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), world, dev)
    value = world * n_envs * A * steps / (ms * 1e-3)
    # ---- per-kernel breakdown of the same step (eager launches bracketed by CUDA events on the launching stream)
    core.set_timing(True)
    nk = min(steps, 50)
    for k in range(nk):
        flush.fill_(k & 0xFF)
        core.step(acts_dev[warmup + k])
    kms, ksteps = batch.get_timing()
    core.set_timing(False)
    live_frac, n_slots = 0.0, 0
    if spec.n_missile_slots and spec.launch_kind:
        names, mi = batch.arena("ms_i")
        live_frac = float((mi[names.index("status")] == 0).float().mean())
        n_slots = int(spec.n_missile_slots)

    # ---- end to end through the VecEnv contract (host numpy in / out, the default API: fresh arrays every step)
    def run_e2e(v):
        out = v.reset()
        for t in range(warmup):
            out = v.step(acts_host[t])       # held across the next call, as a runner does: the output-slot pool reaches its steady state
        torch.cuda.synchronize(); _barrier(world)
        done = 0
        t0 = time.perf_counter()
        for k in range(steps):
            out = v.step(acts_host[warmup + k])
            done += int(np.count_nonzero(out[-2]))        # finished AGENT episodes (a reduction over [N, A, 1] per axis costs 70 us of host time)
        _barrier(world)
        s_ = _max_over_ranks(time.perf_counter() - t0, world, dev)
        return {"value": world * n_envs * A * steps / s_, "unit": UNIT, "h2d_bytes_per_step": v.h2d_bytes_per_step,
                "d2h_bytes_per_step": v.d2h_bytes_per_step, "ms_per_step": s_ / steps * 1e3}, done
    e2e, done_envs = run_e2e(ve)
    e2e["api"] = "VecEnv.step(numpy) with the default copy=True: obs / rewards / dones stay valid for as long as the caller holds them (views of a pool of pinned output slots; a slot is rewritten only when nothing references its arrays)"
    if e2e_views:
        ve.copy = False
        e2e["views"], _ = run_e2e(ve)
        e2e["views"]["api"] = "copy=False: one pinned D2H buffer, the returned views are valid until the next step"
        ve.copy = True

    # ---- the zero-copy runner path (rollout.DeviceRollout: collect -> BatchedEnv.step -> insert on CUDA tensors, a torch replay
    # buffer, nothing crosses PCIe; reference runner/share_jsbsim_runner.py:157-223).  The policy is a stub that hands out
    # the same synthetic action rows (no network: the learner side is out of scope), so this is the env + buffer cost.
    if e2e_views:
        from aircombat_selfplay_b200.rollout import DeviceRollout
        ro = DeviceRollout(core, T=32)
        ro.warmup()
        cursor = [0]
        zeros1 = torch.zeros(n_envs * A, 1, device=dev)

        def stub_policy(share, obs, ha, hc, masks):
            a = acts_dev[cursor[0] % total].view(n_envs * A, -1)
            cursor[0] += 1
            return zeros1, a, zeros1, ha, hc
        ro.run(stub_policy, warmup)
        torch.cuda.synchronize(); _barrier(world)
        t0 = time.perf_counter()
        ro.run(stub_policy, steps)
        torch.cuda.synchronize(); _barrier(world)
        s_ = _max_over_ranks(time.perf_counter() - t0, world, dev)
        e2e["device_rollout"] = {"value": world * n_envs * A * steps / s_, "unit": UNIT, "ms_per_step": s_ / steps * 1e3,
                                 "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                 "api": "rollout.DeviceRollout.run: collect (stub policy) -> BatchedEnv.step -> DeviceRolloutBuffer.insert, "
                                        "all on CUDA tensors; wall clock around the loop, L2 not flushed"}

    # ---- rooflines of the dominant kernel (k_env_substeps*): DESIGN.md section 5
    peak, peak_src = _peaks()
    n_state, n_out = len(capi.state_field_names()), len(capi.output_field_names())
    base_bytes = (2 * n_state + n_out + 11 + 11) * 8 + 2 * 4 + (4 + spec.shoot_dim) * 4
    bytes_per_agent_step = base_bytes + n_slots * live_frac * MISSILE_SLOT_BYTES
    k_ms = kms["substeps"] / max(ksteps, 1)
    kname = KERNELS[batch.get_option("frame_split_effective")]
    flops = FLOPS_PER_SUBSTEP * SUBSTEPS
    ach_tf = n_envs * A * flops / (k_ms * 1e-3) / 1e12
    ach_gb = n_envs * A * bytes_per_agent_step / (k_ms * 1e-3) / 1e9
    roof = {"bound": "fp64", "kernel": kname, "achieved": ach_tf, "peak": (fp64_peak or 0.0) / 1e12, "unit": "TFLOP/s",
            "frac": (ach_tf / (fp64_peak / 1e12)) if fp64_peak else None,
            "peak_source": "measured live (acs_bench_fp64_peak: dependent-free DFMA loop on every SM; MEASURED_PEAKS.json has no fp64 figure)",
            "flops_per_agent_step": flops, "flops_note": "fp64 flops the substep kernel executes for the FDM (ncu: 2*DFMA + DMUL + DADD of one frame x 12, current kernels); missile arithmetic not counted",
            "traffic": _ncu_traffic(kname, config, n_envs), "kernel_ms": k_ms,
            "kernel_share_of_step": kms["substeps"] / max(sum(kms.values()), 1e-12),
            "missiles_ms": kms.get("missiles", 0.0) / max(ksteps, 1), "post_ms": kms["post"] / max(ksteps, 1),
            "reset_ms": kms["reset"] / max(ksteps, 1),
            "hbm": {"achieved": ach_gb, "peak": peak, "unit": "GB/s", "frac": ach_gb / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_agent_step": bytes_per_agent_step, "missile_slots": n_slots,
                    "missile_live_fraction": live_frac,
                    "note": "secondary view: fusing the K substeps leaves the kernel fp64-pipe / issue bound, not HBM bound"}}
    res = {"value": value, "ms_per_step": ms / steps, "e2e": e2e, "roofline": roof, "clocks": clocks, "A": A, "hier": bool(core.hier),
           "gpu_launches": batch.get_option("launches_per_step") * steps, "agent_episodes_finished_in_e2e": done_envs,
           "missile_live_fraction": live_frac}
    ve.close()
    return res


def run_b200(args):
    import torch
    world, rank, local = _dist_init("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the simulator has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from aircombat_selfplay_b200 import capi
    config, n_envs, desc = WORKLOADS[args.workload]
    if args.envs:
        n_envs = args.envs
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    fp64_peak = None
    if not args.no_fp64_peak:
        try:
            fp64_peak = capi.fp64_peak_flops(local)
        except Exception:          # measurement helper only
            fp64_peak = None
    m = measure(config, n_envs, args.steps, args.warmup, args.seed, world, rank, local, dev, flush, fp64_peak)
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(desc, config, n_envs, m["A"], m["hier"]),
            "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "clocks": m["clocks"], "roofline": m["roofline"],
            "agent_episodes_finished_in_e2e": m["agent_episodes_finished_in_e2e"],
            "parity": "FDM parity is against the restated CPU oracle (unpinned: no runnable JSBSim here; only its atmosphere model is pinned by JSBSim's own reference data); the env layer is pinned "
                      "by golden trajectories of the reference's own Python (DESIGN.md section 3)"}
    # ---- the other BASELINE.json configurations, short samples of the same measurement (headline stays configs[1])
    if not args.no_workloads and args.workload == "1v1_noweapon" and not args.envs:
        extra = {}
        for name, (cfg_w, n_w, desc_w) in [("1v1_noweapon_65536", ("1v1/NoWeapon/Selfplay", 65536, "configs[1] at 65536 envs per GPU")),
                                            ("1v1_shoot", WORKLOADS["1v1_shoot"]), ("2v2_shoot", WORKLOADS["2v2_shoot"]),
                                            ("4v4", WORKLOADS["4v4"])]:
            try:
                w = measure(cfg_w, n_w, args.extra_steps, args.extra_warmup, args.seed, world, rank, local, dev, flush, fp64_peak,
                            e2e_views=False)
                r = w["roofline"]
                extra[name] = {"config": workload_config(desc_w, cfg_w, n_w, w["A"], w["hier"]), "value": w["value"],
                               "ms_per_step": w["ms_per_step"], "e2e": w["e2e"], "kernel": r["kernel"], "kernel_ms": r["kernel_ms"],
                               "missiles_ms": r["missiles_ms"], "post_ms": r["post_ms"], "reset_ms": r["reset_ms"], "roofline_fp64_frac": r["frac"],
                               "roofline_hbm_frac": r["hbm"]["frac"], "missile_live_fraction": w["missile_live_fraction"],
                               "steps": args.extra_steps, "warmup": args.extra_warmup}
            except Exception as exc:
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
        extra["north_star"] = "2v2_shoot"
        line["workloads"] = extra
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_env_baseline(config, rounds=10 ** 9, budget_s=args.cpu_seconds)
        print(json.dumps(line), flush=True)


def run_reference(args):
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config, n_envs, desc = WORKLOADS[args.workload]
    if args.envs:
        n_envs = args.envs
    from aircombat_selfplay_b200.tasks import TASKS, load_spec
    spec = load_spec(config)
    # every "step" of this arm is a bounded sample of the workload: `ref_rounds` env-steps of (host cores x 2) envs;
    # W warm-up steps untimed, then at least K steps AND at least `ref_min_seconds` of wall clock timed (a 0.3 s sample
    # moved the rate by +-30 %), the whole run capped at 150 s
    b = cpu_env_baseline(config, rounds=args.steps * args.ref_rounds, warm_rounds=args.warmup * args.ref_rounds, budget_s=150.0,
                         min_s=args.ref_min_seconds)
    hier = bool(TASKS[spec.name]["hier"])
    line = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(desc, config, n_envs, spec.n_agents, hier),
            "cpu_baseline": b, "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's own size)")
    ap.add_argument("--workload", default="1v1_noweapon", choices=sorted(WORKLOADS))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp64-peak", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="wall budget of the CPU baseline sample")
    ap.add_argument("--ref-rounds", type=int, default=10, help="--impl reference: env-steps per bench step")
    ap.add_argument("--ref-min-seconds", type=float, default=8.0, help="--impl reference: minimum timed wall clock per worker")
    ap.add_argument("--no-workloads", action="store_true", help="headline workload only (skip the `workloads` samples)")
    ap.add_argument("--extra-steps", type=int, default=40, help="timed steps of each `workloads` sample")
    ap.add_argument("--extra-warmup", type=int, default=60, help="warm-up steps of each `workloads` sample (missiles in flight)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    try:
        main()
    finally:
        try:
            import torch.distributed as _dist
            if _dist.is_available() and _dist.is_initialized():
                _dist.destroy_process_group()
        except Exception:
            pass
