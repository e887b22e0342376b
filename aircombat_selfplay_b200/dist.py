"""Multi-GPU plumbing: one process per GPU, each owning a contiguous slice of the job's environments.

The env step has NO data-path collective -- environments are independent (the reference runs each in its own OS
process, reference envs/env_wrappers.py:231-267).  ``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU
tests) is used for exactly the two exchanges the reference's runners perform across environments:
  * episode statistics summed over all envs at a log interval (reference runner/share_jsbsim_runner.py:124-138);
  * the self-play opponent's actor weights pushed to every rollout worker when the opponent changes
    (reference runner/share_jsbsim_runner.py:400-407, runner/selfplay_jsbsim_runner.py:236-260).
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
    """(world_size, rank, local_rank) from the torchrun environment (1, 0, 0 when not launched distributed)."""
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None):
    ws, rank, local = world()
    if ws > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=ws)
    return ws, rank, local


def env_slice(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [offset, offset + count) of the job's envs owned by ``rank``; the first ``total % world`` ranks
    get one extra env.  The offset is what ``AcsTaskConfig.env_offset`` takes, so per-env RNG streams do not depend on
    how the job is sharded."""
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def make_sharded_vec_env(config_name: str, total_envs: int, share: bool | None = None, seed: int = 0, **kw):
    """This rank's ``BatchedVecEnv`` / ``ShareBatchedVecEnv`` over its slice of ``total_envs``."""
    from .env_wrappers import BatchedVecEnv, ShareBatchedVecEnv
    from .tasks import load_spec
    ws, rank, local = world()
    offset, count = env_slice(total_envs, ws, rank)
    if share is None:
        share = load_spec(config_name, kw.get("config_dir")).share_obs
    cls = ShareBatchedVecEnv if share else BatchedVecEnv
    return cls(config_name, count, device=local, seed=seed, env_offset=offset, **kw)


def all_reduce_stats(stats: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum of per-rank episode statistics (finished episodes, reward sums, step counts, ...) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    keys = sorted(stats)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def broadcast_module(module: torch.nn.Module, src: int = 0):
    """Broadcast every parameter and buffer of ``module`` from ``src`` (opponent-pool weight push), as one flat bucket
    per dtype: a ~50k-parameter actor is a single small collective, sized for launch latency rather than bandwidth."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype: Dict[torch.dtype, list] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dt, ts_ in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts_])
        dist.broadcast(flat, src=src)
        o = 0
        for t in ts_:
            n = t.numel()
            t.copy_(flat[o:o + n].view_as(t))
            o += n
