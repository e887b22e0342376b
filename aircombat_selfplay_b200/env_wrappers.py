"""The VecEnv contract of the reference (reference envs/env_wrappers.py:48-462) over the batched device simulator.

The reference runs N environments as N OS processes (``SubprocVecEnv`` / ``ShareSubprocVecEnv``) or N Python objects
(``DummyVecEnv`` / ``ShareDummyVecEnv``) and stacks their numpy results.  Here the N environments ARE one device batch:

    reset() -> obs [N, A, D]                                   (Share*: (obs, share_obs [N, A, A*D]))
    step(actions [N, A, act]) -> obs, rewards [N, A, 1], dones [N, A, 1], infos [N] of dict
                                                               (Share*: obs, share_obs, rewards, dones, infos)
    auto-reset of an env whose agents are all done, the reset observation returned in place
    (reference envs/env_wrappers.py:191-204,380-393) -- done on the device, inside the step call.

The four reference class names are kept, constructor signature included (a list of env factories): the first factory
is called once to learn the scenario (it must return one of this package's env classes), the rest only contribute
their count.  ``BatchedVecEnv`` / ``ShareBatchedVecEnv`` are the direct constructors.  One step moves the action array
host->device and one packed output buffer device->host (pinned memory, one copy each way).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch

from . import taskspec as ts
from .capi import AcsError
from .envs import BatchedEnv, LazyInfo, _SingleEnvBase


class VecEnv(ABC):
    """reference envs/env_wrappers.py:48-121"""
    closed = False

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def step_async(self, actions):
        pass

    @abstractmethod
    def step_wait(self):
        pass

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


class ShareVecEnv(VecEnv):
    """reference envs/env_wrappers.py:325-336"""

    def __init__(self, num_envs, observation_space, share_observation_space, action_space):
        super().__init__(num_envs, observation_space, action_space)
        self.share_observation_space = share_observation_space


class BatchedVecEnv(VecEnv):
    """N environments of one scenario on one GPU behind the reference's VecEnv contract."""
    share = False

    def __init__(self, config_name: str, num_envs: int, device: int = 0, seed: int = 0, env_offset: int = 0,
                 config_dir: Optional[str] = None, substeps: Optional[int] = None, controller_path: Optional[str] = None,
                 copy: bool = True, allow_random_controller: bool = False):
        self.core = BatchedEnv(config_name, num_envs, device=device, seed=seed, env_offset=env_offset, config_dir=config_dir,
                               substeps=substeps, controller_path=controller_path, auto_reset=True,
                               allow_random_controller=allow_random_controller)
        if self.share:
            ShareVecEnv.__init__(self, num_envs, self.core.observation_space, self.core.share_observation_space,
                                 self.core.action_space)
        else:
            VecEnv.__init__(self, num_envs, self.core.observation_space, self.core.action_space)
        self.num_agents = self.core.num_agents
        self.copy = copy
        with torch.cuda.device(self.core.device):
            self._act_host = torch.empty((num_envs, self.num_agents, self.core.act_dim), dtype=torch.int32).pin_memory()
            self._act_np = self._act_host.numpy()
            self._act_dev = self.core._act_in          # the step graph's static input buffer: H2D lands directly in it
        # Output slots.  A step lands its packed outputs in a pinned host buffer with ONE D2H copy and returns numpy views of
        # it.  ``copy=True`` (the default) promises what the reference gives: arrays that stay valid for as long as the caller
        # holds them.  Instead of copying every array out of a single buffer (three host copies per step, a quarter of the
        # end-to-end step at 4096 envs), the wrapper keeps a small pool of pinned buffers and never writes into one whose
        # arrays are still referenced from outside (CPython reference counts of the arrays it handed out, and of the views
        # the caller derived from them -- a derived view keeps its parent alive).  A runner that drops last step's arrays when
        # it takes the next ones cycles through two slots.  Past ``MAX_SLOTS`` live buffers the arrays are copied out instead.
        # ``copy=False``: always slot 0, valid until the next step.
        self._slots = []
        self._cur, self._copy_out = self._new_slot(), False
        self._new_slot()                               # a caller that holds last step's arrays across a step alternates between two
        self._pending = False
        # the whole host-facing step -- pinned H2D of the actions, (controller +) env kernels, D2H of the packed outputs --
        # replays as ONE CUDA graph (one per output slot): one launch per step instead of a copy, four kernels and a copy
        self.use_cuda_graph = True
        self.h2d_bytes_per_step = self._act_host.numel() * 4
        self.d2h_bytes_per_step = self._cur["host"].numel()

    MAX_SLOTS = 8

    # ------------------------------------------------------------------ output slots
    def _build_slot(self):
        b, N, A = self.core.batch, self.num_envs, self.num_agents
        with torch.cuda.device(self.core.device):
            host = torch.empty(b.out_buf.shape, dtype=torch.uint8).pin_memory()
        views = b.host_views(host)
        src = {"info": views["info"], "heading": self.core.spec.obs_kind == ts.OBS_HEADING}
        infos = np.empty(N, dtype=object)
        for i in range(N):
            infos[i] = LazyInfo(src, i)
        obs = views["obs"]
        out = {"obs": obs, "rewards": views["rewards"].reshape(N, A, 1),
               "dones": views["dones"].view(np.bool_).reshape(N, A, 1),      # the kernel writes 0 / 1: a view, not a conversion
               "infos": infos}
        if self.share:
            # share_obs [N, A, A*D]: every agent's row is the concatenation of all agents' observations (reference
            # envs/JSBSim/envs/env_base.py:183-189), a read-only stride-0 view of ``obs`` -- never copied over PCIe or
            # materialised; the runner's buffer insert makes the one copy it needs
            D = obs.shape[-1]
            out["share"] = np.broadcast_to(obs.reshape(N, 1, A * D), (N, A, A * D))
        # every array a caller can end up holding, directly or as the parent of a view it derived
        return {"host": host, "views": views, "out": out, "graph": None, "epoch": -1, "watched": list(views.values()) + list(out.values())}

    @staticmethod
    def _refcounts(slot):
        import sys
        return [sys.getrefcount(x) for x in slot["watched"]]

    def _new_slot(self):
        slot = self._build_slot()                      # (its locals are gone: only the slot references the arrays now)
        slot["baseline"] = self._refcounts(slot)
        self._slots.append(slot)
        return slot

    @classmethod
    def _slot_free(cls, slot):
        return cls._refcounts(slot) == slot["baseline"]

    def _pick_slot(self):
        """The buffer this step writes: slot 0 for ``copy=False``; otherwise one whose arrays nobody outside holds."""
        if not self.copy:
            return self._slots[0], False
        for slot in self._slots:
            if self._slot_free(slot):
                return slot, False
        if len(self._slots) < self.MAX_SLOTS:
            return self._new_slot(), False
        if getattr(self, "_scratch", None) is None:       # every slot is held by the caller: land in a buffer that is never
            self._scratch = self._build_slot()            # handed out, and copy the arrays out of it
        return self._scratch, True

    @property
    def _views(self):
        return self._cur["views"]

    @property
    def _host(self):
        return self._cur["host"]

    @property
    def _infos(self):
        return self._cur["out"]["infos"]

    # ------------------------------------------------------------------ helpers
    def _fetch(self):
        self._cur["host"].copy_(self.core.batch.out_buf, non_blocking=True)
        torch.cuda.current_stream(self.core.device).synchronize()

    def _put_actions(self, actions, h2d=True):
        a = self._act_np
        if isinstance(actions, np.ndarray) and actions.shape == a.shape:
            np.copyto(a, actions, casting="unsafe")
        else:   # nested lists of per-agent actions, Tuple-space samples included (reference tests pass these)
            for i, env_act in enumerate(actions):
                for j, x in enumerate(env_act):
                    if isinstance(x, (tuple, list)):
                        x = np.concatenate([np.atleast_1d(np.asarray(y)).ravel() for y in x])
                    a[i, j, :] = np.asarray(x).ravel()
        if h2d:
            self._act_dev.copy_(self._act_host, non_blocking=True)

    def _outputs(self, keys):
        o = self._cur["out"]
        if self._copy_out:      # pool exhausted: the reference's semantics by copying (infos stay views of slot 0)
            c = {k: (o[k].copy() if k in ("obs", "rewards", "dones") else o[k]) for k in keys}
            if "share" in keys:
                N, A, D = c["obs"].shape
                c["share"] = np.broadcast_to(c["obs"].reshape(N, 1, A * D), (N, A, A * D))
            return tuple(c[k] for k in keys)
        return tuple(o[k] for k in keys)

    # ------------------------------------------------------------------ VecEnv
    def reset(self):
        self._cur, self._copy_out = self._pick_slot()
        with torch.cuda.device(self.core.device):
            self.core.reset()
            self._fetch()
        return self._outputs(("obs", "share")) if self.share else self._outputs(("obs",))[0]

    def _capture(self, slot):
        core = self.core
        core._warm_for_capture()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._act_dev.copy_(self._act_host, non_blocking=True)
            core._step_body(self._act_dev)
            slot["host"].copy_(core.batch.out_buf, non_blocking=True)
        slot["graph"], slot["epoch"] = g, core._epoch

    def step_async(self, actions):
        core = self.core
        self._cur, self._copy_out = self._pick_slot()
        slot = self._cur
        with torch.cuda.device(core.device):
            if self.use_cuda_graph and not core._timing:
                if not core._was_reset:
                    raise AcsError("step() called before reset()")
                self._put_actions(actions, h2d=False)
                if slot["graph"] is None or slot["epoch"] != core._epoch:
                    self._capture(slot)
                slot["graph"].replay()
                self._graphed = True
            else:
                self._put_actions(actions)
                core.step(self._act_dev)
                self._graphed = False
        self._pending = True

    def step_wait(self):
        assert self._pending, "step_wait() without step_async()"
        with torch.cuda.device(self.core.device):
            if self._graphed:      # the D2H copy is the graph's last node
                torch.cuda.current_stream(self.core.device).synchronize()
            else:
                self._fetch()
        self._pending = False
        return self._outputs(("obs", "share", "rewards", "dones", "infos") if self.share else ("obs", "rewards", "dones", "infos"))

    def render(self, mode, filepath):
        """reference DummyVecEnv.render (envs/env_wrappers.py:170-172): the TacView text log of env 0."""
        if mode == "txt":
            if getattr(self, "_acmi", None) is None:
                from .tacview import AcmiWriter
                self._acmi = AcmiWriter(self.core, 0)
            step = int(self._views["info"][0, 0, 2])
            self._acmi.write(filepath, step * self.core.time_interval)

    def close_extras(self):
        self.core.close()


class ShareBatchedVecEnv(BatchedVecEnv, ShareVecEnv):
    share = True


def _from_env_fns(cls, env_fns, **kw):
    env_fns = list(env_fns)
    assert len(env_fns) > 0
    template = env_fns[0]()
    if not isinstance(template, _SingleEnvBase):
        raise TypeError("env factories must return aircombat_selfplay_b200 env classes (SingleControlEnv, SingleCombatEnv, "
                        "MultipleCombatEnv)")
    core = template.core
    args = dict(config_name=core.config_name, num_envs=len(env_fns), device=core.device.index or 0,
                seed=core.seed_value, substeps=core.spec.substeps)
    args.update(kw)
    template.close()
    return cls(**args)


class DummyVecEnv(BatchedVecEnv):
    """reference envs/env_wrappers.py:123-181 -- same constructor: ``DummyVecEnv([env_fn, ...])``"""

    def __new__(cls, env_fns, **kw):
        return _from_env_fns(BatchedVecEnv, env_fns, **kw)


class SubprocVecEnv(BatchedVecEnv):
    """reference envs/env_wrappers.py:231-322 -- ``SubprocVecEnv([env_fn, ...], context='spawn', in_series=1)``; the
    process arguments are accepted and ignored: there are no worker processes."""

    def __new__(cls, env_fns, context="spawn", in_series=1, **kw):
        return _from_env_fns(BatchedVecEnv, env_fns, **kw)


class ShareDummyVecEnv(ShareBatchedVecEnv):
    """reference envs/env_wrappers.py:339-372"""

    def __new__(cls, env_fns, **kw):
        return _from_env_fns(ShareBatchedVecEnv, env_fns, **kw)


class ShareSubprocVecEnv(ShareBatchedVecEnv):
    """reference envs/env_wrappers.py:414-462"""

    def __new__(cls, env_fns, context="spawn", in_series=1, **kw):
        return _from_env_fns(ShareBatchedVecEnv, env_fns, **kw)
