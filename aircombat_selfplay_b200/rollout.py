"""Zero-copy rollout path: the runner's ``collect -> envs.step -> insert`` loop on CUDA tensors.

The reference's runners move everything through numpy (``_t2n`` after the policy, numpy replay buffers,
reference runner/share_jsbsim_runner.py:12,157-223 and algorithms/utils/buffer.py:26-170,270-350); with the env step on
the device that forces one PCIe round trip and a chain of host copies per step.  ``DeviceRolloutBuffer`` is the same
buffer (same fields, shapes ``[T(+1), N, A, ...]``, same mask / active-mask / recurrent-state rules as
``ShareJSBSimRunner.insert``) held in device memory, and ``DeviceRollout.run`` the loop that fills it from
``BatchedEnv.step``: observations never leave HBM, the policy consumes and produces device tensors, nothing
synchronises with the host until the caller reads the buffer.  The learner side (PPO / MAPPO) is out of scope; this is
the adapter a trainer plugs its ``get_actions`` into.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .envs import BatchedEnv


class DeviceRolloutBuffer:
    """Fields of the reference's ``SharedReplayBuffer`` (algorithms/utils/buffer.py:270-350) as device tensors.
    ``obs[t]`` is o_t, ``actions[t]`` a_t, ``rewards[t]`` r_t, ``masks[t + 1] = 1 - env_done_t``,
    ``active_masks[t + 1] = 1 - agent_done_t`` (1 again once the whole env is done, i.e. reset)."""

    def __init__(self, T: int, n_envs: int, n_agents: int, obs_dim: int, share_dim: int, act_dim: int, hidden: int = 128,
                 layers: int = 1, device="cuda", dtype=torch.float32):
        z = lambda *s, dt=dtype: torch.zeros(s, dtype=dt, device=device)      # noqa: E731
        self.T, self.step = T, 0
        self.obs, self.share_obs = z(T + 1, n_envs, n_agents, obs_dim), z(T + 1, n_envs, n_agents, share_dim)
        self.actions, self.rewards = z(T, n_envs, n_agents, act_dim), z(T, n_envs, n_agents, 1)
        self.masks, self.active_masks = z(T + 1, n_envs, n_agents, 1) + 1, z(T + 1, n_envs, n_agents, 1) + 1
        self.action_log_probs, self.value_preds = z(T, n_envs, n_agents, 1), z(T + 1, n_envs, n_agents, 1)
        self.rnn_states_actor = z(T + 1, n_envs, n_agents, layers, hidden)
        self.rnn_states_critic = z(T + 1, n_envs, n_agents, layers, hidden)

    def insert(self, obs, share_obs, actions, rewards, dones, action_log_probs, values, rnn_states_actor, rnn_states_critic):
        """``ShareJSBSimRunner.insert`` + ``SharedReplayBuffer.insert`` (reference runner/share_jsbsim_runner.py:198-223,
        algorithms/utils/buffer.py:312-350) without leaving the device.  dones: bool/uint8 [N, A]."""
        t = self.step
        dones = dones.bool()
        dones_env = dones.all(dim=-1)                                             # [N]
        keep = (~dones_env).to(self.masks.dtype).view(-1, 1, 1)
        self.obs[t + 1].copy_(obs)
        self.share_obs[t + 1].copy_(share_obs)
        self.actions[t].copy_(actions)
        self.rewards[t].copy_(rewards.view(rewards.shape[0], rewards.shape[1], 1))
        self.masks[t + 1] = keep.expand_as(self.masks[t + 1])
        active = (~dones).to(self.masks.dtype).unsqueeze(-1)
        self.active_masks[t + 1] = torch.where(dones_env.view(-1, 1, 1), torch.ones_like(active), active)
        self.action_log_probs[t].copy_(action_log_probs)
        self.value_preds[t].copy_(values)
        self.rnn_states_actor[t + 1] = rnn_states_actor * keep.unsqueeze(-1)      # recurrent state of finished envs restarts at 0
        self.rnn_states_critic[t + 1] = rnn_states_critic * keep.unsqueeze(-1)
        self.step = (t + 1) % self.T

    def after_update(self):
        for x in (self.obs, self.share_obs, self.masks, self.active_masks, self.rnn_states_actor, self.rnn_states_critic):
            x[0].copy_(x[-1])


class DeviceRollout:
    """``policy(share_obs, obs, rnn_actor, rnn_critic, masks) -> (values, actions, log_probs, rnn_actor, rnn_critic)`` on flat
    ``[N * A, ...]`` device tensors -- the signature of the reference's ``policy.get_actions`` (runner/share_jsbsim_runner.py:159-165)."""

    def __init__(self, env: BatchedEnv, T: int, hidden: int = 128, layers: int = 1):
        self.env = env
        sp = env.spec
        share = sp.n_agents * sp.obs_dim if sp.share_obs else sp.obs_dim
        self.buffer = DeviceRolloutBuffer(T, env.n_envs, env.n_agents, sp.obs_dim, share, env.act_dim, hidden, layers, device=env.device)

    def warmup(self):
        """``runner.warmup`` (reference runner/share_jsbsim_runner.py:140-154): reset, first observation into slot 0."""
        obs, share = self.env.reset()
        self.buffer.step = 0
        self.buffer.obs[0].copy_(obs)
        self.buffer.share_obs[0].copy_(share if share is not None else obs)

    @torch.no_grad()
    def run(self, policy: Callable, n_steps: Optional[int] = None):
        """collect -> step -> insert for ``n_steps`` (default: one buffer), all on the env's stream; returns agent-steps done."""
        b, env = self.buffer, self.env
        N, A = env.n_envs, env.n_agents
        for _ in range(n_steps or b.T):
            t = b.step
            flat = lambda x: x.reshape(N * A, *x.shape[2:])                        # noqa: E731  np.concatenate(buffer.x[step])
            values, actions, logp, ha, hc = policy(flat(b.share_obs[t]), flat(b.obs[t]), flat(b.rnn_states_actor[t]),
                                                   flat(b.rnn_states_critic[t]), flat(b.masks[t]))
            actions = actions.view(N, A, -1)
            obs, share, rew, done, _ = env.step(actions.to(torch.int32))
            b.insert(obs, share if share is not None else obs, actions, rew, done, logp.view(N, A, 1), values.view(N, A, 1),
                     ha.view(N, A, *ha.shape[1:]), hc.view(N, A, *hc.shape[1:]))
        return (n_steps or b.T) * N * A
