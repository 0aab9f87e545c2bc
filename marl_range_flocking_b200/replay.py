"""Device-resident replay storage for batched rollouts ("next" scope, PyTorch).

The reference's `ReplayBufferMaddpg` (learners/maddpg_official_rnn/memory_rnn.py:8-103) stores one
transition per `add_record` call and samples chunks of consecutive steps with Python loops. With E
envs stepping together, one step is one time slice: the buffer is time-major `(T, E, ...)`, a slice
is written with one `copy_` per field (no per-env loop, nothing leaves HBM) and a minibatch of
chunks is one advanced-indexing gather. `get_minibatch` returns the reference's seven tensors in the
reference's shapes (memory_rnn.py:95-101), so a learner written against it can consume them as is.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


class DeviceReplay:
    def __init__(self, num_envs: int, num_agents: int, obs_dim: int, act_dim: int, capacity_steps: int,
                 device=None, chunk_size: int = 10):
        E, N, T = num_envs, num_agents, int(capacity_steps)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.obs, self.next_obs = z(T, E, N, obs_dim), z(T, E, N, obs_dim)
        self.actions = z(T, E, N, act_dim)
        self.rewards, self.dones = z(T, E, N, 1), z(T, E, N, 1)
        self.episode_end = torch.zeros(T, E, dtype=torch.bool, device=device)
        self.capacity, self.num_envs, self.num_agents, self.chunk_size = T, E, N, chunk_size
        self.counter = 0                     # time slices written so far

    def __len__(self) -> int:                # transitions stored, like buffer_counter (memory_rnn.py:41)
        return min(self.counter, self.capacity) * self.num_envs

    def add(self, tr: Dict[str, torch.Tensor]) -> None:
        """Sink for `rollout.collect`: one time slice for all envs (batched `add_record`, memory_rnn.py:53-67)."""
        t = self.counter % self.capacity
        E, N = self.num_envs, self.num_agents
        self.obs[t].copy_(tr["obs"].reshape(E, N, -1))
        self.next_obs[t].copy_(tr["next_obs"].reshape(E, N, -1))
        self.actions[t].copy_(tr["actions"].reshape(E, N, -1))
        self.rewards[t].copy_(tr["reward"].reshape(E, N, 1))
        self.dones[t].copy_(tr["agent_done"].reshape(E, N, 1))          # bool -> float, memory_rnn.py:65
        self.episode_end[t].copy_(tr["episode_end"])
        self.counter += 1

    def get_minibatch(self, batch_size: int = 128, generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, ...]:
        """`batch_size` chunks of `chunk_size` consecutive steps of one env each. Returns (states,
        rewards, next_states, dones, actors_states, actors_next_states, actors_actions) shaped
        (B,C,N,k), (B,C,N,1), (B,C,N,k), (B,C,N,1), (N,B,C,k), (N,B,C,k), (N,B,C,A)."""
        filled = min(self.counter, self.capacity)
        C = self.chunk_size
        if filled < C:
            raise ValueError(f"need at least {C} stored steps, have {filled}")
        dev = self.obs.device
        # chunk starts are drawn in LOGICAL time (0 = oldest stored slice), so a chunk never straddles
        # the write head of the ring
        t0 = torch.randint(0, filled - C + 1, (batch_size,), device=dev, generator=generator)
        oldest = self.counter % self.capacity if self.counter > self.capacity else 0
        t0 = t0 + oldest
        e = torch.randint(0, self.num_envs, (batch_size,), device=dev, generator=generator)
        tt = (t0[:, None] + torch.arange(C, device=dev)[None, :]) % self.capacity          # (B, C)
        ee = e[:, None].expand(-1, C)
        states, next_states = self.obs[tt, ee], self.next_obs[tt, ee]                      # (B, C, N, k)
        rewards, dones, actions = self.rewards[tt, ee], self.dones[tt, ee], self.actions[tt, ee]
        per_agent = lambda x: x.permute(2, 0, 1, 3).contiguous()                           # (N, B, C, .)
        return states, rewards, next_states, dones, per_agent(states), per_agent(next_states), per_agent(actions)
