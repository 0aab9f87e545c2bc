"""Device-resident replay storage for batched rollouts ("next" scope, PyTorch).

The reference's `ReplayBufferMaddpg` (learners/maddpg_official_rnn/memory_rnn.py:8-103) and `ReplayBufferVDN`
(learners/vdn/utils.py:7-69) store one transition per call and sample chunks of consecutive steps with Python
loops. With E envs stepping together, one step is one time slice: storage is time-major `(T, E, ...)`, a slice is
written with one `copy_` per field (or, with `TrajectoryReplay`, directly by the rollout kernel) and a minibatch
of chunks is one advanced-indexing gather. Both reference return layouts are offered:

  get_minibatch  -> the seven tensors of memory_rnn.py:95-101 (MADDPG-RNN)
  sample_chunk   -> the five tensors of vdn/utils.py:54-58 (VDN)

With E = 1 both classes hold exactly what the reference buffers hold and, given the same chunk starts, return
exactly what they return (tests/test_reference_pinning.py feeds both sides the same transitions).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


def _chunk_index(starts: torch.Tensor, envs: torch.Tensor, chunk: int, capacity: int) -> Tuple[torch.Tensor, torch.Tensor]:
    tt = (starts[:, None] + torch.arange(chunk, device=starts.device)[None, :]) % capacity          # (B, C)
    return tt, envs[:, None].expand(-1, chunk)


class _ChunkSampler:
    """Chunk-start bookkeeping shared by the two buffers: logical time 0 = oldest stored slice, so a chunk never
    straddles the write head of the ring."""
    capacity: int
    counter: int
    num_envs: int

    def _draw(self, batch_size: int, chunk: int, device, generator, starts, envs):
        filled = min(self.counter, self.capacity)
        if filled < chunk:
            raise ValueError(f"need at least {chunk} stored steps, have {filled}")
        lo = getattr(self, "window", 1) - 1 if self.counter > self.capacity else 0   # history older than the ring is gone
        if starts is None:
            starts = torch.randint(lo, filled - chunk + 1, (batch_size,), device=device, generator=generator)
        else:
            starts = torch.as_tensor(starts, device=device, dtype=torch.long)
            if int(starts.min()) < 0 or int(starts.max()) > filled - chunk:
                raise ValueError("chunk start outside the stored range")
        if envs is None:
            envs = torch.randint(0, self.num_envs, (starts.numel(),), device=device, generator=generator)
        else:
            envs = torch.as_tensor(envs, device=device, dtype=torch.long)
        oldest = self.counter % self.capacity if self.counter > self.capacity else 0
        return _chunk_index(starts + oldest, envs, chunk, self.capacity)


class DeviceReplay(_ChunkSampler):
    """Replay fed one time slice at a time by `rollout.collect` (sink=`add`)."""

    def __init__(self, num_envs: int, num_agents: int, obs_dim: int, act_dim: int, capacity_steps: int,
                 device=None, chunk_size: int = 10):
        E, N, T = num_envs, num_agents, int(capacity_steps)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.obs, self.next_obs = z(T, E, N, obs_dim), z(T, E, N, obs_dim)
        self.actions = z(T, E, N, act_dim)
        self.rewards, self.dones = z(T, E, N, 1), z(T, E, N, 1)
        self.episode_end = torch.zeros(T, E, dtype=torch.bool, device=device)
        self.env_done = torch.zeros(T, E, dtype=torch.bool, device=device)
        self.capacity, self.num_envs, self.num_agents, self.chunk_size = T, E, N, chunk_size
        self.counter = 0                     # time slices written so far

    def __len__(self) -> int:                # transitions stored, like buffer_counter (memory_rnn.py:41)
        return min(self.counter, self.capacity) * self.num_envs

    def add(self, tr: Dict[str, torch.Tensor]) -> None:
        """Sink for `rollout.collect`: one time slice for all envs (batched `add_record`, memory_rnn.py:53-67;
        batched `put`, vdn/utils.py:18-28)."""
        t = self.counter % self.capacity
        E, N = self.num_envs, self.num_agents
        for key in ("obs", "next_obs"):
            v = tr[key]
            if hasattr(v, "window"):          # vec_env.ObsRing handle: store the materialised window
                v = v.window()
            getattr(self, key)[t].copy_(v.reshape(E, N, -1))
        self.actions[t].copy_(tr["actions"].reshape(E, N, -1))
        self.rewards[t].copy_(tr["reward"].reshape(E, N, 1))
        self.dones[t].copy_(tr["agent_done"].reshape(E, N, 1))          # bool -> float, memory_rnn.py:65
        self.episode_end[t].copy_(tr["episode_end"])
        self.env_done[t].copy_(tr["env_done"] if "env_done" in tr else tr["episode_end"])
        self.counter += 1

    def get_minibatch(self, batch_size: int = 128, generator: Optional[torch.Generator] = None, starts=None,
                      envs=None) -> Tuple[torch.Tensor, ...]:
        """`batch_size` chunks of `chunk_size` consecutive steps of one env each. Returns (states,
        rewards, next_states, dones, actors_states, actors_next_states, actors_actions) shaped
        (B,C,N,k), (B,C,N,1), (B,C,N,k), (B,C,N,1), (N,B,C,k), (N,B,C,k), (N,B,C,A) -- memory_rnn.py:95-101.
        `starts` / `envs` fix the chunks (logical time of the first step, env index) instead of drawing them."""
        tt, ee = self._draw(batch_size, self.chunk_size, self.obs.device, generator, starts, envs)
        states, next_states = self.obs[tt, ee], self.next_obs[tt, ee]                      # (B, C, N, k)
        rewards, dones, actions = self.rewards[tt, ee], self.dones[tt, ee], self.actions[tt, ee]
        per_agent = lambda x: x.permute(2, 0, 1, 3).contiguous()                           # (N, B, C, .)
        return states, rewards, next_states, dones, per_agent(states), per_agent(next_states), per_agent(actions)

    def sample_chunk(self, batch_size: int, chunk_size: int, generator: Optional[torch.Generator] = None, starts=None,
                     envs=None) -> Tuple[torch.Tensor, ...]:
        """VDN layout (learners/vdn/utils.py:31-58): (state (B,C,N,obs), action (B,C,N), reward (B,C,N),
        new_state (B,C,N,obs), terminal (B,C,1) = the env-level done `int(done[1])` of train_flock.py:101)."""
        tt, ee = self._draw(batch_size, chunk_size, self.obs.device, generator, starts, envs)
        return (self.obs[tt, ee], self.actions[tt, ee][..., 0], self.rewards[tt, ee][..., 0], self.next_obs[tt, ee],
                self.env_done[tt, ee].float()[..., None])


class TrajectoryReplay(_ChunkSampler):
    """Replay whose storage IS the trajectory buffer of `VecEnv.rollout_n` (flock_rollout_n): the kernel writes every
    step's range row, reward and done flags straight into the next `T` time slices (zero copy); only the actions the
    policy produced are copied in by the caller. Observations are kept as ONE time-major stream of range rows,
    `rows[t + 1]` = the row observed after step t (`rows[0]` = the observation the first step started from), so
        obs[t] = rows[t],  next_obs[t] = rows[t + 1]
    for v2 / uw_discrete, and for uw the (4, k) newest-first window is rows t, t-1, t-2, t-3 with the rows that lie
    before the episode's first observation zeroed (gym_flock_uw.py:100-102, 120-123) -- a time-major history needs no
    materialised window. With auto-reset rollouts `rows[t + 1]` of a restarted env is its new first observation, i.e.
    the stream stays consistent across episode ends (the terminal observation is not kept; `env_done[t]` masks the
    bootstrap, as the learners do)."""

    def __init__(self, num_envs: int, num_agents: int, k: int, act_dim: int, capacity_steps: int, device=None,
                 chunk_size: int = 10, window: int = 1):
        E, N, T = num_envs, num_agents, int(capacity_steps)
        self.rows = torch.zeros(T + 1, E, N, k, dtype=torch.float32, device=device)
        self.actions = torch.zeros(T, E, N, act_dim, dtype=torch.float32, device=device)
        self.rewards = torch.zeros(T, E, N, 1, dtype=torch.float32, device=device)
        self.agent_done = torch.zeros(T, E, N, dtype=torch.bool, device=device)
        self.env_done = torch.zeros(T, E, dtype=torch.bool, device=device)
        self.episode_end = torch.zeros(T, E, dtype=torch.bool, device=device)   # env_done | restarts the caller reports
        self.capacity, self.num_envs, self.num_agents, self.k = T, E, N, k
        self.chunk_size, self.window = chunk_size, int(window)
        self.counter = 0

    def __len__(self) -> int:
        return min(self.counter, self.capacity) * self.num_envs

    def next_views(self, num_steps: int):
        """The `vec_env.Trajectory` of the next `num_steps` slices (views of this buffer) for `VecEnv.rollout_n`.
        Slices are contiguous in time, so a rollout must not wrap: capacity must be a multiple of the rollout length."""
        from .vec_env import Trajectory
        t = self.counter % self.capacity
        if t + num_steps > self.capacity:
            raise ValueError("rollout would wrap the ring: make capacity_steps a multiple of the rollout length")
        if t == 0 and self.counter > 0:                  # the stream wraps: slice 0 continues from the last row written
            self.rows[0].copy_(self.rows[self.capacity])
        return Trajectory(self.rows[t + 1:t + 1 + num_steps], self.rewards[t:t + num_steps], self.agent_done[t:t + num_steps],
                          self.env_done[t:t + num_steps], None)

    def begin(self, first_rows: torch.Tensor) -> None:
        """Install the observation the next rollout starts from (the env's current newest range row, (E, N, k))."""
        self.rows[self.counter % self.capacity].copy_(first_rows)

    def commit(self, num_steps: int, actions: Optional[torch.Tensor] = None, restarted: Optional[torch.Tensor] = None) -> None:
        """Account for `num_steps` slices the kernel has written; `actions` (T, E, N, A) are copied in when given.
        `restarted` (E,) bool: envs the caller restarts after this rollout for a reason the kernel does not see (time
        limit, main.py:38-40) -- their observation history ends with the last slice."""
        t = self.counter % self.capacity
        if actions is not None:
            self.actions[t:t + num_steps].copy_(actions.reshape(num_steps, self.num_envs, self.num_agents, -1))
        self.episode_end[t:t + num_steps].copy_(self.env_done[t:t + num_steps])
        if restarted is not None:
            self.episode_end[t + num_steps - 1] |= restarted
        self.counter += num_steps

    def _obs(self, tt: torch.Tensor, ee: torch.Tensor) -> torch.Tensor:
        """(B, C, N, window * k): the observation BEFORE step tt (rows[tt] and, for uw, the older rows of the episode)."""
        if self.window == 1:
            return self.rows[tt, ee]
        parts, alive = [], torch.ones_like(tt, dtype=torch.bool)
        for r in range(self.window):
            tr = (tt - r) % self.capacity            # rows[0] == rows[capacity] once the stream has wrapped
            row = self.rows[tr, ee]
            if r:
                # row r exists only if no episode ended at the steps tt-r .. tt-1 and the slice is stored at all
                stored = (tt - r >= 0) | (self.counter > self.capacity)
                alive = alive & stored & ~self.episode_end[tr, ee]
                row = row * alive[..., None, None]
            parts.append(row)
        return torch.cat(parts, dim=-1)

    def get_minibatch(self, batch_size: int = 128, generator: Optional[torch.Generator] = None, starts=None, envs=None):
        """The seven tensors of memory_rnn.py:95-101, gathered from the row stream."""
        tt, ee = self._draw(batch_size, self.chunk_size, self.rows.device, generator, starts, envs)
        states, next_states = self._obs(tt, ee), self._obs_after(tt, ee)
        rewards, dones = self.rewards[tt, ee], self.agent_done[tt, ee].float()[..., None]
        actions = self.actions[tt, ee]
        per_agent = lambda x: x.permute(2, 0, 1, 3).contiguous()
        return states, rewards, next_states, dones, per_agent(states), per_agent(next_states), per_agent(actions)

    def _obs_after(self, tt: torch.Tensor, ee: torch.Tensor) -> torch.Tensor:
        if self.window == 1:
            return self.rows[tt + 1, ee]
        # the window after step tt: newest row rows[tt + 1]; a restart at step tt clears the history
        cur = self._obs(tt, ee)
        keep = ~self.episode_end[tt, ee]
        older = cur[..., :(self.window - 1) * self.k] * keep[..., None, None]
        return torch.cat([self.rows[tt + 1, ee], older], dim=-1)

    def sample_chunk(self, batch_size: int, chunk_size: int, generator: Optional[torch.Generator] = None, starts=None,
                     envs=None):
        """VDN layout (learners/vdn/utils.py:31-58)."""
        tt, ee = self._draw(batch_size, chunk_size, self.rows.device, generator, starts, envs)
        return (self._obs(tt, ee), self.actions[tt, ee][..., 0], self.rewards[tt, ee][..., 0], self._obs_after(tt, ee),
                self.env_done[tt, ee].float()[..., None])
