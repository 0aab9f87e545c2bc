"""Batched rollout loop: the vectorised replacement of the reference's `main.py:24-51`.

The reference plays one env, one step at a time: `get_actions` -> `env.step` -> `reward += 1` on the
last allowed step if nobody collided (`main.py:38-40`) -> `replay_buffer.add_record` ->
`score += sum(reward) / env.num_particles` (`main.py:44`), with a host sync per step on `done[1]`.
Here E envs advance together, finished or timed-out envs restart in place on the device (masked
reset), and nothing is read back until the caller asks for the statistics.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .vec_env import VecEnv

Policy = Callable[[torch.Tensor], torch.Tensor]
Sink = Callable[[Dict[str, torch.Tensor]], None]


def random_policy(env: VecEnv) -> Policy:
    """The canonical random policy (action_space sampling, Philox stream of the env)."""
    return lambda obs: env.random_actions()


@torch.no_grad()
def collect(env: VecEnv, policy: Policy, num_steps: int, max_episode_steps: int = 250, dt: float = 0.1,
            time_limit_bonus: float = 1.0, sink: Optional[Sink] = None, reset_first: bool = True) -> Dict[str, float]:
    """Advance all envs `num_steps` steps under `policy`; returns the episode statistics.

    * `max_episode_steps` = `MAX_STEPS` of `main.py:103`: an env that reaches it without a collision
      gets `time_limit_bonus` added to every agent's reward (`main.py:38-40`) and is restarted, as is
      any env whose step ended with a collision (`done[1]`).
    * `sink`, if given, receives one dict per step with device tensors `obs`, `actions`, `reward`,
      `next_obs`, `agent_done`, `env_done` (collision, the VDN terminal flag of vdn/train_flock.py:101),
      `episode_end` (collision or time limit) -- the batched form of `replay_buffer.add_record`
      (`learners/maddpg_official_rnn/memory_rnn.py:53-67`). Tensors are views that the next step
      overwrites: a sink must copy what it keeps (e.g. one `index_copy_` into a replay tensor).
    """
    if env.auto_reset:
        raise ValueError("collect() drives the resets itself; build the VecEnv with auto_reset=False")
    obs = env.reset() if reset_first else env.observation
    if hasattr(obs, "window"):            # uw ring layout: this Python loop works on the materialised window
        obs = env.observation
    for _ in range(int(num_steps)):
        prev_obs = obs.clone() if sink is not None else None
        actions = policy(obs)
        _, reward, (agent_done, env_done), _ = env.step(actions, dt)
        obs = env.observation
        timed_out = (env._ep_len >= max_episode_steps) & ~env_done
        if time_limit_bonus:
            reward += timed_out.to(reward.dtype)[:, None, None] * time_limit_bonus
            # the episode return counter tracks the reward the learner sees
            env._ep_return_fx += timed_out.to(torch.int64) * int(round(time_limit_bonus * 4294967296.0)) * env.num_particles
        episode_end = env_done | timed_out
        if sink is not None:
            terminal_obs = obs.clone()
        env.reset(mask=episode_end, keep_outputs=True)       # obs of restarted envs = first obs of the new episode
        if sink is not None:
            sink(dict(obs=prev_obs, actions=actions, reward=reward, next_obs=terminal_obs, agent_done=agent_done,
                      env_done=env_done, episode_end=episode_end))
        obs = env.observation
    return env.stats()


class GraphedRollout:
    """`collect()` without the Python loop: `steps_per_replay` iterations of policy -> `env.step` -> time-limit
    bonus -> masked restart are captured ONCE into a CUDA graph and replayed, so a whole rollout costs one graph
    launch per `steps_per_replay` steps and no host work in between (with the fused policies of `policies.py`
    a step is two kernels plus the bookkeeping elementwise ops). Results are identical to `collect()`: the same
    kernels run in the same order on the same buffers.

    The policy must be capturable: static shapes, no host synchronisation, all work on the current stream --
    e.g. `lambda obs: actors.forward_fused(obs, out=buf)`, `env.random_actions`, or a plain PyTorch module.
    """

    def __init__(self, env: VecEnv, policy: Policy, steps_per_replay: int = 32, max_episode_steps: int = 250,
                 dt: float = 0.1, time_limit_bonus: float = 1.0, warmup_steps: int = 2):
        if env.auto_reset:
            raise ValueError("GraphedRollout drives the resets itself; build the VecEnv with auto_reset=False")
        self.env, self.policy, self.steps_per_replay = env, policy, int(steps_per_replay)
        self._args = (max_episode_steps, dt, time_limit_bonus)
        self._bonus_fx = int(round(time_limit_bonus * 4294967296.0)) * env.num_particles
        self._stream = torch.cuda.Stream(device=env.device)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._warmup_steps = int(warmup_steps)

    @torch.no_grad()
    def _one_step(self) -> None:
        env = self.env
        max_episode_steps, dt, bonus = self._args
        actions = self.policy(env.observation)
        _, reward, (_, env_done), _ = env.step(actions, dt)
        timed_out = (env._ep_len >= max_episode_steps) & ~env_done
        if bonus:
            reward += timed_out.to(reward.dtype)[:, None, None] * bonus
            env._ep_return_fx += timed_out.to(torch.int64) * self._bonus_fx
        env.reset(mask=env_done | timed_out, keep_outputs=True)

    @torch.no_grad()
    def capture(self) -> None:
        """Warm up (lazy initialisations must not happen under capture), then record the graph. The warm-up
        steps are real steps of the rollout."""
        env = self.env
        self._stream.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(self._stream):
            for _ in range(self._warmup_steps):
                self._one_step()
            self._stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                for _ in range(self.steps_per_replay):
                    self._one_step()
        torch.cuda.current_stream(env.device).wait_stream(self._stream)
        self._graph = g

    @torch.no_grad()
    def run(self, num_replays: int = 1) -> Dict[str, float]:
        """Advance `num_replays * steps_per_replay` steps; returns the episode statistics (the one host read)."""
        if self._graph is None:
            self.capture()
        env = self.env
        self._stream.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(self._stream):
            for _ in range(int(num_replays)):
                self._graph.replay()
        torch.cuda.current_stream(env.device).wait_stream(self._stream)
        return env.stats()
