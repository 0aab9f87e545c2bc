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
      `next_obs`, `agent_done`, `episode_end` -- the batched form of `replay_buffer.add_record`
      (`learners/maddpg_official_rnn/memory_rnn.py:53-67`). Tensors are views that the next step
      overwrites: a sink must copy what it keeps (e.g. one `index_copy_` into a replay tensor).
    """
    if env.auto_reset:
        raise ValueError("collect() drives the resets itself; build the VecEnv with auto_reset=False")
    obs = env.reset() if reset_first else env.observation
    for _ in range(int(num_steps)):
        prev_obs = obs.clone() if sink is not None else None
        actions = policy(obs)
        obs, reward, (agent_done, env_done), _ = env.step(actions, dt)
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
                      episode_end=episode_end))
        obs = env.observation
    return env.stats()
