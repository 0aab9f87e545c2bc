"""Drop-in for the reference's `environments/gym_flock_v2.py`: `MultiAgentEnv` + `make_env`.

Heading-controlled flock: action = (linear velocity, angular velocity) per agent, periodic
min-image ranges to the k nearest neighbours as observation, collision-only reward. The whole
`step()` of gym_flock_v2.py:71-83 is one sm_100a kernel launch; this class only shapes the
containers the MADDPG learners and `main.py` expect.
"""
from __future__ import annotations

import torch

from ._single import SingleEnvBase
from .spaces import Box


class MultiAgentEnv(SingleEnvBase):
    variant = "v2"

    def __init__(self, agents, k, collision_distance, normalize_distance=False, rigid_boundary=False,
                 range_start=(0, 100), sensor_range=7, max_linear_velocity=2.5, desired_distance=15,
                 device=None, seed=0):
        self._make(agents, k, collision_distance, normalize_distance, rigid_boundary, range_start, sensor_range,
                   max_linear_velocity, desired_distance, device=device, seed=seed)
        n = self.num_particles
        # gym_flock_v2.py:58-60
        self.action_space = [Box(low=-1.5, high=1.5, shape=(2,)) for _ in range(n)]
        self.observation_space = [Box(low=0, high=range_start[1], shape=(n, self.k)),
                                  [Box(low=0, high=range_start[1], shape=(self.k,)) for _ in range(n)]]

    def _obs(self):
        d = self.vec.observation[0]
        return {"critic": d.clone(), "actors": d.clone()}      # independent tensors (gym_flock_v2.py:127-133)

    def reset(self):
        self._reset_until_free()
        return self._obs()

    def step(self, action, dt=0.1):
        if isinstance(action, (list, tuple)):
            action = torch.stack([torch.as_tensor(a) for a in action])
        _, reward, _, _ = self.vec.step(action.reshape(1, self.num_particles, 2), dt)
        return self._obs(), reward[0].clone(), self._dones(), {}


def make_env(args) -> MultiAgentEnv:
    """gym_flock_v2.py:418-429: build the env from an argparse namespace."""
    return MultiAgentEnv(agents=args.nb_agents, k=args.k, collision_distance=args.collision_distance,
                         normalize_distance=False, range_start=args.range_start, sensor_range=args.sensor_range)
