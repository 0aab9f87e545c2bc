"""Drop-in for the reference's `environments/gym_flock_uw.py` ("underwater", continuous).

Action = velocity direction per agent (normalised to unit length times dt), Euclidean ranges,
observation = the last four k-NN range rows, newest first, shape (N, 4, k); reward = collision
penalty + centre-of-mass term + angular-change term (gym_flock_uw.py:206-221).
"""
from __future__ import annotations

import torch

from ._single import SingleEnvBase
from .spaces import Box


class MultiAgentEnv(SingleEnvBase):
    variant = "uw"

    def __init__(self, agents, k, collision_distance, normalize_distance=False, rigid_boundary=False,
                 range_start=(0, 100), sensor_range=7, max_linear_velocity=2.5, desired_distance=15,
                 device=None, seed=0):
        self._make(agents, k, collision_distance, normalize_distance, rigid_boundary, range_start, sensor_range,
                   max_linear_velocity, desired_distance, device=device, seed=seed)
        # gym_flock_uw.py:57-58 (the declared observation shape is stale upstream; kept verbatim)
        self.action_space = Box(low=-1, high=1, shape=(2,))
        self.observation_space = Box(low=0, high=100, shape=(self.k + 2,))

    @property
    def observation_memory(self):
        return self.vec.observation[0]

    def reset(self):
        self._reset_until_free()
        return self.vec.observation[0].clone()

    def step(self, action, dt=0.1):
        if isinstance(action, (list, tuple)):
            action = torch.stack([torch.as_tensor(a).reshape(2) for a in action])
        obs, reward, _, _ = self.vec.step(action.reshape(1, self.num_particles, 2), dt)
        return obs[0].clone(), reward[0].clone(), self._dones(), {}
