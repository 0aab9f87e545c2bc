"""Drop-in for the reference's `environments/gym_flock_uw_discrete.py` (discrete actions, VDN).

Action id -> (linear, angular) means through the 10-entry dictionary of
gym_flock_uw_discrete.py:59-75, Gaussian actuation noise sigma = 0.1 (drawn in-kernel from a
Philox4x32-10 stream, or injected), heading integrator, unit-speed motion; reward = collision
(-9) + global heading alignment (gym_flock_uw_discrete.py:260-276).
"""
from __future__ import annotations

import torch

from ._single import SingleEnvBase
from .spaces import Box, Discrete

# gym_flock_uw_discrete.py:59-75
ACTION_DICTIONARY = {0: [0.2, -1.2], 1: [0.2, -0.5], 2: [0.2, 0], 3: [0.2, 0.5], 4: [0.2, 1.2],
                     5: [0.6, -1.2], 6: [0.6, -0.5], 7: [0.6, 0], 8: [0.6, 0.5], 9: [0.6, 1.2]}


class MultiAgentEnv(SingleEnvBase):
    variant = "uwd"

    def __init__(self, agents, k=4, collision_distance=3, normalize_distance=False, rigid_boundary=False,
                 range_start=(0, 100), sensor_range=7, max_linear_velocity=2.5, desired_distance=15,
                 device=None, seed=0):
        self._make(agents, k, collision_distance, normalize_distance, rigid_boundary, range_start, sensor_range,
                   max_linear_velocity, desired_distance, device=device, seed=seed)
        self.collision_temp = collision_distance
        self.action_dictionary = dict(ACTION_DICTIONARY)
        n = self.num_particles
        # gym_flock_uw_discrete.py:98-99 (Discrete(k), not Discrete(10), as upstream)
        self.action_space = [Discrete(self.k) for _ in range(n)]
        self.observation_space = [Box(low=0, high=self.sensor_range, shape=(self.k,)) for _ in range(n)]

    def reset(self):
        self._reset_until_free()
        return self.vec.observation[0].clone()

    def step(self, action, dt=0.1, noise=None):
        """`action`: (N,) float- or int-typed ids. Ids outside the dictionary raise KeyError like the
        reference's dict lookup (gym_flock_uw_discrete.py:329); `noise` (N, 2) optionally injects the
        actuation noise instead of the in-kernel Philox draw."""
        action = torch.as_tensor(action).reshape(self.num_particles)
        ids = action.detach().to("cpu", torch.float64)
        bad = ~((ids.trunc() >= 0) & (ids.trunc() <= 9))          # NaN compares False -> bad
        if bool(bad.any()):
            raise KeyError(float(ids[bad][0]))
        nz = None if noise is None else torch.as_tensor(noise).reshape(1, self.num_particles, 2)
        obs, reward, _, _ = self.vec.step(action.reshape(1, self.num_particles), dt, noise=nz)
        return obs[0].clone(), reward[0].clone(), self._dones(), {}
