"""Shared machinery of the three single-env drop-in classes (E = 1 façade over `VecEnv`).

Each façade reproduces the exact containers the reference's learners consume:
fresh tensors every call, `dones` as the tuple `(Tensor[N] bool, python bool)` -- converting
`env_done` to a python bool is the one host synchronisation per step, as upstream
(gym_flock_v2.py:315).
"""
from __future__ import annotations

import torch

from .vec_env import VecEnv


class SingleEnvBase:
    variant = "v2"
    _MAX_RESET_ROUNDS = 64      # x max_reset_attempts draws; the reference recurses without bound

    def _make(self, agents, k, collision_distance, normalize_distance, rigid_boundary, range_start, sensor_range,
              max_linear_velocity, desired_distance, device=None, seed=0):
        self.vec = VecEnv(self.variant, 1, agents, k, collision_distance, normalize_distance, rigid_boundary,
                          range_start, sensor_range, max_linear_velocity, desired_distance, device=device, seed=seed)
        v = self.vec
        self.num_particles, self.k = v.num_particles, v.k
        self.rigid_boundary, self.boundary = v.rigid_boundary, v.boundary
        self.desired_distance, self.range_start = desired_distance, range_start
        self.sensor_range, self.max_linear_velocity = sensor_range, max_linear_velocity
        self.collision_distance = collision_distance
        self.normalize_distances = normalize_distance
        self.memory_size = 4
        self.device = v.device

    # --- reference attributes, as (N, ...) tensors -------------------------------------------
    @property
    def positions(self):
        return self.vec.positions[0]

    @property
    def headings(self):
        return self.vec.headings[0]

    @property
    def prev_headings(self):
        return self.vec.prev_headings[0]

    @property
    def velocities(self):
        return self.vec.velocities[0]

    @property
    def nearest_neighbors(self):
        return self.vec.nearest_neighbors[0].long()

    @property
    def distances_to_nearest_neighbors(self):
        return self.vec.distances_to_nearest_neighbors[0]

    @property
    def collisions(self):
        return self.vec.collisions[0]

    # --- helpers -----------------------------------------------------------------------------
    def _reset_until_free(self):
        """reset(): redraw until the start is collision free (gym_flock_v2.py:105-108)."""
        for _ in range(self._MAX_RESET_ROUNDS):
            self.vec.reset()
            if not bool(self.vec.dones[1][0].item()):
                return
        raise RuntimeError(
            f"reset(): no collision-free start after {self._MAX_RESET_ROUNDS * self.vec.max_reset_attempts} draws "
            f"(agents={self.num_particles}, range_start={self.range_start}, collision_distance="
            f"{self.collision_distance}); the reference would die with RecursionError here")

    def _dones(self):
        agent_done, env_done = self.vec.dones
        return agent_done[0].clone(), bool(env_done[0].item())

    def render(self):
        """Host-side visualisation is out of scope; kept as a no-op hook."""
        return None

    def close(self):
        return None
