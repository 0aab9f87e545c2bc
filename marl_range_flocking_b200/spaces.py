"""Minimal `gym.spaces` look-alikes.

The reference builds its `action_space` / `observation_space` from gym 0.21 `spaces.Box` and
`spaces.Discrete` (gym_flock_v2.py:58-60, gym_flock_uw.py:57-58, gym_flock_uw_discrete.py:98-99)
and the learners read only `.shape`, `.n`, `.low`, `.high` from them. gym is not a dependency
here, so these two classes carry exactly those attributes (plus `sample()`/`contains()`).
"""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete:
    def __init__(self, n: int):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self) -> int:
        return int(np.random.randint(self.n))

    def contains(self, x) -> bool:
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"
