"""marl_range_flocking_b200 -- B200-native batched range-only flocking environments.

Public surface:
  VecEnv                       device-resident batched env (E envs x N agents), one fused sm_100a
                               kernel per step through libflock_b200.so
  gym_flock_v2 / gym_flock_uw / gym_flock_uw_discrete
                               single-env drop-ins for the reference modules of the same names
                               (`MultiAgentEnv`, and `make_env` for v2)
The package never falls back to CPU or plain PyTorch: importing `VecEnv` works anywhere, but
constructing one without libflock_b200.so or without a CUDA device raises.
"""
from .vec_env import VecEnv  # noqa: F401
from ._lib import FlockError, load_library  # noqa: F401

__all__ = ["VecEnv", "FlockError", "load_library"]
