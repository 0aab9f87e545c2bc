"""ctypes binding of libflock_b200.so (the C ABI declared in include/flock_b200.h).

There is NO fallback: if the shared library is missing or cannot be loaded this module raises,
and every entry point raises `FlockError` on a non-zero return code.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLOCK_LIBRARY_PATH lets a developer A/B an alternative build of the same ABI (still a CUDA library)
LIB_PATH = os.environ.get("FLOCK_LIBRARY_PATH") or os.path.join(_HERE, "libflock_b200.so")

FLOCK_ABI_VERSION = 2
FLOCK_MAX_K = 8
FLOCK_MAX_AGENTS = 8192
FLOCK_RESET_KEEP_OUTPUTS = 1
VARIANT_IDS = {"v2": 0, "uw": 1, "uwd": 2}
STAT_NAMES = ("episodes", "episode_steps", "episode_return_fx", "reset_attempts", "reset_gave_up")


class FlockError(RuntimeError):
    """A libflock_b200 call failed (the message comes from flock_last_error())."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libflock_b200 error {code}: {message}")
        self.code = code


class FlockCfg(ctypes.Structure):
    """flock_cfg_t"""
    _fields_ = [
        ("variant", ctypes.c_int32), ("num_envs", ctypes.c_int32), ("num_agents", ctypes.c_int32),
        ("k", ctypes.c_int32), ("rigid_boundary", ctypes.c_int32), ("periodic", ctypes.c_int32),
        ("obs_hist", ctypes.c_int32), ("env_offset", ctypes.c_int32),
        ("boundary", ctypes.c_float), ("range_lo", ctypes.c_float), ("reset_hi", ctypes.c_float),
        ("heading_hi", ctypes.c_float), ("sensor_range", ctypes.c_float),
        ("collision_distance", ctypes.c_float), ("reset_collision_distance", ctypes.c_float),
        ("max_linear_velocity", ctypes.c_float), ("act_noise_std", ctypes.c_float),
        ("range_noise_std", ctypes.c_float), ("seed", ctypes.c_uint64),
    ]


BUFFER_FIELDS = ("x", "y", "h", "prev_h", "vx", "vy", "obs", "obs_head", "nn_idx", "reward",
                 "agent_done", "env_done", "reset_epoch", "ep_return_fx", "ep_len", "stats")


class FlockBuffers(ctypes.Structure):
    """flock_buffers_t"""
    _fields_ = [(n, ctypes.c_void_p) for n in BUFFER_FIELDS]


class NoiseCounters(ctypes.Structure):
    """flock_noise_counters_t"""
    _fields_ = [("env_step", ctypes.c_void_p), ("env_epoch", ctypes.c_void_p)]


def noise_counters(counters):
    """ctypes argument for the `counters` parameter of the fused-noise entry points: None, or a pair of device
    tensors (env_step int32 [E], env_epoch int32/uint32 [E]) such as `VecEnv.noise_counters`."""
    if counters is None:
        return None
    step, epoch = counters
    return ctypes.byref(NoiseCounters(None if step is None else step.data_ptr(), None if epoch is None else epoch.data_ptr()))


_lib = None


def load_library() -> ctypes.CDLL:
    """Load libflock_b200.so or raise (never falls back to a CPU implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m marl_range_flocking_b200.build` "
            "(or __graft_entry__.build()). marl_range_flocking_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float
    sig = {
        "flock_create": (i32, [ctypes.POINTER(FlockCfg), i32, ctypes.POINTER(vp)]),
        "flock_destroy": (None, [vp]),
        "flock_bind": (i32, [vp, ctypes.POINTER(FlockBuffers)]),
        "flock_reset": (i32, [vp, vp, vp, i32, i32, vp]),
        "flock_step": (i32, [vp, vp, f32, vp, vp]),
        "flock_step_n": (i32, [vp, i32, f32, vp]),
        "flock_random_actions": (i32, [vp, u32, vp, vp]),
        "flock_rollout_n": (i32, [vp, i32, vp, f32, vp, vp, vp, vp, vp, vp]),
        "flock_obs_window": (i32, [vp, vp, vp]),
        "flock_step_host": (i32, [vp, vp, f32, vp, vp, vp, vp, vp, vp]),
        "flock_step_host_async": (i32, [vp, vp, f32, vp, vp, vp, vp, vp, vp]),
        "flock_wait_host": (i32, [vp]),
        "flock_get_step_index": (u32, [vp]),
        "flock_set_step_index": (i32, [vp, u32]),
        "flock_launch_count": (u64, [vp]),
        "flock_path": (i32, [vp]),
        "flock_set_tiled_mode": (i32, [vp, i32]),
        "flock_pairs_evaluated": (u64, [vp, i32]),
        "flock_set_auto_reset": (i32, [vp, i32, i32]),
        "flock_actor_packed_bytes": (ctypes.c_size_t, [i32]),
        "flock_actor_pack": (i32, [i32, i32, i32, i32, i32, ctypes.POINTER(vp), vp, vp]),
        "flock_actor_forward": (i32, [vp, vp, vp, i32, i32, i32, vp]),
        "flock_actor_forward_ou": (i32, [vp, vp, vp, i32, i32, i32, vp, f32, f32, f32, f32, u64, u32, i32, vp, vp]),
        "flock_actor_forward_ring": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, f32, f32, f32, f32, u64, u32, i32, vp, vp]),
        "flock_rnn_actor_packed_bytes": (ctypes.c_size_t, [i32]),
        "flock_rnn_actor_pack": (i32, [i32, i32, i32, i32, i32, ctypes.POINTER(vp), vp, vp]),
        "flock_rnn_actor_forward": (i32, [vp, ctypes.POINTER(vp), vp, vp, vp, vp, i32, i32, i32, vp]),
        "flock_rnn_actor_forward_ou": (i32, [vp, ctypes.POINTER(vp), vp, vp, vp, vp, i32, i32, i32, vp, f32, f32, f32, f32,
                                           u64, u32, i32, vp, vp]),
        "flock_qnet_forward": (i32, [ctypes.POINTER(vp), i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, u64, u32, i32, vp, vp]),
        "flock_gru_tc_packed_bytes": (ctypes.c_size_t, [i32, i32]),
        "flock_gru_tc_pack": (i32, [i32, i32, i32, i32, ctypes.POINTER(vp), vp, vp]),
        "flock_qnet_forward_tc": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, u64, u32, i32, vp, vp]),
        "flock_rnn_actor_forward_tc": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, f32, f32, f32, f32, u64, u32, i32, vp, vp]),
        "flock_last_error": (ctypes.c_char_p, []),
        "flock_abi_version": (i32, []),
        "flock_debug_sincos": (i32, [vp, i32, vp, vp, vp]),
        "flock_debug_normal2": (i32, [vp, i32, vp, vp]),
        "flock_debug_philox": (i32, [vp, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.flock_abi_version() != FLOCK_ABI_VERSION:
        raise RuntimeError(f"libflock_b200 ABI {lib.flock_abi_version()} != expected {FLOCK_ABI_VERSION}")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise FlockError(code, load_library().flock_last_error().decode("utf-8", "replace"))
