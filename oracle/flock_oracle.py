"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of the C oracle (oracle/flock_oracle.c).

Only tests/, `__graft_entry__.smoke()` and the cpu_baseline / `--impl reference` legs of
bench.py may import this module; the product package never does (it fails loudly when its
CUDA library is missing instead of falling back to anything here).

`OracleEnv` mirrors the batched semantics of the product `VecEnv` (state SoA `[E][N]`, obs
`[E][N][H][k]`) so that parity tests compare like with like; the per-function citations into
the reference live in the C file.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "flock_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_OUT_DIR, "libflock_oracle.so")

VARIANTS = {"v2": 0, "uw": 1, "uwd": 2}


def build_oracle(force: bool = False) -> str:
    """Compile the C oracle with gcc (strict IEEE: no contraction, no fast-math)."""
    if not force and os.path.isfile(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC):
        return _LIB_PATH
    os.makedirs(_OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-fno-fast-math", "-fexcess-precision=standard", "-Wall", "-o", _LIB_PATH, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


class _Cfg(ctypes.Structure):
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


class _Buf(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("x", "y", "h", "prev_h", "vx", "vy", "obs", "nn", "reward", "agent_done", "env_done",
                 "reset_epoch", "ep_return_fx", "ep_len", "stats")]


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        _lib.orc_reset.restype = ctypes.c_int
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def variant_defaults(variant: str, range_start, collision_distance: float):
    """heading_hi / reset_hi / reset collision distance / periodic / obs_hist per variant.

    v2: gym_flock_v2.py:87-96 (full box, 1.5*pi), periodic step metric (:76);
    uw: gym_flock_uw.py:87-92 (half box via integer floor-div, 2*pi), 4-deep history (:59);
    uwd: gym_flock_uw_discrete.py:125-133,145 (full box, pi/1.2, reset collision distance 4).
    """
    r0, r1 = range_start
    if variant == "v2":
        return dict(heading_hi=np.float32(np.pi * 1.5), reset_hi=np.float32(r1),
                    reset_collision_distance=np.float32(collision_distance), periodic=1, obs_hist=1,
                    act_noise_std=np.float32(0.0))
    if variant == "uw":
        return dict(heading_hi=np.float32(np.pi * 2), reset_hi=np.float32(r1 // 2),
                    reset_collision_distance=np.float32(collision_distance), periodic=0, obs_hist=4,
                    act_noise_std=np.float32(0.0))
    if variant == "uwd":
        return dict(heading_hi=np.float32(np.pi / 1.2), reset_hi=np.float32(r1),
                    reset_collision_distance=np.float32(4.0), periodic=0, obs_hist=1,
                    act_noise_std=np.float32(0.1))
    raise ValueError(variant)


class OracleEnv:
    """CPU oracle for E independent flocking envs (see module docstring)."""

    def __init__(self, variant: str, num_envs: int, agents: int, k: int, collision_distance: float,
                 range_start=(0, 100), sensor_range: float = 7, max_linear_velocity: float = 2.5,
                 rigid_boundary: bool = False, seed: int = 0, env_offset: int = 0,
                 reset_collision_distance=None, act_noise_std=None, periodic=None, nthreads: int = 1,
                 range_noise_std: float = 0.0):
        assert agents >= k + 1 and 1 <= k <= 16
        d = variant_defaults(variant, range_start, collision_distance)
        if reset_collision_distance is not None:
            d["reset_collision_distance"] = np.float32(reset_collision_distance)
        if act_noise_std is not None:
            d["act_noise_std"] = np.float32(act_noise_std)
        if periodic is not None:
            d["periodic"] = int(periodic)
        self.variant, self.E, self.N, self.k, self.H = variant, num_envs, agents, k, d["obs_hist"]
        self.nthreads = nthreads
        self.cfg = _Cfg(VARIANTS[variant], num_envs, agents, k, int(rigid_boundary), d["periodic"],
                        d["obs_hist"], env_offset, float(range_start[1]), float(range_start[0]),
                        float(d["reset_hi"]), float(d["heading_hi"]), float(sensor_range),
                        float(collision_distance), float(d["reset_collision_distance"]),
                        float(max_linear_velocity), float(d["act_noise_std"]), float(range_noise_std), seed)
        E, N, H = num_envs, agents, self.H
        f32 = np.float32
        self.x = np.zeros((E, N), f32); self.y = np.zeros((E, N), f32); self.h = np.zeros((E, N), f32)
        self.prev_h = np.zeros((E, N), f32); self.vx = np.zeros((E, N), f32); self.vy = np.zeros((E, N), f32)
        self.obs = np.zeros((E, N, H, k), f32); self.nn = np.zeros((E, N, k), np.int32)
        self.reward = np.zeros((E, N), f32); self.agent_done = np.zeros((E, N), np.uint8)
        self.env_done = np.zeros((E,), np.uint8); self.reset_epoch = np.zeros((E,), np.uint32)
        self.ep_return_fx = np.zeros((E,), np.int64); self.ep_len = np.zeros((E,), np.int32)
        self.stats = np.zeros((8,), np.uint64)
        self.step_index = 0
        self._buf = _Buf(*[_ptr(a).value for a in (self.x, self.y, self.h, self.prev_h, self.vx, self.vy,
                                                    self.obs, self.nn, self.reward, self.agent_done,
                                                    self.env_done, self.reset_epoch, self.ep_return_fx,
                                                    self.ep_len, self.stats)])
        self._lib = _load()

    # -- state injection ------------------------------------------------------------------
    def set_state(self, x, y, h, prev_h=None, obs=None):
        self.x[...] = x; self.y[...] = y; self.h[...] = h
        if prev_h is not None:
            self.prev_h[...] = prev_h
        if obs is not None:
            self.obs[...] = np.asarray(obs, np.float32).reshape(self.obs.shape)

    def reset(self, mask=None, init=None, max_attempts: int = 64, keep_outputs: bool = False) -> int:
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        ini = None if init is None else np.ascontiguousarray(init, np.float32)
        if ini is not None:
            assert ini.shape == (3, self.E, self.N)
        rc = self._lib.orc_reset(ctypes.byref(self.cfg), ctypes.byref(self._buf),
                                 None if m is None else _ptr(m), None if ini is None else _ptr(ini),
                                 ctypes.c_int(max_attempts), ctypes.c_int(int(keep_outputs)),
                                 ctypes.c_int(self.nthreads))
        self._lib.orc_range_noise(ctypes.byref(self.cfg), ctypes.byref(self._buf), None if m is None else _ptr(m))
        return rc

    def step(self, actions, dt: float = 0.1, noise=None):
        aw = 1 if self.variant == "uwd" else 2
        a = np.ascontiguousarray(actions, np.float32)
        assert a.size == self.E * self.N * aw
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        self._lib.orc_step(ctypes.byref(self.cfg), ctypes.byref(self._buf), _ptr(a),
                           None if nz is None else _ptr(nz), ctypes.c_float(dt), ctypes.c_int(self.nthreads))
        self.step_index += 1
        self._lib.orc_range_noise(ctypes.byref(self.cfg), ctypes.byref(self._buf), None)

    def random_actions(self, step_offset: int = 0) -> np.ndarray:
        aw = 1 if self.variant == "uwd" else 2
        a = np.zeros((self.E, self.N, aw) if aw == 2 else (self.E, self.N), np.float32)
        self._lib.orc_random_actions(ctypes.byref(self.cfg), ctypes.byref(self._buf), ctypes.c_uint32(step_offset),
                                     _ptr(a))
        return a


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32).copy(); k = np.asarray(key, np.uint32).copy()
    out = np.zeros(4, np.uint32)
    _load().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def sincosf(h) -> tuple:
    h = np.ascontiguousarray(h, np.float32).ravel()
    s = np.zeros_like(h); c = np.zeros_like(h)
    _load().orc_sincosf(_ptr(h), ctypes.c_int(h.size), _ptr(s), _ptr(c))
    return s, c


def normal2(words) -> np.ndarray:
    w = np.ascontiguousarray(words, np.uint32).ravel()
    z = np.zeros(w.size, np.float32)
    _load().orc_normal2(_ptr(w), ctypes.c_int(w.size // 2), _ptr(z))
    return z


def max_threads() -> int:
    return _load().orc_max_threads()
