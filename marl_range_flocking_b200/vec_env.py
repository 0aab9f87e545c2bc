"""Device-resident batched flocking environment (`VecEnv`).

E independent instances of the reference's `MultiAgentEnv` (environments/gym_flock_v2.py,
gym_flock_uw.py, gym_flock_uw_discrete.py) live in HBM as float32 structure-of-arrays `[E][N]`
tensors owned by torch; every `step()` is ONE fused sm_100a kernel launched through the C ABI of
libflock_b200.so (include/flock_b200.h). No host synchronisation happens in `reset`/`step`:
`env_done` (the reference's python bool `dones[1]`, gym_flock_v2.py:315) stays on the device.

PyTorch is used for device memory, streams and (in `dist.py`) NCCL plumbing only.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, NamedTuple, Optional, Tuple

import torch

import contextlib

from . import _lib
from ._lib import FlockBuffers, FlockCfg, VARIANT_IDS, check

_NULL_GUARD = contextlib.nullcontext()

# reference per-variant constants
_HEADING_HI = {"v2": math.pi * 1.5, "uw": math.pi * 2, "uwd": math.pi / 1.2}   # v2:96, uw:92, uwd:133
_OBS_HIST = {"v2": 1, "uw": 4, "uwd": 1}                                         # uw: memory_size, :59


def _f32(v: float) -> float:
    return float(torch.tensor(v, dtype=torch.float32).item())


class ObsRing(NamedTuple):
    """Handle of a uw env's observation history in the ring layout: `ring` is `(E, H, N, k)`, `head` `(E,)` int32;
    row r of the reference's newest-first `(N, H, k)` window is `ring[e, (head[e] + r) % H]`. The fused actor kernel
    (`BatchedActors.forward_fused`) and the replay writer take the handle as it is; anything else calls `window()`."""
    ring: torch.Tensor
    head: torch.Tensor

    def window(self) -> torch.Tensor:
        """The materialised `(E, N, H, k)` window (plain torch gather; `VecEnv.observation` uses the library kernel)."""
        E, H, N, k = self.ring.shape
        slot = (self.head.long()[:, None] + torch.arange(H, device=self.ring.device)[None, :]) % H        # (E, H)
        rows = self.ring[torch.arange(E, device=self.ring.device)[:, None], slot]                        # (E, H, N, k)
        return rows.permute(0, 2, 1, 3).contiguous()


class Trajectory(NamedTuple):
    """Time-major buffers written by `VecEnv.rollout_n` (flock_rollout_n): slice t = step t of the launch."""
    obs: torch.Tensor          # (T, E, N, k) the new range row of every step
    reward: torch.Tensor       # (T, E, N, 1)
    agent_done: torch.Tensor   # (T, E, N) bool
    env_done: torch.Tensor     # (T, E) bool
    nn: Optional[torch.Tensor]  # (T, E, N, k) int32 or None


class VecEnv:
    """Batched, device-resident version of the reference `MultiAgentEnv`.

    Constructor arguments up to `desired_distance` mirror `MultiAgentEnv.__init__`
    (gym_flock_v2.py:21-32; `normalize_distance` / `desired_distance` are accepted and ignored
    exactly as the reference's live code path ignores them). Keyword-only arguments
    are the batched extensions: `seed` / `env_offset` (Philox key and global env index of env 0),
    `auto_reset` (finished envs restart in place after each step), `tiled_mode` (kernel choice for
    N > 32, see flock_set_tiled_mode), `range_noise_std` (optional N(0, std) sensing noise on the
    observed ranges; 0 = the reference's noise-free sensing), `obs_layout` (uw only: "window" keeps the
    reference's materialised newest-first `(N, 4, k)` window in HBM, shifted every step; "ring" keeps a
    4-slot ring per env and writes only the new row -- 41 instead of 125 bytes per agent-step; the
    consumers of this package read the ring in place, `observation` materialises the window on request).
    """

    def __init__(self, variant: str, num_envs: int, agents: int, k: int = 4, collision_distance: float = 3,
                 normalize_distance: bool = False, rigid_boundary: bool = False, range_start=(0, 100),
                 sensor_range: float = 7, max_linear_velocity: float = 2.5, desired_distance: float = 15, *,
                 device=None, seed: int = 0, env_offset: int = 0, auto_reset: bool = False,
                 max_reset_attempts: int = 64, reset_collision_distance: Optional[float] = None,
                 act_noise_std: Optional[float] = None, periodic: Optional[bool] = None,
                 track_velocities: bool = True, track_neighbors: bool = True, tiled_mode: int = 0,
                 range_noise_std: float = 0.0, obs_layout: str = "window"):
        if variant not in VARIANT_IDS:
            raise ValueError(f"variant must be one of {sorted(VARIANT_IDS)}, got {variant!r}")
        if normalize_distance:
            raise NotImplementedError("normalize_distance=True is never enabled by the reference (make_env passes "
                                      "False, gym_flock_v2.py:424) and is not implemented")
        if obs_layout not in ("window", "ring"):
            raise ValueError(f"obs_layout must be 'window' or 'ring', got {obs_layout!r}")
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("marl_range_flocking_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"VecEnv lives on a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.variant = variant
        self.num_envs, self.num_particles, self.k = int(num_envs), int(agents), int(k)
        self.rigid_boundary = bool(rigid_boundary)
        self.range_start = tuple(range_start)
        self.boundary = range_start[1]
        self.sensor_range = sensor_range
        self.max_linear_velocity = max_linear_velocity
        self.collision_distance = collision_distance
        self.desired_distance = desired_distance
        self.memory_size = 4
        self.obs_hist = _OBS_HIST[variant]
        self.auto_reset = bool(auto_reset)
        self.max_reset_attempts = int(max_reset_attempts)
        r0, r1 = range_start
        reset_hi = (r1 // 2) if variant == "uw" else r1                       # gym_flock_uw.py:87-89
        if reset_collision_distance is None:
            reset_collision_distance = 4.0 if variant == "uwd" else collision_distance   # uwd:145
        if act_noise_std is None:
            act_noise_std = 0.1 if variant == "uwd" else 0.0                  # uwd:333-334
        if periodic is None:
            periodic = variant == "v2"                                        # gym_flock_v2.py:76
        self.cfg = FlockCfg(VARIANT_IDS[variant], self.num_envs, self.num_particles, self.k,
                            int(self.rigid_boundary), int(bool(periodic)), self.obs_hist, int(env_offset),
                            float(r1), float(r0), float(reset_hi), _f32(_HEADING_HI[variant]), float(sensor_range),
                            float(collision_distance), float(reset_collision_distance), float(max_linear_velocity),
                            float(act_noise_std), float(range_noise_std), int(seed) & 0xFFFFFFFFFFFFFFFF)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.flock_create(ctypes.byref(self.cfg), self.device.index, ctypes.byref(handle)))
        self._h = handle
        self.tiled = self.lib.flock_path(self._h) == 1
        check(self.lib.flock_set_tiled_mode(self._h, int(tiled_mode)))
        check(self.lib.flock_set_auto_reset(self._h, int(self.auto_reset), self.max_reset_attempts))
        E, N, H = self.num_envs, self.num_particles, self.obs_hist
        f32 = dict(dtype=torch.float32, device=self.device)
        z = lambda *shape, **kw: torch.zeros(*shape, **{**f32, **kw})
        self._x, self._y, self._hd = z(E, N), z(E, N), z(E, N)    # updated in place on both kernel paths
        self._prev_h = z(E, N)
        self._vx = z(E, N) if track_velocities else None
        self._vy = z(E, N) if track_velocities else None
        # obs | reward | agent_done | env_done share one allocation so that the host-call path moves
        # the step results with a single device->host copy
        self._out_slab, (self._obs_buf, self._reward, self._agent_done, self._env_done) = self._alloc_outputs(self.device)
        # uw ring layout: the slab's obs region holds [E][H][N][k] slots; row r of the window is slot (head + r) % H
        self.obs_ring = obs_layout == "ring" and self.obs_hist > 1
        self._obs_head = z(E, dtype=torch.int32) if self.obs_ring else None
        self._obs_window = None            # materialised on request (ring layout only)
        self._nn = z(E, N, self.k, dtype=torch.int32) if track_neighbors else None
        self._reset_epoch = z(E, dtype=torch.int32)          # uint32 counter, int32 storage
        self._ep_return_fx = z(E, dtype=torch.int64)
        self._ep_len = z(E, dtype=torch.int32)
        self._stats = z(8, dtype=torch.int64)
        ptr = lambda t: None if t is None else t.data_ptr()
        bufs = FlockBuffers(
            ptr(self._x), ptr(self._y), ptr(self._hd), ptr(self._prev_h), ptr(self._vx), ptr(self._vy), ptr(self._obs_buf),
            ptr(self._obs_head), ptr(self._nn), ptr(self._reward),
            ptr(self._agent_done), ptr(self._env_done), ptr(self._reset_epoch), ptr(self._ep_return_fx),
            ptr(self._ep_len), ptr(self._stats))
        check(self.lib.flock_bind(self._h, ctypes.byref(bufs)))
        self._host = None   # pinned host mirrors for step_host
        self._host_async = None   # side stream + validated action buffers of step_host_async
        self._dev_index = self.device.index
        self._act_shape = torch.Size((E, N) if variant == "uwd" else (E, N, 2))
        if self.obs_ring:       # the same bytes, viewed as [E][H][N][k] slots
            self._ring = ObsRing(self._obs_buf.view(E, H, N, self.k), self._obs_head)
            self._obs_view = self._ring
        else:
            self._obs_view = self._obs_buf if self.obs_hist > 1 else self._obs_buf[:, :, 0, :]

    # ------------------------------------------------------------------------------------------
    def _alloc_outputs(self, device, pin: bool = False):
        E, N, H, k = self.num_envs, self.num_particles, self.obs_hist, self.k
        n_obs, n_rew = E * N * H * k * 4, E * N * 4
        total = n_obs + n_rew + E * N + E
        slab = (torch.zeros(total, dtype=torch.uint8, pin_memory=True) if pin
                else torch.zeros(total, dtype=torch.uint8, device=device))
        obs = slab[:n_obs].view(torch.float32).view(E, N, H, k)
        reward = slab[n_obs:n_obs + n_rew].view(torch.float32).view(E, N, 1)
        agent_done = slab[n_obs + n_rew:n_obs + n_rew + E * N].view(torch.bool).view(E, N)
        env_done = slab[n_obs + n_rew + E * N:].view(torch.bool).view(E)
        return slab, (obs, reward, agent_done, env_done)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self.lib.flock_destroy(h)
            except Exception:
                pass

    def close(self):
        self.__del__()

    @property
    def _obs(self) -> torch.Tensor:
        """(E, N, H, k) newest-first observation buffer. Window layout: the env's own buffer (zero copy). Ring
        layout: materialised on request by flock_obs_window into a buffer that the next request overwrites."""
        if not self.obs_ring:
            return self._obs_buf
        if self._obs_window is None:
            self._obs_window = torch.empty_like(self._obs_buf)
        with self._dev_guard():
            check(self.lib.flock_obs_window(self._h, self._obs_window.data_ptr(), self._stream()))
        return self._obs_window

    @property
    def obs_handle(self) -> ObsRing:
        """The observation ring (uw, obs_layout="ring") for consumers that read it in place."""
        if not self.obs_ring:
            raise RuntimeError("VecEnv was not built with obs_layout='ring'")
        return self._ring

    def _stream(self) -> int:
        # raw handle of torch's current stream on our device (follows stream contexts and graph capture)
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    def _dev_guard(self):
        # kernels must be launched with our device current; the common single-GPU-per-process case
        # needs no switch at all
        if torch.cuda.current_device() == self._dev_index:
            return _NULL_GUARD
        return torch.cuda.device(self.device)

    def _as_input(self, t, shape, name) -> torch.Tensor:
        # fast path: already a contiguous float32 tensor of the right shape on our device, 8-byte aligned (the
        # kernels read action / noise pairs as float2; a view with an odd element offset would fault)
        if (type(t) is torch.Tensor and t.dtype is torch.float32 and t.device == self.device and t.is_contiguous()
                and t.shape == shape and not (t.data_ptr() & 7)):
            return t
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        t = t.to(device=self.device, dtype=torch.float32)
        if t.numel() != math.prod(shape):
            raise ValueError(f"{name} has {t.numel()} elements, expected shape {tuple(shape)}")
        t = t.reshape(shape).contiguous()
        if t.data_ptr() & 7:
            t = t.clone()                 # a fresh allocation is at least 256-byte aligned
        return t

    # ---- views of the device state (zero copy; overwritten by the next step) ------------------
    @property
    def x(self) -> torch.Tensor:
        return self._x

    @property
    def y(self) -> torch.Tensor:
        return self._y

    @property
    def headings(self) -> torch.Tensor:
        return self._hd

    @property
    def positions(self) -> torch.Tensor:
        """(E, N, 2), the reference's `positions` layout (a stacked copy of the SoA state)."""
        return torch.stack((self.x, self.y), dim=-1)

    @property
    def velocities(self) -> torch.Tensor:
        """(E, N, 2) displacement of the last step = reference `velocities` (gym_flock_v2.py:349)."""
        if self._vx is None:
            raise RuntimeError("VecEnv was built with track_velocities=False")
        return torch.stack((self._vx, self._vy), dim=-1)

    @property
    def prev_headings(self) -> torch.Tensor:
        if self.variant == "uwd":   # uwd copies headings every step (gym_flock_uw_discrete.py:249-252)
            return torch.where(self._ep_len[:, None] > 0, self.headings, self._prev_h)
        return self._prev_h

    @property
    def nearest_neighbors(self) -> torch.Tensor:
        if self._nn is None:
            raise RuntimeError("VecEnv was built with track_neighbors=False")
        return self._nn

    @property
    def distances_to_nearest_neighbors(self) -> torch.Tensor:
        if self.obs_ring:    # the newest row = the head slot of every env's ring
            E = self.num_envs
            return self._ring.ring[torch.arange(E, device=self.device), self._obs_head.long()]
        return self._obs_buf[:, :, 0, :]

    @property
    def collisions(self) -> torch.Tensor:
        """(E, N, k) int64 0/1, `_computeCollisions` (gym_flock_v2.py:212-215)."""
        return (self.distances_to_nearest_neighbors < self.collision_distance).long()

    @property
    def observation(self) -> torch.Tensor:
        """(E, N, k) for v2 / uwd, (E, N, 4, k) newest first for uw (materialised on request in the ring layout)."""
        return self._obs if self.obs_hist > 1 else self._obs_buf[:, :, 0, :]

    @property
    def reward(self) -> torch.Tensor:
        return self._reward

    @property
    def dones(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._agent_done, self._env_done

    @property
    def step_index(self) -> int:
        return int(self.lib.flock_get_step_index(self._h))

    @step_index.setter
    def step_index(self, v: int):
        check(self.lib.flock_set_step_index(self._h, int(v) & 0xFFFFFFFF))

    def pairs_evaluated(self, reset: bool = False) -> int:
        """Row x neighbour pairs evaluated by the pruned large-swarm kernel so far (synchronises)."""
        return int(self.lib.flock_pairs_evaluated(self._h, int(reset)))

    @property
    def noise_counters(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(ep_len, reset_epoch): the per-env DEVICE counters to pass as `counters=` to the fused policy kernels
        (`forward_fused(..., ou_state=...)`, `sample_action_fused`), so that their exploration draws advance with the
        env even when the launch is replayed from a CUDA graph (flock_noise_counters_t)."""
        return self._ep_len, self._reset_epoch

    @property
    def launch_count(self) -> int:
        return int(self.lib.flock_launch_count(self._h))

    # ---- the environment API -----------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None, init_state: Optional[torch.Tensor] = None,
              keep_outputs: bool = False) -> torch.Tensor:
        """Batched `reset()`. `mask`: (E,) bool, only those envs; `init_state`: (3, E, N) x/y/heading
        to install instead of drawing (parity injection); `keep_outputs`: leave reward / dones of the
        step that ended the episode untouched (auto-reset semantics)."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device).reshape(self.num_envs)
            m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
            m = m.contiguous()
        ini = None if init_state is None else self._as_input(init_state, torch.Size((3, self.num_envs, self.num_particles)), "init_state")
        with self._dev_guard():
            check(self.lib.flock_reset(self._h, None if m is None else m.data_ptr(),
                                       None if ini is None else ini.data_ptr(), self.max_reset_attempts,
                                       _lib.FLOCK_RESET_KEEP_OUTPUTS if keep_outputs else 0, self._stream()))
        return self.observation

    def step(self, actions, dt: float = 0.1, noise=None):
        """Batched `step(action, dt)`: actions (E, N, 2) [v2, uw] or (E, N) float ids [uwd].

        Returns `(obs, reward (E,N,1), (agent_done (E,N) bool, env_done (E,) bool), {})`; all are
        views of env-owned device tensors. With `obs_layout="ring"` (uw) `obs` is the `ObsRing` handle, which
        the fused actor kernel and the replay writer consume in place; `observation` materialises the window."""
        a = self._as_input(actions, self._act_shape, "actions")
        nz = None if noise is None else self._as_input(noise, torch.Size((self.num_envs, self.num_particles, 2)), "noise")
        with self._dev_guard():
            rc = self.lib.flock_step(self._h, a.data_ptr(), dt, None if nz is None else nz.data_ptr(), self._stream())
            if rc:
                check(rc)
            info: Dict = {}   # auto-reset, when enabled, happens inside flock_step (flock_set_auto_reset)
        return self._obs_view, self._reward, (self._agent_done, self._env_done), info

    def step_n(self, num_steps: int, dt: float = 0.1):
        """`num_steps` steps with the canonical in-kernel random actions (one persistent launch when N <= 32)."""
        with self._dev_guard():
            check(self.lib.flock_step_n(self._h, int(num_steps), float(dt), self._stream()))
        return self.observation, self._reward, (self._agent_done, self._env_done), {}

    def alloc_trajectory(self, num_steps: int, with_neighbors: bool = False) -> Trajectory:
        """Time-major buffers for `rollout_n` (obs | reward | agent_done | env_done of every step)."""
        T, E, N, k = int(num_steps), self.num_envs, self.num_particles, self.k
        dev = self.device
        return Trajectory(torch.zeros(T, E, N, k, device=dev), torch.zeros(T, E, N, 1, device=dev),
                          torch.zeros(T, E, N, dtype=torch.bool, device=dev), torch.zeros(T, E, dtype=torch.bool, device=dev),
                          torch.zeros(T, E, N, k, dtype=torch.int32, device=dev) if with_neighbors else None)

    def rollout_bytes_per_agent_step(self, with_actions: bool = True) -> int:
        """Algorithmic HBM bytes of one agent-step inside `rollout_n`: that step's action in, its range row, reward and
        done flag out (the state never leaves the registers)."""
        aw = 4 if self.variant == "uwd" else 8
        return (aw if with_actions else 0) + 4 * self.k + 4 + 1

    def rollout_n(self, actions: Optional[torch.Tensor], out: Trajectory, dt: float = 0.1) -> Trajectory:
        """Streamed rollout (flock_rollout_n): `T = out.obs.shape[0]` steps in ONE kernel launch (N <= 32), the state in
        registers; step t reads `actions[t]` (`(T, E, N, 2)`, or `(T, E, N)` ids for uwd; None = the canonical Philox
        random actions) and writes slice t of `out`. With `auto_reset=True` finished envs restart inside the kernel as
        they would between `step()` calls. Bit-identical to T calls of `step()`; afterwards the env holds the state and
        the outputs of the last step. The buffers of a `replay.DeviceReplay` can be passed as `out` (zero copy)."""
        T = out.obs.shape[0]
        E, N, k = self.num_envs, self.num_particles, self.k
        if (out.obs.shape != (T, E, N, k) or out.reward.numel() != T * E * N or out.agent_done.shape != (T, E, N)
                or out.env_done.shape != (T, E)):
            raise ValueError("trajectory buffers do not match (T, E, N, k)")
        for t_ in (out.obs, out.reward, out.agent_done, out.env_done):
            if not t_.is_contiguous() or t_.device != self.device:
                raise ValueError("trajectory buffers must be contiguous tensors on the env's device")
        if out.obs.dtype != torch.float32 or out.reward.dtype != torch.float32 or out.agent_done.element_size() != 1:
            raise ValueError("trajectory dtypes: obs / reward float32, dones bool or uint8")
        a_ptr = None
        if actions is not None:
            a = self._as_input(actions, torch.Size((T,) + tuple(self._act_shape)), "actions")
            a_ptr = a.data_ptr()
        nn_ptr = None
        if out.nn is not None:
            if out.nn.shape != (T, E, N, k) or out.nn.dtype != torch.int32 or not out.nn.is_contiguous():
                raise ValueError("trajectory nn must be a contiguous int32 (T, E, N, k) tensor")
            nn_ptr = out.nn.data_ptr()
        with self._dev_guard():
            check(self.lib.flock_rollout_n(self._h, T, a_ptr, float(dt), out.obs.data_ptr(), out.reward.data_ptr(),
                                           out.agent_done.data_ptr(), out.env_done.data_ptr(), nn_ptr, self._stream()))
        return out

    def random_actions(self, step_offset: int = 0) -> torch.Tensor:
        """The canonical random actions `step_n` would apply `step_offset` steps from now."""
        E, N = self.num_envs, self.num_particles
        out = torch.empty((E, N) if self.variant == "uwd" else (E, N, 2), dtype=torch.float32, device=self.device)
        with self._dev_guard():
            check(self.lib.flock_random_actions(self._h, int(step_offset) & 0xFFFFFFFF, out.data_ptr(), self._stream()))
        return out

    def step_host(self, actions_cpu: torch.Tensor, dt: float = 0.1, noise_cpu: Optional[torch.Tensor] = None):
        """End-to-end host form: host actions in, host obs / reward / dones out (pinned mirrors are
        reused between calls). The call synchronises the stream."""
        E, N = self.num_envs, self.num_particles
        if self._host is None:
            slab, (o, r, ad, ed) = self._alloc_outputs(None, pin=True)
            self._host = dict(slab=slab, obs=o, reward=r, agent_done=ad, env_done=ed,
                              ptrs=(o.data_ptr(), r.data_ptr(), ad.data_ptr(), ed.data_ptr()),
                              ret=(o if self.obs_hist > 1 else o[:, :, 0, :], r, (ad, ed), {}))
        if actions_cpu.device.type != "cpu" or actions_cpu.dtype != torch.float32 or not actions_cpu.is_contiguous():
            raise ValueError("step_host wants a contiguous float32 CPU tensor")
        want = E * N * (1 if self.variant == "uwd" else 2)
        if actions_cpu.numel() != want:
            raise ValueError(f"actions has {actions_cpu.numel()} elements, expected {want}")
        hb = self._host
        with self._dev_guard():
            rc = self.lib.flock_step_host(self._h, actions_cpu.data_ptr(), dt,
                                          None if noise_cpu is None else noise_cpu.data_ptr(), *hb["ptrs"], self._stream())
            if rc:
                check(rc)
        return hb["ret"]

    def step_host_async(self, actions_cpu: torch.Tensor, dt: float = 0.1, noise_cpu: Optional[torch.Tensor] = None) -> None:
        """`step_host` without the wait: pinned host actions in, results into the pinned host mirrors by
        one packed copy, all on this env's own side stream; returns at once. `wait_host()` blocks until
        the results have landed and returns them. Two VecEnvs driven alternately overlap one batch's
        result copy with the other's action copy and step (bench.py `e2e`). Do not mix with `step()` on
        the same env without a synchronisation in between. The per-call Python work is two ctypes calls."""
        hs = self._host_async
        if hs is None:
            E, N = self.num_envs, self.num_particles
            if self._host is None:
                slab, (o, r, ad, ed) = self._alloc_outputs(None, pin=True)
                self._host = dict(slab=slab, obs=o, reward=r, agent_done=ad, env_done=ed,
                                  ptrs=(o.data_ptr(), r.data_ptr(), ad.data_ptr(), ed.data_ptr()),
                                  ret=(o if self.obs_hist > 1 else o[:, :, 0, :], r, (ad, ed), {}))
            stream = torch.cuda.Stream(device=self.device)
            stream.wait_stream(torch.cuda.current_stream(self.device))   # after reset() etc.
            hs = self._host_async = dict(stream=stream, raw=stream.cuda_stream, ok=set(),
                                         want=E * N * (1 if self.variant == "uwd" else 2))
        key = (actions_cpu.data_ptr(), actions_cpu.numel())
        if key not in hs["ok"]:          # validated once per buffer: is_pinned() alone costs more than the launch
            if (actions_cpu.device.type != "cpu" or actions_cpu.dtype != torch.float32 or not actions_cpu.is_contiguous()
                    or not actions_cpu.is_pinned()):
                raise ValueError("step_host_async wants a pinned contiguous float32 CPU tensor")
            if actions_cpu.numel() != hs["want"]:
                raise ValueError(f"actions has {actions_cpu.numel()} elements, expected {hs['want']}")
            hs["ok"].add(key)
        if torch.cuda.current_device() != self._dev_index:
            with self._dev_guard():
                rc = self.lib.flock_step_host_async(self._h, key[0], dt, None if noise_cpu is None else noise_cpu.data_ptr(),
                                                    *self._host["ptrs"], hs["raw"])
        else:
            rc = self.lib.flock_step_host_async(self._h, key[0], dt, None if noise_cpu is None else noise_cpu.data_ptr(),
                                                *self._host["ptrs"], hs["raw"])
        if rc:
            check(rc)

    def wait_host(self):
        """Block until the last `step_host_async` has delivered; returns (obs, reward, (agent_done, env_done), {})
        as pinned host tensors (reused by the next call)."""
        rc = self.lib.flock_wait_host(self._h)
        if rc:
            check(rc)
        return self._host["ret"]

    # ---- checkpoint / injection --------------------------------------------------------------
    def get_state(self) -> Dict[str, torch.Tensor]:
        return dict(x=self.x.clone(), y=self.y.clone(), headings=self.headings.clone(), prev_headings=self._prev_h.clone(),
                    obs=self._obs.clone(), reset_epoch=self._reset_epoch.clone(), ep_len=self._ep_len.clone(),
                    ep_return_fx=self._ep_return_fx.clone(), stats=self._stats.clone(),
                    step_index=torch.tensor(self.step_index, dtype=torch.int64))

    def set_state(self, state: Dict[str, torch.Tensor]) -> None:
        self._x.copy_(state["x"]); self._y.copy_(state["y"]); self._hd.copy_(state["headings"])
        if "prev_headings" in state:
            self._prev_h.copy_(state["prev_headings"])
        if "obs" in state:
            win = state["obs"].reshape(self._obs_buf.shape)                       # (E, N, H, k) newest first
            if self.obs_ring:                                                      # ring with head 0: slot r = row r
                self._ring.ring.copy_(win.permute(0, 2, 1, 3))
                self._obs_head.zero_()
            else:
                self._obs_buf.copy_(win)
        for key, buf in (("reset_epoch", self._reset_epoch), ("ep_len", self._ep_len),
                         ("ep_return_fx", self._ep_return_fx), ("stats", self._stats)):
            if key in state:
                buf.copy_(state[key])
        if "step_index" in state:
            self.step_index = int(state["step_index"])

    # ---- logging statistics ------------------------------------------------------------------
    def stats_tensor(self) -> torch.Tensor:
        """(8,) int64 device tensor of episode statistics flushed by reset (see FLOCK_STAT_*)."""
        return self._stats

    def episode_returns(self) -> torch.Tensor:
        """(E,) float64 running return of the open episodes: sum_t sum_i reward / N (main.py:44)."""
        return self._ep_return_fx.double() / 4294967296.0 / self.num_particles

    def stats(self) -> Dict[str, float]:
        s = self._stats.tolist()      # one device->host read
        n = max(s[0], 1)
        return dict(episodes=s[0], mean_episode_length=s[1] / n,
                    mean_episode_return=s[2] / 4294967296.0 / self.num_particles / n,
                    reset_attempts=s[3], reset_gave_up=s[4])
