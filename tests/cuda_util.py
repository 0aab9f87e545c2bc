"""Shared helpers of the GPU parity tests: build matching (VecEnv, OracleEnv) pairs and compare
EVERY buffer bit for bit."""
from __future__ import annotations

import numpy as np
import torch

from oracle.flock_oracle import OracleEnv


def make_pair(variant, E, N, k, cd, rs=(0, 100), sr=7.0, seed=1234, env_offset=0, rigid=False, **kw):
    from marl_range_flocking_b200 import VecEnv
    env = VecEnv(variant, E, N, k, cd, rigid_boundary=rigid, range_start=rs, sensor_range=sr, seed=seed,
                 env_offset=env_offset, **kw)
    okw = {k_: v_ for k_, v_ in kw.items() if k_ in ("reset_collision_distance", "act_noise_std", "periodic", "range_noise_std")}
    # tiled_mode / auto_reset / ... only exist on the CUDA side
    orc = OracleEnv(variant, E, N, k, cd, range_start=rs, sensor_range=sr, seed=seed, env_offset=env_offset,
                    rigid_boundary=rigid, nthreads=8, **okw)
    return env, orc


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def assert_same(name, got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if got.dtype == np.float32:
        # bit-exact, except that NaN payloads are not compared
        g, w = bits(got), bits(want)
        both_nan = np.isnan(got) & np.isnan(want)
        bad = (g != w) & ~both_nan
    else:
        bad = got != want
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{name}: {int(bad.sum())} / {bad.size} elements differ; first {idx.tolist()} "
                             f"got {got[tuple(idx[0])]!r} want {want[tuple(idx[0])]!r}")


def compare_all(env, orc, check_reward=True, tag=""):
    torch.cuda.synchronize()
    assert_same(tag + "x", env.x, orc.x)
    assert_same(tag + "y", env.y, orc.y)
    assert_same(tag + "h", env.headings, orc.h)
    assert_same(tag + "prev_h", env.prev_headings, orc.prev_h)
    assert_same(tag + "vx", env._vx, orc.vx)
    assert_same(tag + "vy", env._vy, orc.vy)
    assert_same(tag + "obs", env._obs, orc.obs)
    if env._nn is not None:
        assert_same(tag + "nn", env.nearest_neighbors, orc.nn)
    assert_same(tag + "agent_done", env.dones[0].to(torch.uint8), orc.agent_done)
    assert_same(tag + "env_done", env.dones[1].to(torch.uint8), orc.env_done)
    if check_reward:
        assert_same(tag + "reward", env.reward[..., 0], orc.reward)
    assert_same(tag + "ep_len", env._ep_len, orc.ep_len)
    assert_same(tag + "ep_return_fx", env._ep_return_fx, orc.ep_return_fx)
    assert_same(tag + "reset_epoch", env._reset_epoch.cpu().numpy().view(np.uint32), orc.reset_epoch)
    assert_same(tag + "stats", env._stats.cpu().numpy().view(np.uint64), orc.stats)
