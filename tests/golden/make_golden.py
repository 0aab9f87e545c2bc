"""Generate the golden vectors that pin oracle/flock_oracle.c to the reference.

Runs the UNMODIFIED reference environments (/root/reference/environments/*.py) on CPU under
oracle/ref_shim.py and records inputs + outputs of `reset()` / `step()`. Only runnable in the
build container (the reference tree does not travel to the GPU box); the produced
tests/golden/*.npz files are committed and are what the tests read.

    python tests/golden/make_golden.py            # regenerates every fixture

Fixtures
  traj_*   free-running trajectories: initial state after reset(), per-step actions (+ injected
           actuation noise for uw_discrete) and the reference's per-step state and outputs.
  edge_*   single steps from hand-built states (ties, coincident agents, walls, NaN/Inf actions,
           zero action, k = N-1, dt != 0.1, rigid boundary), state injected into the reference
           by attribute assignment.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import NoiseInjector, load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _obs_array(variant, obs):
    if variant == "v2":
        assert torch.equal(obs["critic"], obs["actors"])
        return obs["actors"].numpy().copy()
    return obs.numpy().copy()


def _snapshot(variant, env, obs, reward, dones):
    d = dict(
        pos=env.positions.numpy().copy(), h=env.headings.numpy().copy(),
        vel=env.velocities.numpy().copy(), obs=_obs_array(variant, obs),
        reward=reward.numpy().copy().astype(np.float32), agent_done=dones[0].numpy().copy(),
        env_done=np.bool_(dones[1]),
    )
    if variant == "v2":
        d["nn"] = env.nearest_neighbors.numpy().copy().astype(np.int32)
    d["prev_h"] = env.prev_headings.numpy().copy()
    return d


def make_env(variant, kw):
    mod = load_reference(variant)
    return mod.MultiAgentEnv(**kw)


def record_traj(name, variant, kw, T, action_fn, dt=0.1, seed=0, noise_std=0.1):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    env = make_env(variant, kw)
    obs0 = env.reset()
    N = env.num_particles
    init = dict(pos0=env.positions.numpy().copy(), h0=env.headings.numpy().copy(),
                obs0=_obs_array(variant, obs0))
    steps = []
    acts, noises = [], []
    for t in range(T):
        a = action_fn(rng, N, t)
        acts.append(a.copy())
        ta = torch.from_numpy(a)
        if variant == "uwd":
            nz = (rng.standard_normal((2, N)) * noise_std).astype(np.float32)
            noises.append(nz.T.copy())       # stored [N][2] = (n_u, n_w)
            with NoiseInjector([torch.from_numpy(nz[0].copy()), torch.from_numpy(nz[1].copy())]):
                obs, rew, dones, _ = env.step(ta, dt)
        else:
            obs, rew, dones, _ = env.step(ta, dt)
        steps.append(_snapshot(variant, env, obs, rew, dones))
    out = dict(init)
    out["actions"] = np.stack(acts)
    if noises:
        out["noise"] = np.stack(noises)
    for key in steps[0]:
        out[key] = np.stack([s[key] for s in steps])
    out["dt"] = np.float32(dt)
    out["variant"] = variant
    for k_, v_ in kw.items():
        out["cfg_" + k_] = np.asarray(v_)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "steps", T, "env_done frac", float(np.mean(out["env_done"])))


def single_step(variant, kw, pos, h, action, dt=0.1, prev_h=None, noise=None, obs_mem=None):
    """Inject a state into a fresh reference env and run ONE step."""
    env = make_env(variant, kw)
    torch.manual_seed(0)
    env.reset() if _safe_reset(env) else None
    env.positions = torch.from_numpy(np.asarray(pos, np.float32).copy())
    env.headings = torch.from_numpy(np.asarray(h, np.float32).copy())
    env.prev_headings = torch.from_numpy(
        np.zeros(len(h), np.float32) if prev_h is None else np.asarray(prev_h, np.float32).copy())
    env.velocities = torch.zeros(len(h), 2)
    if variant == "uw":
        env.observation_memory = torch.from_numpy(
            np.zeros((len(h), 4, kw["k"]), np.float32) if obs_mem is None else np.asarray(obs_mem, np.float32).copy())
    if variant == "uwd":
        env.collision_distance = env.collision_temp
    ta = torch.from_numpy(np.asarray(action, np.float32).copy())
    if variant == "uwd":
        nz = np.zeros((len(h), 2), np.float32) if noise is None else np.asarray(noise, np.float32)
        with NoiseInjector([torch.from_numpy(nz[:, 0].copy()), torch.from_numpy(nz[:, 1].copy())]):
            obs, rew, dones, _ = env.step(ta, dt)
    else:
        obs, rew, dones, _ = env.step(ta, dt)
    return _snapshot(variant, env, obs, rew, dones)


def _safe_reset(env):
    # reset() allocates prev_headings etc.; it may recurse forever for crowded configs, so the
    # attributes it would create are assigned by single_step() instead.
    return False


def record_edges():
    cases = []

    def add(label, variant, kw, pos, h, action, dt=0.1, **extra):
        snap = single_step(variant, kw, pos, h, action, dt, **extra)
        c = dict(label=label, variant=variant, kw=kw, pos_in=np.asarray(pos, np.float32),
                 h_in=np.asarray(h, np.float32), action=np.asarray(action, np.float32), dt=np.float32(dt))
        for k_, v_ in extra.items():
            if v_ is not None:
                c[k_ + "_in"] = np.asarray(v_, np.float32)
        c.update({"out_" + k_: v_ for k_, v_ in snap.items()})
        cases.append(c)

    kw_v2 = dict(agents=6, k=3, collision_distance=1.0, range_start=(0, 20), sensor_range=9)
    # tiny linear velocity (clamped to 0.005) keeps the layouts essentially intact
    still = np.zeros((6, 2), np.float32)
    # 1. interior, well separated
    pos = np.array([[2, 2], [5, 3], [9, 9], [14, 4], [17, 16], [3, 15]], np.float32)
    h = np.linspace(0.1, 4.0, 6).astype(np.float32)
    add("v2_plain", "v2", kw_v2, pos, h, np.array([[1.0, 0.3]] * 6, np.float32))
    # 2. on / beyond walls: x = B exactly, x slightly below 0 after the move, x == 0 start
    pos = np.array([[19.9999, 5], [0.0001, 7], [20.0, 10], [0.001, 12], [10, 19.99999], [10, 0.00001]], np.float32)
    h = np.array([0.0, np.pi, 0.0, np.pi, np.pi / 2, -np.pi / 2], np.float32)
    add("v2_walls", "v2", kw_v2, pos, h, np.array([[2.5, 0.0]] * 6, np.float32))
    add("v2_walls_rigid", "v2", dict(kw_v2, rigid_boundary=True), pos, h, np.array([[2.5, 0.0]] * 6, np.float32))
    # 3. NaN / Inf / out-of-range actions
    pos = np.array([[2, 2], [5, 3], [9, 9], [14, 4], [17, 16], [3, 15]], np.float32)
    act = np.array([[np.nan, 0.1], [0.5, np.nan], [np.inf, 0.2], [-np.inf, -9.0], [100.0, 100.0], [1e-9, -1e-9]], np.float32)
    add("v2_bad_actions", "v2", kw_v2, pos, np.ones(6, np.float32), act)
    # 4. dt != 0.1 (test_flock.py uses 0.05 / 0.2)
    add("v2_dt", "v2", kw_v2, pos, h, np.array([[1.5, -1.0]] * 6, np.float32), dt=0.25)
    # 5. periodic neighbours across the wall + collision
    pos = np.array([[0.3, 10], [19.8, 10.2], [10, 0.2], [10.3, 19.9], [5, 5], [15, 15]], np.float32)
    add("v2_periodic_pairs", "v2", kw_v2, pos, np.zeros(6, np.float32), still)
    # 6. k = N-1
    kw_full = dict(agents=5, k=4, collision_distance=0.5, range_start=(0, 30), sensor_range=100)
    pos = np.array([[1, 1], [4, 5], [20, 22], [28, 3], [15, 15]], np.float32)
    add("v2_k_full", "v2", kw_full, pos, np.linspace(0, 3, 5).astype(np.float32), np.array([[0.7, 0.2]] * 5, np.float32))
    # 7. large headings (range reduction)
    add("v2_big_heading", "v2", kw_v2, np.array([[2, 2], [5, 3], [9, 9], [14, 4], [17, 16], [3, 15]], np.float32),
        np.array([100.0, -250.5, 1e4, -3e4, 12345.678, 710.0], np.float32), np.array([[2.0, 1.0]] * 6, np.float32))

    kw_uw = dict(agents=6, k=3, collision_distance=1.0, range_start=(0, 20), sensor_range=7)
    pos = np.array([[9, 9], [10, 10.5], [11, 9], [10, 8], [3, 15], [10, 10]], np.float32)
    hh = np.array([0.1, 0.2, 0.3, 1.0, 2.0, 6.0], np.float32)
    act = np.array([[0.5, 0.5], [0, 0], [-1, 0.2], [1e-30, 0], [3e30, 1e30], [0.3, -0.9]], np.float32)
    add("uw_first_step", "uw", kw_uw, pos, hh, act)
    mem = np.arange(6 * 4 * 3, dtype=np.float32).reshape(6, 4, 3)
    add("uw_history_prevh", "uw", kw_uw, pos, hh, act, dt=0.05, prev_h=hh, obs_mem=mem)
    add("uw_nan_action", "uw", kw_uw, pos, hh,
        np.array([[np.nan, 1], [np.inf, 1], [1, -np.inf], [0.1, 0.1], [-0.0, 0.0], [1, 1]], np.float32))

    kw_uwd = dict(agents=6, k=3, collision_distance=1.0, range_start=(0, 20), sensor_range=7)
    pos = np.array([[9, 9], [10, 10.5], [11, 9], [10, 8], [3, 15], [16, 3]], np.float32)
    hh = np.array([0.1, 0.15, 0.2, 0.5, 0.12, 2.0], np.float32)
    nz = np.array([[0.05, 0.3], [-0.3, -0.51], [0.0, 0.01], [2.5, -0.02], [-0.19, 0.0], [0.01, 0.024]], np.float32)
    add("uwd_ids", "uwd", kw_uwd, pos, hh, np.array([0, 3, 9, 5.7, 2, 7], np.float32), noise=nz)
    add("uwd_dt", "uwd", kw_uwd, pos, hh, np.array([4, 4, 1, 1, 8, 6], np.float32), dt=0.2, noise=nz[::-1].copy())

    flat = {}
    for i, c in enumerate(cases):
        for k_, v_ in c.items():
            if k_ == "kw":
                for kk, vv in v_.items():
                    flat[f"c{i}_cfg_{kk}"] = np.asarray(vv)
            else:
                flat[f"c{i}_{k_}"] = np.asarray(v_)
    flat["num_cases"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(OUT, "edge_cases.npz"), **flat)
    print("edge_cases", len(cases))


def main(only=None):
    u15 = lambda rng, N, t: rng.uniform(-1.5, 1.5, (N, 2)).astype(np.float32)      # action_space v2:58
    # BASELINE config 1: main.py defaults (main.py:96-101)
    record_traj("traj_v2_n10_k4", "v2",
                dict(agents=10, k=4, collision_distance=2.5, range_start=(0, 50), sensor_range=14),
                1000, u15, seed=0)
    record_traj("traj_v2_n32_k8", "v2",
                dict(agents=32, k=8, collision_distance=0.5, range_start=(0, 100), sensor_range=30),
                200, u15, seed=1)
    # shared-critic trainer parameters (learners/maddpg_shared_critic/train_flock.py:36), warm-up
    # actions torch.rand in [0,1) (:92) alternating with the actor range [-1,1)
    uw_act = lambda rng, N, t: (rng.uniform(0, 1, (N, 2)) if t % 2 else rng.uniform(-1, 1, (N, 2))).astype(np.float32)
    record_traj("traj_uw_n8_k3", "uw",
                dict(agents=8, k=3, collision_distance=3, range_start=(0, 50), sensor_range=7),
                400, uw_act, seed=2)
    record_traj("traj_uw_n32_k3", "uw",
                dict(agents=32, k=3, collision_distance=0.5, range_start=(0, 200), sensor_range=7),
                200, uw_act, dt=0.05, seed=3)
    # dense, fast-moving uw world: collisions (reward -5 branch), wall wraps, dt = 0.5
    record_traj("traj_uw_n10_k3_dense", "uw",
                dict(agents=10, k=3, collision_distance=2, range_start=(0, 30), sensor_range=7),
                400, uw_act, dt=0.5, seed=6)
    # VDN trainer parameters (learners/vdn/train_flock.py:78); ids over the whole dictionary
    ids = lambda rng, N, t: rng.integers(0, 10, N).astype(np.float32)
    record_traj("traj_uwd_n8_k4", "uwd", dict(agents=8, k=4, range_start=[0, 50]), 400, ids, seed=4)
    record_traj("traj_uwd_n16_k4", "uwd",
                dict(agents=16, k=4, collision_distance=0.5, range_start=[0, 100]), 200, ids, dt=0.2, seed=5)
    # north-star horizon (1e-5 after 1000 steps) for the other two variants, at the BASELINE config 3 / 4 parameters
    record_traj("traj_uw_n32_k3_1000", "uw",
                dict(agents=32, k=3, collision_distance=0.5, range_start=(0, 200), sensor_range=7),
                1000, uw_act, seed=7)
    record_traj("traj_uwd_n16_k4_1000", "uwd",
                dict(agents=16, k=4, collision_distance=0.5, range_start=[0, 100], sensor_range=7),
                1000, ids, seed=8)
    record_edges()


if __name__ == "__main__":
    main()
