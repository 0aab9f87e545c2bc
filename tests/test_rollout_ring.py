"""GPU: the uw observation ring (41 B / agent-step layout), the streamed rollout kernel (flock_rollout_n) and the
sensing noise fused into the step epilogue -- each against the CPU oracle and against the plain step path, bit for bit."""
import numpy as np
import pytest
import torch

from tests.cuda_util import assert_same, compare_all, make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("E,N,k,auto", [(64, 32, 3, False), (50, 8, 3, True), (33, 12, 4, False), (20, 5, 2, True),
                                        (4, 48, 3, False), (3, 40, 4, True)])
def test_uw_ring_layout_is_the_same_env(E, N, k, auto):
    """obs_layout="ring": [E][4][N][k] slots + a per-env head, only the new row written per step. Everything the
    env exposes -- the materialised window included -- must equal the oracle (= the window-layout env) bit for bit,
    across steps, masked resets, step_n, auto-reset and the host-buffer call."""
    kw = dict(auto_reset=True, max_reset_attempts=16, reset_collision_distance=1.5) if auto else {}
    env, orc = make_pair("uw", E, N, k, 1.5 if auto else 0.5, (0, 30 if auto else 100), 7.0, seed=17, obs_layout="ring", **kw)
    assert env.obs_ring and env.obs_handle.ring.shape == (E, 4, N, k)
    env.reset()
    orc.reset(max_attempts=16 if auto else 64)
    compare_all(env, orc, tag="reset:")
    for t in range(30):
        a = orc.random_actions()
        orc.step(a, 0.1)
        if auto:
            orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=16)
        obs, *_ = env.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"step {t}:")
        assert torch.equal(obs.window(), env.observation)            # the torch gather and the library kernel agree
        assert_same("newest row", env.distances_to_nearest_neighbors, orc.obs[:, :, 0, :])
    if not auto:
        mask = orc.env_done.copy()
        mask[::2] = 1
        orc.reset(mask=mask)
        env.reset(mask=torch.from_numpy(mask).cuda().bool())
        compare_all(env, orc, tag="masked reset:")
        for n in (1, 2, 3, 6):                                       # every residue of the head rotation
            env.step_n(n, 0.1)
            for _ in range(n):
                orc.step(orc.random_actions(), 0.1)
            compare_all(env, orc, tag=f"step_n({n}):")
    a = orc.random_actions()
    orc.step(a, 0.1)
    if auto:
        orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=16)
    obs, rew, (ad, ed), _ = env.step_host(torch.from_numpy(a).pin_memory(), 0.1)       # host gets the window
    assert_same("host obs", obs, orc.obs)
    assert_same("host reward", rew[..., 0], orc.reward)
    compare_all(env, orc, tag="host:")
    # checkpoint round trip through the window form
    from marl_range_flocking_b200 import VecEnv
    twin = VecEnv("uw", E, N, k, 1.5 if auto else 0.5, range_start=(0, 30 if auto else 100), sensor_range=7.0, seed=17,
                  obs_layout="ring", **{k_: v_ for k_, v_ in kw.items()})
    twin.set_state(env.get_state())
    act = env.random_actions()
    env.step(act, 0.1)
    twin.step(act, 0.1)
    assert torch.equal(env.observation, twin.observation) and torch.equal(env.x, twin.x)


def test_ring_and_sensing_noise():
    env, orc = make_pair("uw", 24, 32, 3, 0.5, (0, 60), 9.0, seed=31, range_noise_std=0.25, obs_layout="ring")
    env.reset()
    orc.reset()
    compare_all(env, orc, tag="noisy reset:")
    for t in range(6):
        a = orc.random_actions()
        orc.step(a, 0.1)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"noisy step {t}:")


@pytest.mark.parametrize("variant,E,N,k", [("v2", 40, 10, 4), ("uw", 20, 32, 3), ("uwd", 30, 16, 4), ("v2", 33, 7, 6)])
def test_sensing_noise_is_fused_into_the_step_launch(variant, E, N, k):
    """range_noise_std > 0 on the small path: ONE launch per step (noise in the epilogue), same bits as the oracle."""
    env, orc = make_pair(variant, E, N, k, 0.5, (0, 60), 9.0, seed=31, range_noise_std=0.25)
    env.reset()
    orc.reset()
    l0 = env.launch_count
    for t in range(8):
        a = orc.random_actions()
        orc.step(a, 0.1)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"noisy step {t}:")
    assert env.launch_count - l0 == 8


def test_fused_actor_reads_the_ring_in_place():
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.policies import BatchedActors
    dev = torch.device("cuda:0")
    E, N, k = 300, 6, 3
    env = VecEnv("uw", E, N, k, 0.5, range_start=(0, 80), sensor_range=7.0, seed=3, device=dev, obs_layout="ring")
    torch.manual_seed(1)
    actors = BatchedActors(N, 4 * k, 400, 300, 2, device=dev)
    with torch.no_grad():
        actors.w3.uniform_(-0.1, 0.1)
    env.reset()
    acts = torch.empty(E, N, 2, device=dev)
    for t in range(7):                                              # the head walks through every slot
        a_ring = actors.forward_fused(env.obs_handle, out=acts).clone()
        a_win = actors.forward_fused(env.observation)
        assert torch.equal(a_ring, a_win), t                       # same inputs, same kernel arithmetic
        env.step(acts, 0.1)
    with torch.no_grad():
        ref = actors(env.obs_handle)                                # PyTorch path materialises the window itself
    assert (actors.forward_fused(env.obs_handle) - ref).abs().max().item() <= 4e-2


ROLL = [
    # variant, E, N, k, cd, world, sensor, auto_reset, obs_layout
    ("v2", 100, 10, 4, 2.5, 50, 14.0, False, "window"),
    ("v2", 100, 10, 4, 2.5, 30, 14.0, True, "window"),
    ("v2", 37, 32, 8, 0.5, 100, 30.0, False, "window"),
    ("v2", 31, 7, 3, 1.0, 12, 9.0, True, "window"),          # generic stride, dense: many in-kernel restarts
    ("uw", 64, 32, 3, 0.5, 200, 7.0, False, "window"),
    ("uw", 64, 32, 3, 0.5, 200, 7.0, False, "ring"),
    ("uw", 50, 8, 3, 2.0, 25, 7.0, True, "ring"),
    ("uw", 33, 12, 4, 1.0, 30, 7.0, True, "window"),
    ("uwd", 64, 16, 4, 0.5, 100, 7.0, False, "window"),
    ("uwd", 33, 8, 4, 2.0, 30, 7.0, True, "window"),
    ("v2", 3, 64, 4, 0.5, 200, 30.0, False, "window"),       # N > 32: the per-step fallback behind the same API
]


@pytest.mark.parametrize("case", ROLL, ids=[f"{c[0]}-E{c[1]}-N{c[2]}-k{c[3]}-{'auto' if c[7] else 'plain'}-{c[8]}" for c in ROLL])
@pytest.mark.parametrize("given_actions", [True, False], ids=["actions", "philox"])
def test_rollout_n_equals_single_steps_and_the_oracle(case, given_actions):
    """flock_rollout_n (T steps in one launch, per-step results streamed into time-major buffers, optional in-kernel
    auto-reset) against (a) the oracle loop step -> reset(done) and (b) a twin env driven by T step() calls."""
    variant, E, N, k, cd, B, sr, auto, layout = case
    T = 23
    kw = dict(auto_reset=True, max_reset_attempts=8, reset_collision_distance=cd) if auto else {}
    if variant == "uw":
        kw["obs_layout"] = layout
    env, orc = make_pair(variant, E, N, k, cd, (0, B), sr, seed=91, **kw)
    twin, _ = make_pair(variant, E, N, k, cd, (0, B), sr, seed=91, **kw)
    for e in (env, twin):
        e.reset()
    orc.reset(max_attempts=8 if auto else 64)
    env.step_n(2, 0.1)                     # start mid-episode, ring head off zero
    twin.step_n(2, 0.1)
    for _ in range(2):
        orc.step(orc.random_actions(), 0.1)
    traj = env.alloc_trajectory(T, with_neighbors=True)
    want = dict(obs=[], reward=[], agent_done=[], env_done=[], nn=[], actions=[])
    rng = np.random.default_rng(5)
    for t in range(T):
        if given_actions:
            a = (rng.integers(0, 10, (E, N)).astype(np.float32) if variant == "uwd"
                 else rng.uniform(-1.5, 1.5, (E, N, 2)).astype(np.float32))
        else:
            a = orc.random_actions()
        want["actions"].append(a)
        orc.step(a, 0.2)
        want["reward"].append(orc.reward.copy())
        want["agent_done"].append(orc.agent_done.copy())
        want["env_done"].append(orc.env_done.copy())
        if auto:
            orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=8)
        want["obs"].append(orc.obs[:, :, 0, :].copy())
        want["nn"].append(orc.nn.copy())
        twin.step(torch.from_numpy(a).cuda(), 0.2)
    acts = torch.from_numpy(np.stack(want["actions"])).cuda() if given_actions else None
    l0 = env.launch_count
    env.rollout_n(acts, traj, 0.2)
    if N <= 32:
        assert env.launch_count - l0 == 1
    torch.cuda.synchronize()
    assert_same("traj obs", traj.obs, np.stack(want["obs"]))
    assert_same("traj reward", traj.reward[..., 0], np.stack(want["reward"]))
    assert_same("traj agent_done", traj.agent_done.to(torch.uint8), np.stack(want["agent_done"]))
    assert_same("traj env_done", traj.env_done.to(torch.uint8), np.stack(want["env_done"]))
    assert_same("traj nn", traj.nn, np.stack(want["nn"]))
    compare_all(env, orc, tag="after rollout:")
    for name in ("x", "y", "headings", "_prev_h", "_reward", "_agent_done", "_env_done", "_ep_len", "_ep_return_fx",
                 "_reset_epoch", "_stats"):
        assert torch.equal(getattr(env, name), getattr(twin, name)), name
    assert torch.equal(env.observation, twin.observation)
    if auto:
        assert int(np.stack(want["env_done"]).sum()) > 0 and env.stats()["episodes"] == int(orc.stats[0]) > 0
    # and the env carries on with plain steps
    a = orc.random_actions()
    orc.step(a, 0.1)
    if auto:
        orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=8)
    env.step(torch.from_numpy(a).cuda(), 0.1)
    compare_all(env, orc, tag="step after rollout:")


def test_rollout_n_full_size_properties():
    """BASELINE config 2 size, 128 steps per launch: sizes the oracle loop is too slow for are checked through the
    size-independent property rollout(T) == T x step on a twin env (itself oracle-checked at full size elsewhere)."""
    from marl_range_flocking_b200 import VecEnv
    mk = lambda: VecEnv("v2", 4096, 10, 4, 2.5, range_start=(0, 50), sensor_range=14.0, seed=0x5EED, auto_reset=True,
                        track_velocities=False)
    a, b = mk(), mk()
    a.reset()
    b.reset()
    T = 128
    traj = a.alloc_trajectory(T)
    a.rollout_n(None, traj, 0.1)
    rew, done = [], []
    for t in range(T):
        b.step(b.random_actions(), 0.1)
        rew.append(b.reward.clone())
        done.append(b.dones[1].clone())
    torch.cuda.synchronize()
    assert torch.equal(traj.reward, torch.stack(rew)) and torch.equal(traj.env_done, torch.stack(done))
    assert torch.equal(a.x, b.x) and torch.equal(a.observation, b.observation) and torch.equal(a._stats, b._stats)
    assert torch.equal(traj.obs[-1], a.observation)
    d = traj.obs
    assert bool((d[..., 1:] >= d[..., :-1]).all()) and bool((d >= 0).all()) and bool((d <= 14.0).all())


def test_rollout_n_writes_straight_into_the_replay_and_the_uw_windows_come_back():
    """flock_rollout_n -> TrajectoryReplay, zero copy: the kernel's time-major range rows ARE the replay storage; the
    (4, k) uw windows a learner samples are rebuilt from the row stream and must equal what a window-layout twin env
    showed step by step (auto-reset restarts included)."""
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.replay import TrajectoryReplay
    E, N, k, T = 48, 8, 3, 20
    mk = lambda layout: VecEnv("uw", E, N, k, 2.0, range_start=(0, 25), sensor_range=7.0, seed=12, auto_reset=True,
                               max_reset_attempts=8, reset_collision_distance=2.0, obs_layout=layout)
    env, twin = mk("ring"), mk("window")
    env.reset()
    twin.reset()
    rp = TrajectoryReplay(E, N, k, 2, capacity_steps=2 * T, device="cuda", chunk_size=5, window=4)
    before, after, term = [], [], []
    gen = torch.Generator(device="cuda").manual_seed(0)
    for chunk in range(2):
        acts = torch.rand(T, E, N, 2, device="cuda", generator=gen) * 2 - 1
        rp.begin(env.distances_to_nearest_neighbors)
        views = rp.next_views(T)
        assert views.obs.data_ptr() == rp.rows[chunk * T + 1].data_ptr()          # zero copy: views of the storage
        env.rollout_n(acts, views, 0.1)
        rp.commit(T, acts)
        for t in range(T):
            before.append(twin.observation.clone())
            twin.step(acts[t], 0.1)
            after.append(twin.observation.clone())
            term.append(twin.dones[1].clone())
    torch.cuda.synchronize()
    assert int(torch.stack(term).sum()) > 0                                       # episodes did end inside the rollouts
    starts = torch.tensor([0, 7, 16, 23, 35]).cuda()
    envs = torch.tensor([0, 5, 17, 30, 47]).cuda()
    s, a, r, s2, d = rp.sample_chunk(5, 5, starts=starts, envs=envs)
    for b in range(5):
        for c in range(5):
            t, e = int(starts[b]) + c, int(envs[b])
            assert torch.equal(s[b, c], before[t][e].reshape(N, 4 * k)), (b, c)
            assert torch.equal(s2[b, c], after[t][e].reshape(N, 4 * k)), (b, c)
            assert float(d[b, c, 0]) == float(term[t][e])
    assert torch.equal(env.observation, twin.observation)
