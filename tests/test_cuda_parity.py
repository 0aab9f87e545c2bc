"""GPU: the sm_100a kernels (through the C ABI / VecEnv) against the CPU oracle -- BIT-EXACT.

Integer outputs (k-NN index lists, dones, episode counters) and every float32 buffer (positions,
headings, displacements, ranges, rewards) must be identical to oracle/flock_oracle.c, which is
itself pinned to the reference by tests/test_oracle_golden.py. Tolerance: 0 ulp.
"""
import numpy as np
import pytest
import torch

from tests.cuda_util import assert_same, compare_all, make_pair

pytestmark = pytest.mark.gpu


def _lib():
    from marl_range_flocking_b200 import load_library
    return load_library()


def test_philox_known_answers_and_stream():
    lib = _lib()
    from oracle import flock_oracle as fo
    rng = np.random.default_rng(0)
    ck = rng.integers(0, 2**32, (4096, 6), dtype=np.uint64).astype(np.uint32)
    ck[0] = 0
    ck[1] = 0xFFFFFFFF
    ck[2] = [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0]
    d = torch.from_numpy(ck.view(np.int32)).cuda()
    out = torch.zeros(4096, 4, dtype=torch.int32, device="cuda")
    assert lib.flock_debug_philox(d.data_ptr(), 4096, out.data_ptr(), None) == 0
    got = out.cpu().numpy().view(np.uint32)
    # Random123 known-answer vectors for philox4x32-10
    assert got[0].tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert got[1].tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert got[2].tolist() == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    want = np.stack([fo.philox4x32_10(r[:4], r[4:]) for r in ck[:256]])
    assert np.array_equal(got[:256], want)


def test_canonical_sincos_and_normal_bit_exact():
    lib = _lib()
    from oracle import flock_oracle as fo
    rng = np.random.default_rng(1)
    h = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e4, 1e4, 100000), rng.uniform(-1e8, 1e8, 20000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 2e9, 1e-30, np.pi / 4, -np.pi / 4]]).astype(np.float32)
    d = torch.from_numpy(h).cuda()
    s = torch.empty_like(d)
    c = torch.empty_like(d)
    assert lib.flock_debug_sincos(d.data_ptr(), h.size, s.data_ptr(), c.data_ptr(), None) == 0
    ws, wc = fo.sincosf(h)
    assert_same("sin", s, ws)
    assert_same("cos", c, wc)
    w = rng.integers(0, 2**32, 400000, dtype=np.uint64).astype(np.uint32)
    w[:4] = [0, 0, 0xFFFFFFFF, 0xFFFFFFFF]
    dw = torch.from_numpy(w.view(np.int32)).cuda()
    z = torch.empty(w.size, dtype=torch.float32, device="cuda")
    assert lib.flock_debug_normal2(dw.data_ptr(), w.size // 2, z.data_ptr(), None) == 0
    assert_same("normal2", z, fo.normal2(w))


SMALL = [
    # variant, E, N, k, cd, range_start, sensor_range
    ("v2", 100, 10, 4, 2.5, (0, 50), 14.0),      # BASELINE config 1/2 parameters
    ("v2", 37, 32, 8, 0.5, (0, 100), 30.0),
    ("v2", 64, 5, 4, 1.0, (0, 20), 9.0),         # k = N-1
    ("v2", 50, 2, 1, 1.0, (0, 20), 9.0),
    ("v2", 31, 7, 3, 1.0, (0, 20), 9.0),
    ("v2", 20, 17, 5, 1.0, (0, 40), 9.0),        # one env per warp, 17 live lanes
    ("uw", 64, 32, 3, 0.5, (0, 200), 7.0),       # BASELINE config 3 parameters
    ("uw", 50, 8, 3, 3.0, (0, 50), 7.0),         # shared-critic trainer parameters
    ("uw", 33, 12, 4, 1.0, (0, 60), 7.0),        # k=4 history (generic obs path)
    ("uwd", 64, 16, 4, 0.5, (0, 100), 7.0),      # BASELINE config 4 parameters
    ("uwd", 33, 8, 4, 3.0, (0, 50), 7.0),        # VDN trainer parameters
    ("uwd", 20, 25, 6, 0.5, (0, 100), 7.0),
    ("uw", 20, 30, 3, 0.5, (0, 200), 7.0),       # stride 32 but N < 32: the generic loop, not the unrolled N == 32 kernel
    ("v2", 9, 31, 8, 0.5, (0, 100), 30.0),
]
TILED = [
    ("v2", 5, 64, 8, 0.5, (0, 200), 30.0),       # TMA path (N % 4 == 0)
    ("v2", 3, 67, 4, 0.5, (0, 200), 30.0),       # plain-load path
    ("v2", 2, 300, 8, 0.3, (0, 400), 50.0),      # 3 tiles per env, ragged last tile
    ("uw", 4, 48, 3, 0.5, (0, 300), 7.0),
    ("uwd", 3, 40, 4, 0.5, (0, 300), 7.0),
    ("uw", 2, 301, 3, 0.3, (0, 400), 7.0),       # pruned Euclidean path, ragged last tile, N not a power of two (true division)
    ("uwd", 2, 260, 4, 0.3, (0, 400), 7.0),
    ("v2", 2, 33, 3, 0.5, (0, 200), 30.0),
]


def _run_parity(variant, E, N, k, cd, rs, sr, T, rigid=False, **kw):
    env, orc = make_pair(variant, E, N, k, cd, rs, sr, rigid=rigid, **kw)
    env.reset()
    orc.reset()
    compare_all(env, orc, tag="reset:")
    for t in range(T):
        a = orc.random_actions()
        dt = 0.1 if t % 3 else 0.25
        orc.step(a, dt)
        env.step(torch.from_numpy(a).cuda(), dt)
        compare_all(env, orc, tag=f"step{t}:")
    # masked reset of the finished envs + a second full reset (closes episodes -> stats)
    mask = orc.env_done.copy()
    mask[::3] = 1
    orc.reset(mask=mask)
    env.reset(mask=torch.from_numpy(mask).cuda().bool())
    compare_all(env, orc, tag="masked reset:")
    a = orc.random_actions()
    orc.step(a, 0.1)
    env.step(torch.from_numpy(a).cuda(), 0.1)
    compare_all(env, orc, tag="after masked reset:")


@pytest.mark.parametrize("case", SMALL, ids=[f"{c[0]}-E{c[1]}-N{c[2]}-k{c[3]}" for c in SMALL])
def test_small_path_bit_exact(case):
    _run_parity(*case, T=40)


@pytest.mark.parametrize("mode", [1, 2], ids=["thread-per-row", "warp-per-row-bitonic"])
@pytest.mark.parametrize("case", TILED, ids=[f"{c[0]}-E{c[1]}-N{c[2]}-k{c[3]}" for c in TILED])
def test_tiled_path_bit_exact(case, mode):
    # T >= 40: crosses two 16-step row-order refreshes (hint_slots invalidated and rebuilt through `inv`)
    _run_parity(*case, T=40, tiled_mode=mode)


def test_rigid_boundary_bit_exact():
    _run_parity("v2", 40, 10, 4, 1.0, (0, 6), 9.0, T=60, rigid=True)
    _run_parity("v2", 2, 64, 4, 0.1, (0, 6), 9.0, T=30, rigid=True, tiled_mode=1)
    _run_parity("v2", 2, 64, 4, 0.1, (0, 6), 9.0, T=30, rigid=True, tiled_mode=2)


def test_dense_world_wraps_and_collides():
    # tiny world: constant wrapping, many collisions, clamped ranges (sensor_range < typical distance)
    _run_parity("v2", 64, 10, 4, 2.5, (0, 8), 2.0, T=80)
    _run_parity("uw", 64, 10, 3, 2.5, (0, 8), 2.0, T=80)
    _run_parity("uwd", 64, 10, 4, 2.5, (0, 8), 2.0, T=80)


@pytest.mark.parametrize("variant,N,k", [("v2", 10, 4), ("uw", 32, 3), ("uwd", 16, 4), ("v2", 32, 8), ("uw", 9, 4)])
def test_step_n_persistent_matches_step_by_step(variant, N, k):
    E, T = 50, 37
    env, orc = make_pair(variant, E, N, k, 0.5, (0, 100), 9.0, seed=77)
    env.reset()
    orc.reset()
    env.step_n(T, 0.1)
    for _ in range(T):
        orc.step(orc.random_actions(), 0.1)
    compare_all(env, orc, tag="step_n:")
    # and once more from a non-zero step index, single step
    env.step_n(1, 0.2)
    orc.step(orc.random_actions(), 0.2)
    compare_all(env, orc, tag="step_n(1):")


def test_step_n_tiled_matches():
    env, orc = make_pair("v2", 2, 64, 4, 0.5, (0, 200), 30.0, seed=5)
    env.reset()
    orc.reset()
    env.step_n(5, 0.1)
    for _ in range(5):
        orc.step(orc.random_actions(), 0.1)
    compare_all(env, orc, tag="step_n tiled:")


def test_random_actions_match_oracle():
    for variant, N, k in (("v2", 10, 4), ("uw", 32, 3), ("uwd", 16, 4)):
        env, orc = make_pair(variant, 20, N, k, 0.5, (0, 100), 9.0, seed=9, env_offset=1000)
        for si in (0, 1, 12345):
            assert_same("actions", env.random_actions(si), orc.random_actions(si))


def test_uwd_injected_noise_and_nan_inf_actions():
    env, orc = make_pair("uwd", 16, 8, 4, 1.0, (0, 50), 7.0)
    env.reset()
    orc.reset()
    rng = np.random.default_rng(3)
    for t in range(10):
        a = rng.integers(0, 10, (16, 8)).astype(np.float32) + rng.uniform(0, 0.9, (16, 8)).astype(np.float32)
        nz = (rng.standard_normal((16, 8, 2)) * 0.1).astype(np.float32)
        orc.step(a, 0.1, noise=nz)
        env.step(torch.from_numpy(a).cuda(), 0.1, noise=torch.from_numpy(nz).cuda())
        compare_all(env, orc, tag=f"uwd noise {t}:")
    env, orc = make_pair("v2", 8, 6, 3, 1.0, (0, 20), 9.0)
    env.reset()
    orc.reset()
    a = orc.random_actions()
    a[0, 0] = [np.nan, 0.1]
    a[1, 1] = [0.5, np.nan]
    a[2, 2] = [np.inf, -np.inf]
    a[3, 3] = [-np.inf, 1e30]
    for t in range(3):
        orc.step(a, 0.1)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"bad actions {t}:")
    env, orc = make_pair("uw", 8, 6, 3, 1.0, (0, 20), 9.0)
    env.reset()
    orc.reset()
    a = orc.random_actions()
    a[0, 0] = [0.0, 0.0]          # zero action -> 0/0 -> nan_to_num -> 0 (gym_flock_uw.py:298)
    a[1, 1] = [np.nan, 1.0]
    a[2, 2] = [3e30, 1e30]
    orc.step(a, 0.1)
    env.step(torch.from_numpy(a).cuda(), 0.1)
    compare_all(env, orc, tag="uw bad actions:")


def test_ties_and_coincident_agents_lower_index_wins():
    # symmetric layout: agent 0 in the middle of a square -> four exactly equal distances
    N, k = 6, 3
    x = np.array([[10, 9, 11, 10, 10, 10]], np.float32)
    y = np.array([[10, 10, 10, 9, 11, 10]], np.float32)   # agent 5 coincides with agent 0
    h = np.zeros((1, N), np.float32)
    for variant in ("v2", "uw", "uwd"):
        env, orc = make_pair(variant, 1, N, k, 0.5, (0, 20), 9.0)
        init = np.stack([x, y, h])
        orc.reset(init=init)
        env.reset(init_state=torch.from_numpy(init).cuda())
        compare_all(env, orc, tag="ties reset:")
        nn = env.nearest_neighbors[0].cpu().numpy()
        assert nn[0].tolist() == [5, 1, 2]        # d=0 (coincident) first, then the lowest indices of the tie
        assert nn[5].tolist() == [0, 1, 2]
        assert not (nn == np.arange(N)[:, None]).any()   # never itself


def test_env_sharding_is_invariant():
    """Splitting the env range over two handles (as two GPUs would) reproduces the single-handle
    results bit for bit: the Philox streams are keyed by the GLOBAL env index."""
    for variant, N, k in (("v2", 10, 4), ("uwd", 16, 4)):
        E = 64
        full, _ = make_pair(variant, E, N, k, 0.5, (0, 100), 9.0, seed=42)
        lo, _ = make_pair(variant, E // 2, N, k, 0.5, (0, 100), 9.0, seed=42, env_offset=0)
        hi, _ = make_pair(variant, E // 2, N, k, 0.5, (0, 100), 9.0, seed=42, env_offset=E // 2)
        for e in (full, lo, hi):
            e.reset()
            e.step_n(25, 0.1)
        torch.cuda.synchronize()
        for name in ("x", "y", "headings", "observation", "reward"):
            want = getattr(full, name)
            got = torch.cat([getattr(lo, name), getattr(hi, name)], dim=0)
            assert torch.equal(want, got), name


def test_step_host_matches_device_path():
    env, orc = make_pair("v2", 128, 10, 4, 2.5, (0, 50), 14.0)
    env.reset()
    orc.reset()
    for t in range(5):
        a = orc.random_actions()
        orc.step(a, 0.1)
        obs, rew, (ad, ed), _ = env.step_host(torch.from_numpy(a).pin_memory(), 0.1)
        assert_same("host obs", obs, orc.obs[:, :, 0, :])
        assert_same("host reward", rew[..., 0], orc.reward)
        assert_same("host agent_done", ad.to(torch.uint8), orc.agent_done)
        assert_same("host env_done", ed.to(torch.uint8), orc.env_done)
    compare_all(env, orc, tag="host:")


@pytest.mark.parametrize("variant,E,N,k,B", [("v2", 256, 10, 4, 30), ("uw", 128, 32, 3, 60), ("uwd", 128, 16, 4, 40),
                                             ("v2", 97, 7, 3, 15), ("v2", 6, 64, 4, 60)])
def test_auto_reset_resets_exactly_the_done_envs(variant, E, N, k, B):
    """auto_reset: step + restart of the finished envs (one fused launch for N <= 32; integrate
    pre-pass, sensing kernel and a follow-up masked multi-CTA reset for N > 32) == oracle step followed by
    reset(mask = env_done, keep_outputs)."""
    env, orc = make_pair(variant, E, N, k, 2.5, (0, B), 14.0, auto_reset=True, reset_collision_distance=2.5,
                         max_reset_attempts=16)
    env.reset()
    orc.reset(max_attempts=16)
    launches0 = env.launch_count
    T = 30
    for t in range(T):
        a = orc.random_actions()
        orc.step(a, 0.1)
        orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=16)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"auto reset {t}:")
    launches = env.launch_count - launches0
    if N <= 32:
        assert launches == T                     # fused: one kernel per step
    else:    # integrate pre-pass + sensing kernel + masked reset (two multi-CTA attempts + the loop kernel) (+ row-order refreshes)
        assert 5 * T <= launches <= 5 * T + T // 16 + 1
    s = env.stats()
    assert s["episodes"] == int(orc.stats[0]) and s["episodes"] > 0
    # host-buffer path with auto-reset: the host sees the restarted envs' first observation too
    a = orc.random_actions()
    orc.step(a, 0.1)
    orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=16)
    obs, rew, (ad, ed), _ = env.step_host(torch.from_numpy(a).pin_memory(), 0.1)
    assert_same("host obs", obs, orc.obs if env.obs_hist > 1 else orc.obs[:, :, 0, :])
    assert_same("host env_done", ed.to(torch.uint8), orc.env_done)
    compare_all(env, orc, tag="auto reset host:")


FULL = [
    # the three BASELINE batch configs, 1000 free-running steps each (the north-star horizon)
    ("v2", 4096, 10, 4, 2.5, (0, 50), 14.0, 1000),    # BASELINE config 2
    ("uw", 4096, 32, 3, 0.5, (0, 200), 7.0, 1000),    # BASELINE config 3
    ("uwd", 8192, 16, 4, 0.5, (0, 100), 7.0, 1000),   # BASELINE config 4 (in-kernel Philox actuation noise)
]


@pytest.mark.parametrize("case", FULL, ids=[f"{c[0]}-E{c[1]}-N{c[2]}" for c in FULL])
def test_full_size_configs_bit_exact(case):
    variant, E, N, k, cd, rs, sr, T = case
    env, orc = make_pair(variant, E, N, k, cd, rs, sr, seed=0x5EED,
                         **({"reset_collision_distance": 1.0} if variant == "uwd" else {}))
    env.reset()
    orc.reset()
    compare_all(env, orc, tag="reset:")
    for t in range(T):
        a = orc.random_actions()
        orc.step(a, 0.1)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        if t % 100 == 99:
            compare_all(env, orc, tag=f"step {t}:")
    compare_all(env, orc, tag="final:")
    # domain properties at full size: ranges ascending, clamped, self never a neighbour
    d = env.distances_to_nearest_neighbors
    assert bool((d[..., 1:] >= d[..., :-1]).all()) and bool((d >= 0).all()) and bool((d <= sr).all())
    nn = env.nearest_neighbors
    assert not bool((nn == torch.arange(N, device=nn.device)[None, :, None]).any())
    assert bool(((nn >= 0) & (nn < N)).all())


@pytest.mark.parametrize("auto_reset", [False, True], ids=["masked-reset", "auto-reset"])
@pytest.mark.parametrize("mode", [1, 2], ids=["thread-per-row", "warp-per-row-bitonic"])
def test_large_swarm_config5_bit_exact(mode, auto_reset):
    """BASELINE config 5 shape (E reduced to 8 to keep the CPU oracle fast): 2048 agents, k = 8, 72 steps = four
    16-step row-order refreshes of the pruned kernel (its neighbour hints are invalidated and rebuilt through the
    inverse row order each time), with a masked reset of a subset in the middle of a refresh period -- or, in the
    auto-reset variant, a world dense enough that envs finish and restart on their own."""
    cd = 0.3 if auto_reset else 0.05
    kw = dict(auto_reset=True, max_reset_attempts=16) if auto_reset else {}
    env, orc = make_pair("v2", 8, 2048, 8, cd, (0, 2000), 100.0, seed=0x5EED, tiled_mode=mode, **kw)
    env.reset()
    orc.reset(max_attempts=16 if auto_reset else 64)
    compare_all(env, orc, tag="reset:")
    T = 72
    restarted = 0
    for t in range(T):
        a = orc.random_actions()
        dt = 0.1 if t % 5 else 0.5
        orc.step(a, dt)
        if auto_reset:
            restarted += int(orc.env_done.sum())
            orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=16)
        env.step(torch.from_numpy(a).cuda(), dt)
        if t < 4 or t % 4 == 3 or t in (16, 17, 32, 33, 41, 42, 48, 49, 64, 65):
            compare_all(env, orc, tag=f"step{t}:")
        if not auto_reset and t == 40:            # mid-period masked reset: envs 1, 4, 7 restart, hints of the rest stay
            mask = np.zeros(8, np.uint8)
            mask[1::3] = 1
            orc.reset(mask=mask)
            env.reset(mask=torch.from_numpy(mask).cuda().bool())
            compare_all(env, orc, tag="masked reset:")
    compare_all(env, orc, tag="final:")
    if auto_reset:
        assert restarted > 0 and env.stats()["episodes"] == int(orc.stats[0]) > 0
    # domain properties: ranges ascending and clamped, self never a neighbour, indices in range
    d, nn = env.distances_to_nearest_neighbors, env.nearest_neighbors
    assert bool((d[..., 1:] >= d[..., :-1]).all()) and bool((d >= 0).all()) and bool((d <= 100.0).all())
    assert not bool((nn == torch.arange(2048, device=nn.device)[None, :, None]).any())
    assert bool(((nn >= 0) & (nn < 2048)).all())


@pytest.mark.parametrize("mode", [1, 2], ids=["thread-per-row-pruned", "warp-per-row-bitonic"])
@pytest.mark.parametrize("variant,k,cd", [("uw", 3, 0.05), ("uwd", 4, 0.05)])
def test_large_swarm_uw_uwd_bit_exact(variant, k, cd, mode):
    """2048-agent uw / uwd swarms (Euclidean metric, gym_flock_uw.py:125-144): the spatially pruned thread-per-row kernel
    and the warp-per-row kernel against the oracle over 40 steps = two row-order refreshes, with a masked reset in
    between. The centre of mass / mean heading behind the rewards is the canonical sequential sum, formed once per env
    by the integrate pre-pass."""
    env, orc = make_pair(variant, 4, 2048, k, cd, (0, 2000), 7.0, seed=0xFACE, tiled_mode=mode)
    env.reset()
    orc.reset(max_attempts=64)
    compare_all(env, orc, tag="reset:")
    for t in range(40):
        a = orc.random_actions()
        dt = 0.1 if t % 5 else 0.5
        orc.step(a, dt)
        env.step(torch.from_numpy(a).cuda(), dt)
        if t < 3 or t % 4 == 3 or t in (16, 17, 25, 26, 32, 33):
            compare_all(env, orc, tag=f"step{t}:")
        if t == 24:
            mask = np.array([0, 1, 0, 1], np.uint8)
            orc.reset(mask=mask)
            env.reset(mask=torch.from_numpy(mask).cuda().bool())
            compare_all(env, orc, tag="masked reset:")
    compare_all(env, orc, tag="final:")


@pytest.mark.parametrize("variant,N,k,layout", [("uw", 301, 3, "window"), ("uw", 301, 3, "ring"), ("uwd", 260, 4, "window")])
def test_dense_euclidean_swarm_with_auto_reset_bit_exact(variant, N, k, layout):
    """Large uw / uwd swarms in a small world: constant wrap-around (far-row flags, poisoned hints), collisions that end
    episodes, in-launch sequence step -> masked auto-reset (hints and flags of the restarted envs dropped), 48 steps =
    three row-order refreshes; uw on both observation layouts. Pruned thread-per-row kernel against the oracle."""
    kw = dict(obs_layout=layout) if variant == "uw" else {}
    env, orc = make_pair(variant, 3, N, k, 0.6, (0, 120), 7.0, seed=0xD0D0, tiled_mode=1, auto_reset=True, max_reset_attempts=8, **kw)
    env.reset()
    orc.reset(max_attempts=8)
    restarted = 0
    for t in range(48):
        a = orc.random_actions()
        dt = 0.1 if t % 4 else 0.4
        orc.step(a, dt)
        restarted += int(orc.env_done.sum())
        orc.reset(mask=orc.env_done.copy(), keep_outputs=True, max_attempts=8)
        env.step(torch.from_numpy(a).cuda(), dt)
        torch.cuda.synchronize()
        assert_same(f"step{t}:x", env.x, orc.x)
        assert_same(f"step{t}:reward", env.reward[..., 0], orc.reward)
        assert_same(f"step{t}:env_done", env.dones[1].to(torch.uint8), orc.env_done)
        obs = env.observation
        assert_same(f"step{t}:obs", obs, orc.obs.reshape(obs.shape))
    assert restarted > 0 and env.stats()["episodes"] == int(orc.stats[0]) > 0


def test_two_large_swarm_envs_of_different_size_share_a_device():
    """The dynamic shared-memory opt-in is per FUNCTION, not per handle: a smaller large-swarm env created after
    a bigger one must not lower the limit under it (8192 agents need > 48 KB). Both are stepped alternately and
    the smaller one is checked against the oracle."""
    from marl_range_flocking_b200 import VecEnv
    big = VecEnv("v2", 1, 8192, 8, 0.02, range_start=(0, 4000), sensor_range=100.0, seed=3)
    small, orc = make_pair("v2", 1, 4096, 8, 0.02, (0, 3000), 100.0, seed=4)
    big.reset()
    small.reset()
    orc.reset()
    for t in range(3):
        big.step(big.random_actions(), 0.1)
        a = orc.random_actions()
        orc.step(a, 0.1)
        small.step(torch.from_numpy(a).cuda(), 0.1)
        big.step(big.random_actions(), 0.1)
    torch.cuda.synchronize()
    compare_all(small, orc, tag="4096 next to 8192:")
    d = big.distances_to_nearest_neighbors
    assert bool(torch.isfinite(d).all()) and bool((d[..., 1:] >= d[..., :-1]).all())


@pytest.mark.parametrize("variant,E,N,k", [("v2", 64, 10, 4), ("uw", 32, 32, 3), ("uwd", 48, 16, 4), ("v2", 2, 96, 8),
                                           ("uw", 3, 200, 3), ("uwd", 3, 130, 4)])   # tiled Euclidean: stale hints / far-row flags after the restore
def test_get_state_set_state_round_trip_continues_bit_equal(variant, E, N, k):
    """Checkpoint / restore: a twin env that receives `get_state()` of a running env must continue
    with identical bits -- state, observation window, episode counters, Philox epochs, statistics."""
    from marl_range_flocking_b200 import VecEnv
    mk = lambda seed: VecEnv(variant, E, N, k, 0.5, range_start=(0, 60), sensor_range=9.0, seed=seed, device="cuda:0")
    a = mk(5)
    a.reset()
    a.step_n(13, 0.1)
    a.reset(mask=a.dones[1].clone())                       # some envs mid-episode, some freshly restarted
    a.step_n(3, 0.1)
    state = {k_: v_.cpu() for k_, v_ in a.get_state().items()}   # through host memory, as a checkpoint file would
    b = mk(5)                                               # same seed = same Philox key; everything else from `state`
    b.set_state(state)
    for t in range(6):
        act = a.random_actions()
        assert torch.equal(act, b.random_actions())
        a.step(act, 0.1)
        b.step(act, 0.1)
    a.step_n(5, 0.1)                                        # in-kernel Philox actions + (uwd) actuation noise
    b.step_n(5, 0.1)
    a.reset(mask=a.dones[1].clone())
    b.reset(mask=b.dones[1].clone())
    torch.cuda.synchronize()
    for name in ("x", "y", "headings", "_prev_h", "_obs", "_reward", "_agent_done", "_env_done", "_ep_len",
                 "_ep_return_fx", "_reset_epoch", "_stats"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_misaligned_action_views_are_copied_not_faulted():
    """A contiguous float32 view with an odd element offset is not 8-byte aligned; the kernels read action pairs
    as float2, so VecEnv must copy such an input (and the C ABI must refuse it) instead of faulting."""
    env, orc = make_pair("v2", 16, 10, 4, 2.5, (0, 50), 14.0)
    env.reset()
    orc.reset()
    a = orc.random_actions()
    buf = torch.zeros(1 + a.size, device="cuda")
    buf[1:] = torch.from_numpy(a).cuda().reshape(-1)
    view = buf[1:].view(16, 10, 2)
    assert view.data_ptr() % 8 == 4
    orc.step(a, 0.1)
    env.step(view, 0.1)
    compare_all(env, orc, tag="misaligned view:")
    rc = env.lib.flock_step(env._h, view.data_ptr(), 0.1, None, None)
    assert rc == -1 and b"aligned" in env.lib.flock_last_error()


def test_batched_rollout_matches_per_step_oracle_loop():
    """rollout.collect (main.py:24-51, batched) against the same loop played on the oracle."""
    from marl_range_flocking_b200.rollout import collect, random_policy
    E, N, k, T, MAXS = 200, 10, 4, 120, 25
    env, orc = make_pair("v2", E, N, k, 2.5, (0, 40), 14.0, seed=11)
    seen = []
    stats = collect(env, random_policy(env), T, max_episode_steps=MAXS,
                    sink=lambda tr: seen.append((tr["reward"].clone(), tr["episode_end"].clone())))
    orc.reset()
    bonus_fx = np.int64(1 << 32) * N
    for t in range(T):
        orc.step(orc.random_actions(), 0.1)
        timed_out = (orc.ep_len >= MAXS) & (orc.env_done == 0)
        orc.reward[timed_out] += np.float32(1.0)
        orc.ep_return_fx[timed_out] += bonus_fx
        end = (orc.env_done != 0) | timed_out
        assert_same(f"reward {t}", seen[t][0][..., 0], orc.reward)
        assert_same(f"episode_end {t}", seen[t][1].to(torch.uint8), end.astype(np.uint8))
        orc.reset(mask=end.astype(np.uint8), keep_outputs=True)
    compare_all(env, orc, tag="rollout:")
    assert stats["episodes"] == int(orc.stats[0]) > E
    assert stats["mean_episode_length"] <= MAXS


def test_single_env_facades_have_the_reference_surface():
    """Shapes / containers the learners rely on (the reference's own stale unittest
    learners/maddpg_official/test.py:49-82 only checks shapes as well)."""
    import argparse
    from marl_range_flocking_b200 import gym_flock_uw, gym_flock_uw_discrete, gym_flock_v2
    args = argparse.Namespace(nb_agents=10, k=4, collision_distance=2.5, range_start=(0, 50), sensor_range=14)
    env = gym_flock_v2.make_env(args)
    assert env.num_particles == 10 and env.k == 4 and len(env.action_space) == 10
    assert env.action_space[0].shape[0] == 2 and env.observation_space[0].shape == (10, 4)
    assert env.observation_space[1][3].shape[0] == 4
    obs = env.reset()
    assert set(obs) == {"critic", "actors"} and obs["actors"].shape == (10, 4) and obs["actors"].is_cuda
    assert obs["actors"].data_ptr() != obs["critic"].data_ptr()
    a = torch.stack([torch.rand(2, device="cuda") * 3 - 1.5 for _ in range(10)])     # MADDPG.py:32
    nobs, reward, done, info = env.step(a)
    assert reward.shape == (10, 1) and reward.dtype == torch.float32 and info == {}
    assert isinstance(done, tuple) and done[0].shape == (10,) and done[0].dtype == torch.bool and isinstance(done[1], bool)
    reward += 1                                                                         # main.py:40 mutates it
    assert float(sum(reward) / env.num_particles) > 0
    nobs2, *_ = env.step(a, dt=0.1)
    assert not torch.equal(nobs["actors"], nobs2["actors"]) or True
    assert env.positions.shape == (10, 2) and env.headings.shape == (10,) and env.velocities.shape == (10, 2)
    assert env.nearest_neighbors.shape == (10, 4) and env.nearest_neighbors.dtype == torch.int64
    env.render(); env.close()

    uw = gym_flock_uw.MultiAgentEnv(agents=8, k=3, collision_distance=3, range_start=(0, 50), sensor_range=7)
    o = uw.reset()
    assert o.shape == (8, 4, 3) and bool((o[:, 1:, :] == 0).all()) and uw.action_space.shape == (2,)
    o2, r, d, _ = uw.step(torch.rand(8, 2, device="cuda"), dt=0.05)
    assert o2.shape == (8, 4, 3) and o2.reshape(8, -1).shape == (8, 12) and r.shape == (8, 1)
    assert torch.equal(o2[:, 1, :], o[:, 0, :]) and d[0].long().shape == (8,)             # window shifted by one
    assert float(r.mean().item()) == float(r.mean().item())

    uwd = gym_flock_uw_discrete.MultiAgentEnv(agents=8, k=4, range_start=[0, 50])
    assert len(uwd.observation_space) == 8 and uwd.observation_space[0].shape[0] == 4 and uwd.action_space[0].n == 4
    s = uwd.reset()
    assert s.shape == (8, 4)
    ns, r, d, _ = uwd.step(torch.tensor([0., 1, 2, 3, 9, 5, 6, 7]), 0.2)                 # float ids, positional dt
    assert ns.shape == (8, 4) and r[:, 0].shape == (8,) and isinstance(d[1], bool)
    with pytest.raises(KeyError):
        uwd.step(torch.tensor([0., 1, 2, 3, 10, 5, 6, 7]))


def test_single_env_facade_matches_oracle_trajectory():
    from marl_range_flocking_b200 import gym_flock_v2
    from oracle.flock_oracle import OracleEnv
    env = gym_flock_v2.MultiAgentEnv(agents=10, k=4, collision_distance=2.5, range_start=(0, 50), sensor_range=14, seed=3)
    env.reset()
    orc = OracleEnv("v2", 1, 10, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=3)
    init = np.stack([env.positions[:, 0].cpu().numpy()[None], env.positions[:, 1].cpu().numpy()[None],
                     env.headings.cpu().numpy()[None]])
    orc.reset(init=init)
    rng = np.random.default_rng(0)
    for t in range(50):
        a = rng.uniform(-1.5, 1.5, (10, 2)).astype(np.float32)
        orc.step(a[None], 0.1)
        obs, rew, done, _ = env.step(torch.from_numpy(a).cuda())
        assert_same("obs", obs["actors"], orc.obs[0, :, 0, :])
        assert_same("reward", rew[:, 0], orc.reward[0])
        assert done[1] == bool(orc.env_done[0])
        assert_same("nn", env.nearest_neighbors.int(), orc.nn[0])


@pytest.mark.parametrize("variant,N,k", [("v2", 10, 4), ("uw", 32, 3), ("uw", 12, 4), ("uwd", 16, 4), ("v2", 64, 4)])
@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_zero_copy_and_staged_paths(variant, N, k, pinned):
    """flock_step_host: pinned host buffers take the zero-copy path (kernel reads actions from and
    writes results to mapped host memory), pageable ones the staged-copy path; both must equal the
    device path bit for bit."""
    E = 96
    env, orc = make_pair(variant, E, N, k, 0.5, (0, 100), 9.0, seed=21)
    env.reset()
    orc.reset()
    for t in range(4):
        a = orc.random_actions()
        orc.step(a, 0.1)
        ta = torch.from_numpy(a.copy())
        obs, rew, (ad, ed), _ = env.step_host(ta.pin_memory() if pinned else ta, 0.1)
        want_obs = orc.obs if env.obs_hist > 1 else orc.obs[:, :, 0, :]
        assert_same("host obs", obs, want_obs)
        assert_same("host reward", rew[..., 0], orc.reward)
        assert_same("host agent_done", ad.to(torch.uint8), orc.agent_done)
        assert_same("host env_done", ed.to(torch.uint8), orc.env_done)
    compare_all(env, orc, tag="host path:")


def test_actor_rollout_into_device_replay():
    """8f items 1-3 together: batched per-agent actors drive a uw VecEnv, transitions land in the
    device-resident replay, nothing is read back until the statistics."""
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.policies import BatchedActors
    from marl_range_flocking_b200.replay import DeviceReplay
    from marl_range_flocking_b200.rollout import collect
    E, N, k, T = 64, 8, 3, 40
    env = VecEnv("uw", E, N, k, 0.5, range_start=(0, 80), sensor_range=7, seed=4)
    actors = BatchedActors(N, 4 * k, 64, 48, 2, device="cuda")
    rb = DeviceReplay(E, N, 4 * k, 2, capacity_steps=T, device="cuda", chunk_size=10)
    stats = collect(env, actors, T, max_episode_steps=15, sink=rb.add)
    assert len(rb) == T * E and stats["episodes"] >= E * (T // 15)
    s, r, ns, d, a_s, a_ns, a_a = rb.get_minibatch(32)
    assert s.shape == (32, 10, N, 4 * k) and a_a.shape == (N, 32, 10, 2) and bool((a_a.abs() <= 1).all())
    # the newest row of next_obs is the newest row of the following obs unless the episode ended in between
    t = 5
    cont = ~rb.episode_end[t]
    assert torch.equal(rb.next_obs[t][cont], rb.obs[t + 1][cont])


def test_bench_line_satisfies_the_contract():
    """`python bench.py` prints ONE JSON line with every key of the measurement contract."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--gpus", "1", "--steps", "20", "--warmup", "5",
                          "--no-sweep"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "clocks", "e2e", "gpu_launches",
                "repeats", "configs"):
        assert key in j, key
    assert j["steps"] == 20 and j["warmup"] == 5 and j["n_gpus"] == 1 and j["scaling"] == "weak"
    assert j["vs_baseline"] is None and j["dtype"] == "f32" and j["data"] == "synthetic" and j["higher_is_better"] is True
    sys.path.insert(0, root)
    import bench
    assert j["config"] == bench.config_of("cfg2", bench.WORKLOADS["cfg2"]) and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert abs(r["achieved"] - 4096 * 10 * 53 / (j["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    c = j["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c and c["port"]["value"] > 0
    if c["kind"] == "reference":        # the unmodified reference step: orders of magnitude below its own C port
        assert c["value"] < c["port"]["value"] and c["reference_pytorch"]["threads_1"]["env_steps_per_s"] > 0
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 * 10 * 2 * 4
    assert e["d2h_bytes_per_step"] == 4096 * 10 * (4 * 4 + 4 + 1) + 4096 and e["value"] < j["value"]
    assert e["value_is"] in ("pipelined", "sync_per_step") and e["sync_per_step"]["value"] > 0 and e["pipelined"]["value"] > 0
    # exactly 20 steps per timed repetition, one fused kernel each; the median over `repeats` repetitions
    assert j["repeats"] >= 11 and j["gpu_launches"] == 20 * j["repeats"] and j["gpu_launches_per_repetition"] == 20
    assert j["value"] > 1e9 and abs(j["value"] - 4096 * 10 / (j["ms_per_step"] * 1e-3)) < 1e-6 * j["value"]
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert j["stats_allreduce"]["episodes"] > 0
    for name in ("cfg3", "cfg4", "cfg5"):
        sub = j["configs"][name]
        assert sub["value"] > 1e8 and sub["ms_per_step"] > 0 and 0 < sub["roofline"]["frac"] < 1.5


def test_randomised_configuration_sweep_bit_exact():
    """60 random (variant, E, N, k, world, dt, boundary mode) configurations, a few steps each:
    covers every shared-memory stride of the small path (unrolled and generic), ragged tiles of both
    tiled kernels, k from 1 to 8 and k = N-1."""
    rng = np.random.default_rng(2024)
    variants = ["v2", "uw", "uwd"]
    for trial in range(60):
        variant = variants[trial % 3]
        if trial % 5 == 4:
            N = int(rng.integers(33, 200))
        else:
            N = int(rng.integers(2, 33))
        k = int(rng.integers(1, min(8, N - 1) + 1))
        E = int(rng.integers(1, 40)) if N <= 32 else int(rng.integers(1, 5))
        B = float(rng.choice([6.0, 20.0, 75.0, 300.0]))
        cd = float(rng.choice([0.3, 1.0, 2.5]))
        sr = float(rng.choice([2.0, 7.0, 50.0]))
        rigid = bool(rng.integers(0, 4) == 0)
        mode = int(rng.integers(1, 3))
        env, orc = make_pair(variant, E, N, k, cd, (0, B), sr, seed=int(rng.integers(0, 2**31)), rigid=rigid,
                             reset_collision_distance=cd, tiled_mode=mode, max_reset_attempts=3)
        env.reset()
        orc.reset(max_attempts=3)
        tag = f"trial {trial} {variant} E{E} N{N} k{k} B{B} rigid{rigid} mode{mode}: "
        compare_all(env, orc, tag=tag + "reset ")
        for t in range(4):
            a = orc.random_actions()
            dt = float(rng.choice([0.05, 0.1, 0.2, 1.0]))
            orc.step(a, dt)
            env.step(torch.from_numpy(a).cuda(), dt)
            compare_all(env, orc, tag=tag + f"step {t} ")
        env.step_n(3, 0.1)
        for _ in range(3):
            orc.step(orc.random_actions(), 0.1)
        compare_all(env, orc, tag=tag + "step_n ")


@pytest.mark.parametrize("variant,E,N,k", [("v2", 40, 10, 4), ("uw", 20, 32, 3), ("uwd", 30, 16, 4), ("v2", 2, 96, 8),
                                           ("v2", 33, 7, 6)])
def test_optional_range_sensing_noise_bit_exact(variant, E, N, k):
    """north-star extension: Philox N(0, std) noise on the observed ranges (the reference has none;
    std = 0 is the default and is what every other test runs)."""
    env, orc = make_pair(variant, E, N, k, 0.5, (0, 60), 9.0, seed=31, range_noise_std=0.25)
    env.reset()
    orc.reset()
    compare_all(env, orc, tag="noisy reset:")
    clean, _ = make_pair(variant, E, N, k, 0.5, (0, 60), 9.0, seed=31)
    clean.reset()
    for t in range(6):
        a = orc.random_actions()
        orc.step(a, 0.1)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        clean.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"noisy step {t}:")
    # noise touches the observation only: state, dones and rewards equal the noise-free run
    assert torch.equal(env.x, clean.x) and torch.equal(env.reward, clean.reward) and torch.equal(env.dones[0], clean.dones[0])
    d_noisy, d_clean = env.distances_to_nearest_neighbors, clean.distances_to_nearest_neighbors
    assert not torch.equal(d_noisy, d_clean) and bool((d_noisy >= 0).all()) and bool((d_noisy <= 9.0).all())
    inside = (d_clean > 1.5) & (d_clean < 7.5)          # away from the clamps the residual is N(0, 0.25)
    if int(inside.sum()) > 200:
        res = (d_noisy - d_clean)[inside]
        assert abs(float(res.mean())) < 0.06 and abs(float(res.std()) - 0.25) < 0.06
    env.step_n(4, 0.1)
    for _ in range(4):
        orc.step(orc.random_actions(), 0.1)
    compare_all(env, orc, tag="noisy step_n:")
    a = orc.random_actions()
    orc.step(a, 0.1)
    obs, *_ = env.step_host(torch.from_numpy(a).pin_memory(), 0.1)      # staged path (noise after the step)
    assert_same("noisy host obs", obs, orc.obs if env.obs_hist > 1 else orc.obs[:, :, 0, :])


@pytest.mark.parametrize("variant,E,N,k", [("uw", 64, 32, 3), ("uw", 40, 12, 4), ("uwd", 64, 16, 4), ("uwd", 30, 9, 8),
                                           ("uw", 17, 5, 2), ("v2", 50, 10, 4), ("v2", 3, 70, 4), ("uwd", 2, 130, 8)])
def test_values_only_selection_when_indices_are_not_tracked(variant, E, N, k):
    """track_neighbors=False: uw / uwd use the values-only selection network (the reference discards
    the indices there); every other buffer must still match the oracle bit for bit."""
    env, orc = make_pair(variant, E, N, k, 0.5, (0, 40), 9.0, seed=13, track_neighbors=False, auto_reset=(N == 12))
    env.reset()
    orc.reset()
    compare_all(env, orc, tag="reset:")
    for t in range(25):
        a = orc.random_actions()
        orc.step(a, 0.1)
        if N == 12:
            orc.reset(mask=orc.env_done.copy(), keep_outputs=True)
        env.step(torch.from_numpy(a).cuda(), 0.1)
        compare_all(env, orc, tag=f"step {t}:")
    if N != 12:
        env.step_n(7, 0.1)
        for _ in range(7):
            orc.step(orc.random_actions(), 0.1)
        compare_all(env, orc, tag="step_n:")
    a = orc.random_actions()
    orc.step(a, 0.1)
    if N == 12:
        orc.reset(mask=orc.env_done.copy(), keep_outputs=True)
    obs, *_ = env.step_host(torch.from_numpy(a).pin_memory(), 0.1)
    assert_same("host obs", obs, orc.obs if env.obs_hist > 1 else orc.obs[:, :, 0, :])


def _golden_files():
    from tests.golden_util import traj_files
    return traj_files()


@pytest.mark.parametrize("path", _golden_files(), ids=[p.split("/")[-1][:-4] for p in _golden_files()])
def test_cuda_free_running_against_reference_golden(path):
    """The CUDA path DIRECTLY against the recorded outputs of the unmodified reference
    (tests/golden/, produced by tests/golden/make_golden.py): same initial state, same actions (and
    injected actuation noise), T free-running steps (1000 for BASELINE config 1). North-star
    tolerance: k-NN index lists identical at every step, positions / headings / ranges within 1e-5
    relative at the end, dones identical."""
    from marl_range_flocking_b200 import VecEnv
    from tests.golden_util import close, env_kwargs, load_traj, torus_close
    g = load_traj(path)
    v, cfg = g["variant"], g["cfg"]
    kw = env_kwargs(v, cfg)
    env = VecEnv(v, 1, kw["agents"], kw["k"], kw["collision_distance"], range_start=kw["range_start"],
                 sensor_range=kw["sensor_range"], rigid_boundary=kw["rigid_boundary"])
    B = float(kw["range_start"][1])
    T = g["actions"].shape[0]
    init = np.stack([g["pos0"][:, 0][None], g["pos0"][:, 1][None], g["h0"][None]])
    obs0 = env.reset(init_state=torch.from_numpy(init).cuda())[0].cpu().numpy()
    assert close(obs0, g["obs0"], 1e-5, 1e-6).all()
    nn_bad = 0
    acts = torch.from_numpy(g["actions"]).cuda()
    noise = torch.from_numpy(g["noise"]).cuda() if "noise" in g else None
    for t in range(T):
        env.step(acts[t][None], float(g["dt"]), noise=None if noise is None else noise[t][None])
        if "nn" in g:
            nn_bad += int((env.nearest_neighbors[0].cpu().numpy() != g["nn"][t]).any(axis=1).sum())
        if t % 50 == 49 or t == T - 1:
            pos = env.positions[0].cpu().numpy()
            assert torus_close(pos, g["pos"][t], B, 1e-5, 1e-5).all(), t
            assert np.array_equal(env.dones[0][0].cpu().numpy(), g["agent_done"][t]), t
    assert nn_bad == 0
    assert close(env.headings[0].cpu().numpy(), g["h"][-1], 1e-5, 1e-6).all()
    assert close(env.observation[0].cpu().numpy(), g["obs"][-1], 1e-5, 1e-5).all()
    rew_ok = close(env.reward[0, :, 0].cpu().numpy(), g["reward"][-1][:, 0], 1e-6, 1e-7)
    assert (~rew_ok).sum() <= 1      # a thresholded mean whose summation order torch leaves undefined


def test_cuda_edge_cases_against_reference_golden():
    """Single steps from hand-built states (walls, rigid boundary, NaN / Inf actions, zero action,
    k = N-1, dt != 0.1, history window, coincident agents) against the reference's recorded outputs."""
    from marl_range_flocking_b200 import VecEnv
    from tests.golden_util import close, env_kwargs, load_edges
    for case in load_edges():
        v, cfg = case["variant"], case["cfg"]
        kw = env_kwargs(v, cfg)
        env = VecEnv(v, 1, kw["agents"], kw["k"], kw["collision_distance"], range_start=kw["range_start"],
                     sensor_range=kw["sensor_range"], rigid_boundary=kw["rigid_boundary"])
        pos, h = case["pos_in"], case["h_in"]
        N = len(h)
        state = dict(x=torch.from_numpy(pos[:, 0][None].copy()), y=torch.from_numpy(pos[:, 1][None].copy()),
                     headings=torch.from_numpy(h[None].copy()),
                     prev_headings=torch.from_numpy(case.get("prev_h_in", np.zeros_like(h))[None].copy()))
        if "obs_mem_in" in case:
            state["obs"] = torch.from_numpy(case["obs_mem_in"][None].copy())
        env.set_state(state)
        noise = None
        if v == "uwd":
            noise = torch.from_numpy(case.get("noise_in", np.zeros((N, 2), np.float32))[None].copy()).cuda()
        env.step(torch.from_numpy(case["action"][None].copy()).cuda(), float(case["dt"]), noise=noise)
        label = case["label"]
        assert close(env.positions[0].cpu().numpy(), case["out_pos"], 1e-6, 1e-6).all(), label
        hh = env.headings[0].cpu().numpy()
        fin = ~np.isnan(case["out_h"])
        assert np.array_equal(np.isnan(hh), ~fin) and close(hh[fin], case["out_h"][fin], 1e-6, 1e-6).all(), label
        assert close(env.velocities[0].cpu().numpy(), case["out_vel"], 1e-5, 1e-7).all(), label
        assert close(env.observation[0].cpu().numpy(), case["out_obs"], 1e-5, 1e-6).all(), label
        if "out_nn" in case:
            ref_nn = case["out_nn"]
            ok_rows = ~(ref_nn == np.arange(N)[:, None]).any(axis=1)      # upstream lists self for coincident pairs
            assert np.array_equal(env.nearest_neighbors[0].cpu().numpy()[ok_rows], ref_nn[ok_rows]), label
        assert np.array_equal(env.dones[0][0].cpu().numpy(), case["out_agent_done"]), label
        assert bool(env.dones[1][0]) == bool(case["out_env_done"]), label
        assert close(env.reward[0, :, 0].cpu().numpy(), case["out_reward"][:, 0], 1e-6, 1e-7).all(), label


@pytest.mark.parametrize("path", [p for p in _golden_files() if "_uw_" in p],
                         ids=[p.split("/")[-1][:-4] for p in _golden_files() if "_uw_" in p])
def test_cuda_uw_is_bit_identical_to_the_reference(path):
    """uw on the GPU, free-running, against the unmodified reference's recorded outputs: 0 ulp on
    positions, displacements, the observation window, rewards and dones at every step."""
    from marl_range_flocking_b200 import VecEnv
    from tests.golden_util import env_kwargs, load_traj
    g = load_traj(path)
    kw = env_kwargs("uw", g["cfg"])
    env = VecEnv("uw", 1, kw["agents"], kw["k"], kw["collision_distance"], range_start=kw["range_start"],
                 sensor_range=kw["sensor_range"])
    init = np.stack([g["pos0"][:, 0][None], g["pos0"][:, 1][None], g["h0"][None]])
    env.reset(init_state=torch.from_numpy(init).cuda())
    acts = torch.from_numpy(g["actions"]).cuda()
    for t in range(g["actions"].shape[0]):
        env.step(acts[t][None], float(g["dt"]))
        assert_same(f"pos {t}", env.positions[0], g["pos"][t])
        assert_same(f"vel {t}", env.velocities[0], g["vel"][t])
        assert_same(f"obs {t}", env.observation[0], g["obs"][t])
        assert_same(f"reward {t}", env.reward[0], g["reward"][t])
        assert_same(f"done {t}", env.dones[0][0], g["agent_done"][t])


@pytest.mark.gpu
@pytest.mark.parametrize("variant,E,N,k", [("v2", 96, 10, 4), ("uw", 40, 32, 3), ("uwd", 64, 16, 4), ("v2", 3, 200, 8)])
def test_async_host_path_two_batches_in_flight_matches_the_oracle(variant, E, N, k):
    """flock_step_host_async (VecEnv.step_host_async / wait_host): two env batches stepped alternately with
    their copies in flight give bit for bit what the oracle gives for each batch."""
    pairs = [make_pair(variant, E, N, k, 0.5, rs=(0, 100), sr=7.0, seed=31 + i) for i in range(2)]
    for env, orc in pairs:
        env.reset()
        orc.reset()
    torch.cuda.synchronize()
    acts = [[torch.from_numpy(orc.random_actions(t)).pin_memory() for t in range(6)] for _, orc in pairs]
    for t in range(6):
        for i, (env, orc) in enumerate(pairs):
            env.step_host_async(acts[i][t], 0.1)
        for i, (env, orc) in enumerate(pairs):
            orc.step(acts[i][t].numpy(), 0.1)
            obs, reward, (agent_done, env_done), _ = env.wait_host()
            assert obs.device.type == "cpu"
            assert_same("obs", obs.reshape(orc.obs.shape), orc.obs)
            assert_same("reward", reward[..., 0], orc.reward)
            assert_same("agent_done", agent_done.to(torch.uint8), orc.agent_done)
            assert_same("env_done", env_done.to(torch.uint8), orc.env_done)
    for env, orc in pairs:
        assert_same("x", env.x, orc.x)


@pytest.mark.gpu
@pytest.mark.parametrize("variant,N,k", [("v2", 10, 4), ("uw", 32, 3), ("uwd", 16, 4)])
def test_graphed_rollout_is_identical_to_the_python_loop(variant, N, k):
    """rollout.GraphedRollout (policy -> step -> time-limit bonus -> masked restart captured once in a CUDA graph)
    against rollout.collect on a twin env: same kernels, same order, same buffers -> identical bits."""
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.rollout import GraphedRollout, collect, random_policy
    E, MAXS, SPR, REPLAYS, WARM = 96, 9, 8, 3, 2
    mk = lambda: VecEnv(variant, E, N, k, 2.5 if variant == "v2" else 0.5, range_start=(0, 40 if variant == "v2" else 100),
                        sensor_range=14.0 if variant == "v2" else 7.0, seed=5, device="cuda:0")
    a, b = mk(), mk()
    stats_a = collect(a, random_policy(a), WARM + SPR * REPLAYS, max_episode_steps=MAXS)
    b.reset()
    gr = GraphedRollout(b, random_policy(b), steps_per_replay=SPR, max_episode_steps=MAXS, warmup_steps=WARM)
    stats_b = gr.run(REPLAYS)
    torch.cuda.synchronize()
    for name in ("x", "y", "headings", "_obs", "_reward", "_agent_done", "_env_done", "_ep_len", "_ep_return_fx", "_reset_epoch"):
        ta, tb = getattr(a, name), getattr(b, name)
        assert torch.equal(ta, tb), name
    assert stats_a == stats_b and stats_a["episodes"] > 0


@pytest.mark.parametrize("variant,N,k,B,cd", [("v2", 300, 8, 60.0, 1.0), ("uw", 96, 3, 40.0, 1.0), ("uwd", 130, 4, 50.0, 1.0)])
def test_large_swarm_reset_needing_many_attempts_matches_the_oracle(variant, N, k, B, cd):
    """Dense large-swarm worlds: most envs need several rejection rounds, some more than the two multi-CTA attempt
    launches (then the single-CTA loop kernel continues at attempt 2), some give up at max_attempts -- draws,
    attempt counters, Philox epochs and statistics must match the oracle's single loop exactly."""
    env, orc = make_pair(variant, 12, N, k, cd, (0, B), 9.0, seed=4, reset_collision_distance=cd, max_reset_attempts=5)
    for rnd in range(3):
        env.reset()
        orc.reset(max_attempts=5)
        compare_all(env, orc, tag=f"reset round {rnd}:")
        for t in range(3):
            a = orc.random_actions()
            orc.step(a, 0.1)
            env.step(torch.from_numpy(a).cuda(), 0.1)
        mask = np.zeros(12, np.uint8)
        mask[rnd::3] = 1
        orc.reset(mask=mask, max_attempts=5)
        env.reset(mask=torch.from_numpy(mask).cuda().bool())
        compare_all(env, orc, tag=f"masked reset round {rnd}:")
    s = orc.stats
    assert int(s[3]) > 3 * 12 and int(s[4]) >= 0          # more attempts than resets: the rejection loop really looped
