"""CPU: independent checks of the C oracle (beyond the reference golden vectors).

A numpy float32 restatement of the v2/uw/uwd step written directly from SURVEY appendix A (IEEE
single operations, brute-force stable sort for the k-NN) must agree with the C oracle BIT FOR BIT,
plus the domain properties the reference implies (permutation equivariance, sorted clamped ranges,
window roll, Philox reset bounds / determinism / sharding invariance).
"""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import flock_oracle as fo
from oracle.flock_oracle import OracleEnv

f32 = np.float32


def fmaf_np(a, b, c):
    """float32 fma through float64: the product is exact in binary64, so only the final conversion
    rounds (up to a ~2^-29 double-rounding corner that the fixed seeds below do not hit)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def np_step(variant, x, y, h, prev_h, act, dt, B, sr, cd, vmax, k, periodic, noise=None):
    """One env, numpy float32, straight from SURVEY appendix A.1 / A.3 / A.4."""
    x, y, h, prev_h = (a.astype(f32).copy() for a in (x, y, h, prev_h))
    dt = f32(dt)
    N = x.size
    with np.errstate(all="ignore"):
        if variant == "v2":
            w = np.clip(act[:, 1], f32(-np.pi / 2), f32(np.pi / 2))
            w = np.where(np.isnan(act[:, 1]), act[:, 1], w).astype(f32)
            h = (h + w * dt).astype(f32)
            u = np.clip(act[:, 0], f32(0.005), f32(vmax))
            u = np.where(np.isnan(act[:, 0]), act[:, 0], u).astype(f32)
            sn, cs = fo.sincosf(h)
            vx, vy = (u * cs).astype(f32), (u * sn).astype(f32)
        elif variant == "uw":
            n = np.sqrt(fmaf_np(act[:, 1], act[:, 1], (act[:, 0] * act[:, 0]).astype(f32))).astype(f32)
            vx, vy = (act[:, 0] / n).astype(f32), (act[:, 1] / n).astype(f32)
        else:
            ids = np.clip(np.nan_to_num(act, nan=0.0), 0, 9).astype(np.int64)
            mu_u = np.where(ids < 5, f32(0.2), f32(0.6)).astype(f32)
            mu_w = np.array([-1.2, -0.5, 0.0, 0.5, 1.2], f32)[ids % 5]
            u = (mu_u + noise[:, 0]).astype(f32)
            w = np.clip((mu_w + noise[:, 1]).astype(f32), f32(-0.025), f32(0.025))
            h = (h + w * dt).astype(f32)
            u = np.clip(u, f32(5e-6), f32(vmax))
            sn, cs = fo.sincosf(h)
            vx, vy = (u * cs).astype(f32), (u * sn).astype(f32)
            n = np.sqrt(fmaf_np(vy, vy, (vx * vx).astype(f32))).astype(f32)
            vx, vy = (vx / n).astype(f32), (vy / n).astype(f32)
        vx, vy = np.nan_to_num(vx).astype(f32), np.nan_to_num(vy).astype(f32)
        vx, vy = (vx * dt).astype(f32), (vy * dt).astype(f32)
        x, y = (x + vx).astype(f32), (y + vy).astype(f32)
    Bf = f32(B)
    for c in (x, y):
        c[...] = np.where(c < Bf, c, f32(0.001))
        c[...] = np.where(c > 0, c, Bf)
    dx = np.abs(x[:, None] - x[None, :]).astype(f32)
    dy = np.abs(y[:, None] - y[None, :]).astype(f32)
    if periodic:
        half = f32(B / 2)
        dx = np.where(dx > half, (Bf - dx).astype(f32), dx)
        dy = np.where(dy > half, (Bf - dy).astype(f32), dy)
    d2 = (((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32) if periodic
          else fmaf_np(dy, dy, (dx * dx).astype(f32)))
    nn = np.zeros((N, k), np.int32)
    dk = np.zeros((N, k), f32)
    for i in range(N):
        cand = [j for j in range(N) if j != i]
        cand.sort(key=lambda j: (d2[i, j], j))          # (d2, j) ascending: ties -> lower index
        nn[i] = cand[:k]
        dk[i] = np.clip(np.sqrt(d2[i, cand[:k]]).astype(f32), 0, f32(sr))
    coll = (dk < f32(cd)).any(axis=1)
    if variant == "v2":
        rew = np.where(coll, f32(-5), f32(0.01)).astype(f32)
    elif variant == "uw":
        sx = f32(0); sy = f32(0)
        for j in range(N):
            sx = f32(sx + x[j]); sy = f32(sy + y[j])
        cx, cy = f32(sx / f32(N)), f32(sy / f32(N))
        ddx, ddy = (x - cx).astype(f32), (y - cy).astype(f32)
        dc = np.sqrt(fmaf_np(ddy, ddy, (ddx * ddx).astype(f32))).astype(f32)
        rcom = np.where(dc < f32(cd * 4), f32(0.01), f32(0)).astype(f32)
        rang = np.where(np.abs((prev_h - h).astype(f32)) > f32(0.27), f32(-0.01), f32(0.001)).astype(f32)
        rew = ((np.where(coll, f32(-5), f32(0.01)).astype(f32) + rcom).astype(f32) + rang).astype(f32)
        prev_h = h.copy()
    else:
        sh = f32(0)
        for j in range(N):
            sh = f32(sh + h[j])
        hm = f32(sh / f32(N))
        ral = np.where(np.abs((hm - h).astype(f32)) > f32(0.20), f32(0), f32(0.1)).astype(f32)
        rew = (np.where(coll, f32(-9), f32(0)).astype(f32) + ral).astype(f32)
        prev_h = h.copy()
    return dict(x=x, y=y, h=h, prev_h=prev_h, vx=vx, vy=vy, nn=nn, dk=dk, coll=coll, rew=rew)


def _bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


@pytest.mark.parametrize("variant,N,k,B,sr,cd", [("v2", 10, 4, 50, 14, 2.5), ("v2", 7, 6, 8, 3, 1.0),
                                                 ("uw", 12, 3, 30, 7, 1.5), ("uwd", 9, 4, 25, 7, 1.0),
                                                 ("v2", 40, 8, 60, 20, 0.7)])
def test_numpy_restatement_agrees_bit_for_bit(variant, N, k, B, sr, cd):
    E, T = 6, 40
    rng = np.random.default_rng(N * 31 + k)
    env = OracleEnv(variant, E, N, k, cd, range_start=(0, B), sensor_range=sr, seed=5)
    env.reset()
    for t in range(T):
        state = [(env.x[e].copy(), env.y[e].copy(), env.h[e].copy(), env.prev_h[e].copy()) for e in range(E)]
        if variant == "uwd":
            a = rng.integers(0, 10, (E, N)).astype(f32)
            nz = (rng.standard_normal((E, N, 2)) * 0.1).astype(f32)
        else:
            a = rng.uniform(-1.5, 1.5, (E, N, 2)).astype(f32)
            nz = None
        dt = 0.1 if t % 4 else 0.3
        env.step(a, dt, noise=nz)
        for e in range(E):
            r = np_step(variant, *state[e], a[e], dt, B, sr, cd, 2.5, k, variant == "v2",
                        None if nz is None else nz[e])
            assert np.array_equal(_bits(r["x"]), _bits(env.x[e])) and np.array_equal(_bits(r["y"]), _bits(env.y[e]))
            assert np.array_equal(_bits(r["h"]), _bits(env.h[e]))
            assert np.array_equal(_bits(r["vx"]), _bits(env.vx[e])) and np.array_equal(_bits(r["vy"]), _bits(env.vy[e]))
            assert np.array_equal(r["nn"], env.nn[e])
            assert np.array_equal(_bits(r["dk"]), _bits(env.obs[e, :, 0, :]))
            assert np.array_equal(r["coll"].astype(np.uint8), env.agent_done[e])
            assert bool(r["coll"].any()) == bool(env.env_done[e])
            assert np.array_equal(_bits(r["rew"]), _bits(env.reward[e]))
            assert np.array_equal(_bits(r["prev_h"]), _bits(env.prev_h[e])) or variant == "v2"


@settings(max_examples=40, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**31 - 1), N=st.integers(3, 20), k=st.integers(1, 8))
def test_permuting_agents_permutes_outputs(seed, N, k):
    k = min(k, N - 1)
    rng = np.random.default_rng(seed)
    B = 40.0
    x, y = rng.uniform(0.01, B, (2, 1, N)).astype(f32)
    h = rng.uniform(0, 4, (1, N)).astype(f32)
    act = rng.uniform(-1.5, 1.5, (1, N, 2)).astype(f32)
    perm = rng.permutation(N)
    a = OracleEnv("v2", 1, N, k, 1.0, range_start=(0, B), sensor_range=15)
    b = OracleEnv("v2", 1, N, k, 1.0, range_start=(0, B), sensor_range=15)
    a.set_state(x, y, h)
    b.set_state(x[:, perm], y[:, perm], h[:, perm])
    a.step(act, 0.1)
    b.step(act[:, perm], 0.1)
    assert np.array_equal(_bits(a.x[0, perm]), _bits(b.x[0])) and np.array_equal(_bits(a.reward[0, perm]), _bits(b.reward[0]))
    assert np.array_equal(_bits(a.obs[0, perm]), _bits(b.obs[0]))       # ranges do not depend on labels
    d = a.obs[0, :, 0, :]
    assert (np.diff(d, axis=1) >= 0).all() and (d >= 0).all() and (d <= 15).all()
    # neighbour SETS map through the permutation whenever the row has no exact distance tie
    inv = np.argsort(perm)
    for i in range(N):
        if len(set(d[perm[i]].tolist())) == k and d[perm[i]].max() < 15:
            assert set(inv[a.nn[0, perm[i]]].tolist()) == set(b.nn[0, i].tolist())


def test_uw_window_rolls_newest_first():
    env = OracleEnv("uw", 3, 8, 3, 0.5, range_start=(0, 60), sensor_range=7, seed=2)
    env.reset()
    assert (env.obs[:, :, 1:, :] == 0).all()
    rows = [env.obs[:, :, 0, :].copy()]
    rng = np.random.default_rng(0)
    for t in range(6):
        env.step(rng.uniform(-1, 1, (3, 8, 2)).astype(f32), 0.1)
        rows.append(env.obs[:, :, 0, :].copy())
        for s in range(4):
            want = rows[-1 - s] if len(rows) > s else np.zeros_like(rows[0])
            assert np.array_equal(env.obs[:, :, s, :], want)


def test_philox_reset_is_collision_free_bounded_and_shard_invariant():
    E, N = 300, 10
    full = OracleEnv("v2", E, N, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=99)
    assert full.reset() == 0
    assert not full.env_done.any() and not full.agent_done.any()
    assert (full.x > 0).all() and (full.x <= 50).all() and (full.y > 0).all() and (full.y <= 50).all()
    assert (full.h > 0).all() and (full.h <= np.float32(1.5 * np.pi)).all()
    assert (full.reset_epoch >= 1).all() and full.reset_epoch.max() > 1      # some envs needed a redraw
    lo = OracleEnv("v2", 100, N, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=99, env_offset=0)
    hi = OracleEnv("v2", 200, N, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=99, env_offset=100)
    lo.reset(); hi.reset()
    assert np.array_equal(np.concatenate([lo.x, hi.x]), full.x) and np.array_equal(np.concatenate([lo.h, hi.h]), full.h)
    again = OracleEnv("v2", E, N, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=99)
    again.reset()
    assert np.array_equal(again.x, full.x)
    other = OracleEnv("v2", E, N, 4, 2.5, range_start=(0, 50), sensor_range=14, seed=100)
    other.reset()
    assert not np.array_equal(other.x, full.x)
    # uw starts in the half box (gym_flock_uw.py:87-89), uwd rejects with collision distance 4
    uw = OracleEnv("uw", 50, 8, 3, 1.0, range_start=(0, 50), sensor_range=7, seed=1)
    uw.reset()
    assert (uw.x <= 25).all() and (uw.y <= 25).all()
    uwd = OracleEnv("uwd", 50, 8, 4, 1.0, range_start=(0, 100), sensor_range=7, seed=1)
    uwd.reset()
    assert (uwd.obs[:, :, 0, 0] >= 4.0).all()


def test_reset_gives_up_after_max_attempts_on_impossible_density():
    env = OracleEnv("v2", 4, 32, 3, 3.0, range_start=(0, 10), sensor_range=7, seed=1)
    assert env.reset(max_attempts=5) == 4
    assert env.env_done.all() and (env.reset_epoch == 5).all() and int(env.stats[4]) == 4


def test_actuation_noise_statistics():
    w = np.random.default_rng(0).integers(0, 2**32, 400000, dtype=np.uint64).astype(np.uint32)
    z = fo.normal2(w)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3 and np.isfinite(z).all()
    assert abs(np.mean(np.abs(z) > 1.96) - 0.05) < 3e-3
