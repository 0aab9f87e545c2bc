"""Fused tensor-core actor kernel (csrc/flock_actor.cu, flock_actor_forward through the C ABI) against
plain PyTorch references of the same op: the fp32 module (`policies.BatchedActors.forward`, itself
CPU-tested against the reference's ActorNetwork layout) and an fp32 emulation of the kernel's
arithmetic (bf16-rounded matrix operands, fp32 accumulation / LayerNorm / head).

Tolerances (floating-point kernel; outputs are tanh values in [-1, 1]):
  * vs. the bf16-operand emulation: 4e-3 absolute -- only summation order and rsqrt/tanh
    implementation differ; a wrong operand layout gives O(1) errors;
  * vs. the fp32 module: 4e-2 absolute -- the cost of bf16 operands (2^-9 relative per operand)
    through two 400/300-wide layers.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _actors(N, in_dims, seed, device):
    from marl_range_flocking_b200.policies import BatchedActors
    torch.manual_seed(seed)
    a = BatchedActors(N, in_dims, 400, 300, 2, device=device)
    with torch.no_grad():                      # non-trivial LayerNorm affine and an O(1) head
        a.g1.uniform_(0.5, 1.5)
        a.be1.uniform_(-0.3, 0.3)
        a.g2.uniform_(0.5, 1.5)
        a.be2.uniform_(-0.3, 0.3)
        a.w3.uniform_(-0.15, 0.15)
        a.b3.uniform_(-0.2, 0.2)
        a.w1.mul_(4.0)
    return a


def _emulate_bf16_operands(a, obs):
    """The kernel's arithmetic in fp32: weights and biases centred over the output features at pack time (LayerNorm
    removes that mean anyway), matrix operands rounded to bf16, everything else fp32."""
    r = lambda t: t.bfloat16().float()
    c = lambda t: t - t.mean(dim=2, keepdim=True)
    E, N = obs.shape[:2]
    x = r(obs.reshape(E, N, -1).transpose(0, 1))
    x = torch.baddbmm(c(a.b1), x, r(c(a.w1)))
    x = F.relu(F.layer_norm(x, x.shape[-1:]) * a.g1 + a.be1)
    x = torch.baddbmm(c(a.b2), r(x), r(c(a.w2)))
    x = F.relu(F.layer_norm(x, x.shape[-1:]) * a.g2 + a.be2)
    return torch.tanh(torch.baddbmm(a.b3, x, a.w3)).transpose(0, 1).contiguous()


@pytest.mark.parametrize("E,N,in_dims", [(128, 1, 12), (64, 10, 4), (300, 5, 12), (4096, 32, 12), (77, 3, 14), (200, 4, 9), (1, 2, 12)])
def test_fused_actor_matches_pytorch(E, N, in_dims):
    dev = torch.device("cuda:0")
    a = _actors(N, in_dims, 11 + E, dev)
    torch.manual_seed(E * 7 + N)
    obs = torch.rand(E, N, in_dims, device=dev) * 7.0          # ranges in [0, sensor_range]
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want32 = a(obs)
            want16 = _emulate_bf16_operands(a, obs)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    got = a.forward_fused(obs)
    torch.cuda.synchronize()
    assert got.shape == (E, N, 2) and got.dtype == torch.float32
    assert torch.isfinite(got).all()
    assert want32.abs().max() > 0.2                              # the comparison is not vacuous
    err16 = (got - want16).abs().max().item()
    err32 = (got - want32).abs().max().item()
    assert err16 <= 4e-3, (err16, err32)
    assert err32 <= 4e-2, (err16, err32)


def test_fused_actor_drives_the_env_and_repacks_after_an_update():
    from marl_range_flocking_b200 import VecEnv
    dev = torch.device("cuda:0")
    E, N, k = 256, 32, 3
    env = VecEnv("uw", E, N, k, 0.5, range_start=(0, 200), sensor_range=7.0, seed=5, device="cuda:0")
    obs = env.reset()
    a = _actors(N, 4 * k, 3, dev)
    acts = torch.empty(E, N, 2, device=dev)
    for _ in range(4):
        a.forward_fused(obs, out=acts)
        obs, reward, dones, _ = env.step(acts, 0.1)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and acts.abs().max() <= 1.0
    with torch.no_grad():
        before = a.forward_fused(obs).clone()
        a.w3.mul_(-1.0)
        a.b3.mul_(-1.0)
        a.pack_fused()
        after = a.forward_fused(obs)
    torch.cuda.synchronize()
    assert torch.allclose(after, -before, atol=1e-6)             # tanh is odd: the new parameters are in use


def test_fused_actor_rejects_unsupported_shapes():
    from marl_range_flocking_b200 import _lib
    from marl_range_flocking_b200.policies import BatchedActors
    dev = torch.device("cuda:0")
    with pytest.raises(_lib.FlockError):
        BatchedActors(2, 12, 256, 128, 2, device=dev).pack_fused()
    with pytest.raises(_lib.FlockError):
        BatchedActors(2, 15, 400, 300, 2, device=dev).pack_fused()


def test_fused_ou_exploration_noise():
    """flock_actor_forward_ou: the Ornstein-Uhlenbeck exploration noise of the shared-critic learner
    (OUActionNoiseGPU, learners/maddpg_shared_critic/utils.py:6-21) fused into the actor launch.
    sigma = 0: the state follows the deterministic recurrence exactly (fp32, 1e-6); sigma > 0: increments have the
    right mean / variance (2.6e5 samples), draws are reproducible per (seed, step) and invariant under env sharding;
    actions = mu + x."""
    dev = torch.device("cuda:0")
    E, N = 4096, 32
    a = _actors(N, 12, 5, dev)
    torch.manual_seed(2)
    obs = torch.rand(E, N, 12, device=dev) * 7.0
    mu_act = a.forward_fused(obs).clone()
    theta, mu, sigma, dt = 0.2, 0.3, 0.15, 1e-2
    # deterministic part
    x = torch.full((E, N, 2), -1.0, device=dev)
    ref = x.clone()
    for t in range(3):
        out = a.forward_fused(obs, ou_state=x, ou_theta=theta, ou_mu=mu, ou_sigma=0.0, ou_dt=dt, seed=9, step=t)
        ref = ref + theta * dt * (mu - ref)
        assert torch.allclose(x, ref, atol=1e-6) and torch.allclose(out, mu_act + x, atol=1e-6)
    # stochastic part
    x0 = torch.zeros(E, N, 2, device=dev)
    x1 = x0.clone()
    out1 = a.forward_fused(obs, ou_state=x1, ou_theta=theta, ou_mu=mu, ou_sigma=sigma, ou_dt=dt, seed=9, step=4).clone()
    x2 = x0.clone()
    a.forward_fused(obs, ou_state=x2, ou_theta=theta, ou_mu=mu, ou_sigma=sigma, ou_dt=dt, seed=9, step=4)
    x3 = x0.clone()
    a.forward_fused(obs, ou_state=x3, ou_theta=theta, ou_mu=mu, ou_sigma=sigma, ou_dt=dt, seed=9, step=5)
    torch.cuda.synchronize()
    assert torch.equal(x1, x2) and not torch.equal(x1, x3)
    assert torch.allclose(out1, mu_act + x1, atol=1e-6)
    z = (x1 - theta * dt * mu) / (sigma * dt ** 0.5)               # the standard normals that were drawn
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01
    assert abs((z[..., 0] * z[..., 1]).mean().item()) < 0.01        # the two action dims are independent
    half = E // 2
    xs = x0[half:].clone()
    a.forward_fused(obs[half:].contiguous(), ou_state=xs, ou_theta=theta, ou_mu=mu, ou_sigma=sigma, ou_dt=dt, seed=9,
                    step=4, env_offset=half)
    torch.cuda.synchronize()
    assert torch.equal(xs, x1[half:])


def test_graphed_rollout_with_the_fused_actor_matches_the_python_loop():
    """policy = fused actor kernel inside rollout.GraphedRollout (everything captured in one CUDA graph) against
    rollout.collect with the same policy: the kernels are deterministic, so the final states are identical."""
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.rollout import GraphedRollout, collect
    dev = torch.device("cuda:0")
    E, N, k = 256, 8, 3
    mk = lambda: VecEnv("uw", E, N, k, 0.5, range_start=(0, 60), sensor_range=7.0, seed=8, device="cuda:0")
    a, b = mk(), mk()
    actors = _actors(N, 4 * k, 13, dev)
    buf_a, buf_b = torch.empty(E, N, 2, device=dev), torch.empty(E, N, 2, device=dev)
    stats_a = collect(a, lambda obs: actors.forward_fused(obs, out=buf_a), 2 + 3 * 8, max_episode_steps=11)
    b.reset()
    gr = GraphedRollout(b, lambda obs: actors.forward_fused(obs, out=buf_b), steps_per_replay=8, max_episode_steps=11,
                        warmup_steps=2)
    stats_b = gr.run(3)
    torch.cuda.synchronize()
    for name in ("x", "y", "_obs", "_reward", "_env_done", "_ep_len", "_ep_return_fx"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(buf_a, buf_b) and stats_a == stats_b


def test_graphed_rollout_with_fused_ou_noise_draws_fresh_normals_every_replay():
    """A host `step` is frozen inside a captured CUDA graph: every replay would redraw the same normals. With
    `counters=env.noise_counters` (per-env DEVICE counters ep_len / reset_epoch added inside the kernel) the draws
    advance with the env: (a) graph replay == the same steps launched eagerly, bit for bit; (b) the increments
    of different replays differ; (c) with the host step alone they repeat -- which is what the counters are for."""
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200.rollout import GraphedRollout, collect
    dev = torch.device("cuda:0")
    E, N, k, SPR = 128, 6, 3, 4
    mk = lambda: VecEnv("uw", E, N, k, 0.5, range_start=(0, 80), sensor_range=7.0, seed=3, device="cuda:0")
    actors = _actors(N, 4 * k, 5, dev)
    kw = dict(ou_theta=0.0, ou_mu=0.0, ou_sigma=0.15, ou_dt=1e-2, seed=21, step=0)    # theta = 0: x is the plain sum of the draws
    actors.forward_fused(torch.zeros(E, N, 4 * k, device=dev))     # pack + one-time kernel configuration outside any capture
    torch.cuda.synchronize()

    def run(env, ou, buf, with_counters, graphed):
        pol = lambda obs: actors.forward_fused(obs, out=buf, ou_state=ou, counters=env.noise_counters if with_counters else None, **kw)
        snaps = []
        if graphed:
            env.reset()
            gr = GraphedRollout(env, pol, steps_per_replay=SPR, max_episode_steps=1000, warmup_steps=0)
            for _ in range(3):
                gr.run(1)
                snaps.append(ou.clone())
        else:
            for r in range(3):
                collect(env, pol, SPR, max_episode_steps=1000, reset_first=(r == 0))
                snaps.append(ou.clone())
        torch.cuda.synchronize()
        return snaps

    z = lambda: torch.zeros(E, N, 2, device=dev)
    g = run(mk(), z(), torch.empty(E, N, 2, device=dev), True, True)
    e = run(mk(), z(), torch.empty(E, N, 2, device=dev), True, False)
    for a, b in zip(g, e):
        assert torch.equal(a, b)                                   # (a) replay-safe: graph == eager
    inc = [g[0], g[1] - g[0], g[2] - g[1]]                         # sum of SPR draws per replay
    assert not torch.allclose(inc[0], inc[1]) and not torch.allclose(inc[1], inc[2])          # (b)
    zz = torch.cat([i.flatten() for i in inc]) / (0.15 * 1e-1 * SPR ** 0.5)
    assert abs(zz.mean().item()) < 0.05 and abs(zz.std().item() - 1.0) < 0.05
    frozen = run(mk(), z(), torch.empty(E, N, 2, device=dev), False, True)
    live = frozen[0] != 0                                          # (c) host step only: identical draws every launch
    assert torch.allclose((frozen[1] - frozen[0])[live], frozen[0][live], rtol=1e-4, atol=1e-6)


def test_bf16_actor_leaves_the_closed_loop_episode_statistics_unchanged():
    """The fused actor rounds its MLP operands to bf16 (<= 4e-2 on a tanh output, 1.6e-2 typical). Closed loop, that
    must not change what a learner sees in aggregate: the same policy run in fp32 PyTorch and through the fused kernel
    on twin envs gives the same episode statistics (thousands of episodes; individual trajectories do diverge)."""
    from marl_range_flocking_b200 import VecEnv
    dev = torch.device("cuda:0")
    E, N, k, T = 2048, 8, 3, 160
    mk = lambda: VecEnv("uw", E, N, k, 1.5, range_start=(0, 30), sensor_range=7.0, seed=77, device="cuda:0", auto_reset=True,
                        max_reset_attempts=16, reset_collision_distance=1.5)
    actors = _actors(N, 4 * k, 21, dev)
    stats = []
    for fused in (False, True):
        env = mk()
        env.reset()
        buf = torch.empty(E, N, 2, device=dev)
        rew = torch.zeros((), device=dev, dtype=torch.float64)
        with torch.no_grad():
            for t in range(T):
                a = actors.forward_fused(env.observation, out=buf) if fused else actors(env.observation)
                _, r, _, _ = env.step(a, 0.1)
                rew += r.double().sum()
        torch.cuda.synchronize()
        s = env.stats()
        stats.append((s["episodes"], s["mean_episode_length"], s["mean_episode_return"], float(rew) / (E * N * T)))
    (e0, l0, r0, m0), (e1, l1, r1, m1) = stats
    assert e0 > 2000 and e1 > 2000
    assert abs(e1 - e0) / e0 < 0.05 and abs(l1 - l0) / l0 < 0.05, stats
    assert abs(r1 - r0) <= 0.05 * abs(r0) + 0.05 and abs(m1 - m0) <= 0.05 * abs(m0) + 1e-3, stats
