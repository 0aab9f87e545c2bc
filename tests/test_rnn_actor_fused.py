"""Fused recurrent MADDPG actor (csrc/flock_rnn_actor.cu, flock_rnn_actor_forward through the C ABI) against plain
PyTorch references of the same op: `policies.BatchedRnnActors.forward` (fp32; CPU-tested against per-agent modules
with the reference's layer layout, learners/maddpg_official_rnn/net.py:14-72) and an fp32 emulation with the
kernel's operand rounding.

Tolerances: the GRU front end is fp32 on both sides -> next hidden state within 2e-5; actions (bf16 MLP operands):
<= 4e-3 absolute against the bf16-operand emulation (a wrong operand layout gives O(1) errors), <= 4e-2 against the
fp32 module (linear in [0, 1], angular in [-1.5, 1.5])."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _net(N, n_obs, seed, dev):
    from marl_range_flocking_b200.policies import BatchedRnnActors
    torch.manual_seed(seed)
    a = BatchedRnnActors(N, n_obs, device=dev)
    with torch.no_grad():                          # O(1) heads so that the comparison is not vacuous
        a.wl.uniform_(-0.1, 0.1)
        a.wa.uniform_(-0.1, 0.1)
        a.bl.uniform_(-0.2, 0.2)
        a.ba.uniform_(-0.2, 0.2)
    return a


def _emulate(a, obs, hidden):
    r = lambda t: t.bfloat16().float()
    E, N = obs.shape[:2]
    _, h = a(obs, hidden)                          # fp32 front end
    x = h.transpose(0, 1)
    x = F.relu(torch.baddbmm(a.b1, r(x), r(a.w1)))
    x = F.relu(torch.baddbmm(a.b2, r(x), r(a.w2)))
    lin = (torch.tanh(torch.baddbmm(a.bl, x, a.wl)) + 1) / 2
    ang = torch.tanh(torch.baddbmm(a.ba, x, a.wa)) * 1.5
    return torch.cat([lin, ang], dim=-1).transpose(0, 1).contiguous(), h


@pytest.mark.parametrize("impl", ["tc", "fp32"])      # front end: tensor cores with split fp16 operands / fp32 CUDA cores
@pytest.mark.parametrize("E,N,n_obs", [(128, 1, 4), (300, 5, 4), (4096, 10, 4), (77, 3, 8), (1, 2, 12), (4096, 32, 16)])
def test_fused_rnn_actor_matches_pytorch(E, N, n_obs, impl):
    dev = torch.device("cuda:0")
    a = _net(N, n_obs, 7 + E, dev)
    torch.manual_seed(E + N)
    obs = torch.rand(E, N, n_obs, device=dev) * 14.0
    hidden = torch.randn(E, N, 32, device=dev) * 0.5
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want32, h32 = a(obs, hidden)
            want16, _ = _emulate(a, obs, hidden)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    got, h = a.forward_fused(obs, hidden, impl=impl)
    torch.cuda.synchronize()
    assert got.shape == (E, N, 2) and h.shape == (E, N, 32)
    assert torch.isfinite(got).all()
    assert torch.allclose(h, h32, atol=2e-5, rtol=2e-5), (h - h32).abs().max().item()
    err16 = (got - want16).abs().max().item()
    err32 = (got - want32).abs().max().item()
    assert err16 <= 4e-3, (err16, err32)
    assert err32 <= 4e-2, (err16, err32)
    assert bool(((got[..., 0] >= 0) & (got[..., 0] <= 1)).all()) and bool((got[..., 1].abs() <= 1.5).all())


def test_fused_rnn_actor_drives_the_env_with_the_state_updated_in_place():
    from marl_range_flocking_b200 import VecEnv
    dev = torch.device("cuda:0")
    E, N, k = 512, 10, 4
    env = VecEnv("v2", E, N, k, 2.5, range_start=(0, 50), sensor_range=14.0, seed=3, device="cuda:0")
    obs = env.reset()
    a = _net(N, k, 1, dev)
    hidden = a.init_hidden(E)
    ref_h = hidden.clone()
    acts = torch.empty(E, N, 2, device=dev)
    for t in range(4):
        with torch.no_grad():
            _, ref_h = a(obs, ref_h)
        a.forward_fused(obs, hidden, out=acts, hidden_out=hidden)          # in place
        assert torch.allclose(hidden, ref_h, atol=1e-4, rtol=1e-4)
        obs, reward, dones, _ = env.step(acts, 0.1)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(hidden).all()


def test_fused_rnn_actor_ou_exploration_noise():
    """flock_rnn_actor_forward_ou: sigma = 0 gives the deterministic recurrence exactly and actions = mu + x;
    sigma > 0 gives unit-normal increments, reproducible per (seed, step)."""
    dev = torch.device("cuda:0")
    E, N = 2048, 10
    a = _net(N, 4, 2, dev)
    torch.manual_seed(4)
    obs = torch.rand(E, N, 4, device=dev) * 14.0
    hidden = torch.randn(E, N, 32, device=dev) * 0.3
    mu_act, _ = a.forward_fused(obs, hidden)
    mu_act = mu_act.clone()
    theta, mu, sigma, dt = 0.15, 0.1, 0.2, 1e-2
    x = torch.full((E, N, 2), 0.5, device=dev)
    ref = x + theta * dt * (mu - x)
    out, _ = a.forward_fused(obs, hidden, ou_state=x, ou_theta=theta, ou_mu=mu, ou_sigma=0.0, ou_dt=dt, seed=3, step=0)
    assert torch.allclose(x, ref, atol=1e-6) and torch.allclose(out, mu_act + x, atol=1e-6)
    x1, x2, x3 = (torch.zeros(E, N, 2, device=dev) for _ in range(3))
    kw = dict(ou_theta=theta, ou_mu=mu, ou_sigma=sigma, ou_dt=dt, seed=3)
    a.forward_fused(obs, hidden, ou_state=x1, step=1, **kw)
    a.forward_fused(obs, hidden, ou_state=x2, step=1, **kw)
    a.forward_fused(obs, hidden, ou_state=x3, step=2, **kw)
    torch.cuda.synchronize()
    assert torch.equal(x1, x2) and not torch.equal(x1, x3)
    z = (x1 - theta * dt * mu) / (sigma * dt ** 0.5)
    assert abs(z.mean().item()) < 0.02 and abs(z.std().item() - 1.0) < 0.02
