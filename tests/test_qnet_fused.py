"""Fused VDN action-selection kernel (csrc/flock_qnet.cu, flock_qnet_forward through the C ABI) against the
plain PyTorch fp32 module of the same op (`policies.BatchedQNet.forward`, itself CPU-tested against the
reference's per-agent QNet layout, learners/vdn/net.py:11-58).

Tolerance (floating-point kernel, fp32 on both sides, only the summation order differs): Q-values and
GRU hidden state within 2e-5 absolute + 2e-5 relative; greedy actions must be EQUAL wherever the top-two
Q gap exceeds 1e-4 (elsewhere either maximiser is accepted).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(N, n_obs, A, rec, seed, dev):
    from marl_range_flocking_b200.policies import BatchedQNet
    torch.manual_seed(seed)
    return BatchedQNet(N, n_obs, A, recurrent=rec, device=dev)


@pytest.mark.parametrize("impl", ["tc", "fp32"])      # tensor-core kernel with split fp16 operands / fp32 CUDA-core kernel
@pytest.mark.parametrize("E,N,n_obs,A,rec", [(8192, 16, 4, 4, True), (8192, 16, 4, 4, False), (300, 5, 8, 8, True),
                                             (257, 3, 3, 5, False), (1, 2, 4, 4, True), (100, 7, 16, 16, True),
                                             (129, 4, 1, 2, True)])
def test_fused_qnet_matches_pytorch(E, N, n_obs, A, rec, impl):
    dev = torch.device("cuda:0")
    net = _net(N, n_obs, A, rec, 3 + E, dev)
    torch.manual_seed(E + N)
    obs = torch.rand(E, N, n_obs, device=dev) * 7.0
    hidden = torch.randn(E, N, 32, device=dev) * 0.5
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            q_ref, h_ref = net(obs, hidden)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    q, h = net.forward_fused(obs, hidden, impl=impl)
    act, h2 = net.sample_action_fused(obs, hidden, epsilon=0.0, impl=impl)
    torch.cuda.synchronize()
    assert torch.allclose(q, q_ref, atol=2e-5, rtol=2e-5), (q - q_ref).abs().max().item()
    if rec:
        assert torch.allclose(h, h_ref, atol=2e-5, rtol=2e-5), (h - h_ref).abs().max().item()
        assert torch.equal(h, h2)
    top2 = q_ref.topk(2, dim=2).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-4
    assert clear.float().mean() > 0.9
    assert torch.equal(act[clear], q_ref.argmax(dim=2).float()[clear])
    assert torch.equal(act, q.argmax(dim=2).float())              # exactly the kernel's own Q-values, first maximum


def test_fused_qnet_epsilon_greedy_is_per_env_reproducible_and_sharding_invariant():
    dev = torch.device("cuda:0")
    E, N, A = 4096, 8, 4
    net = _net(N, 4, A, True, 9, dev)
    obs = torch.rand(E, N, 4, device=dev) * 7.0
    hidden = torch.zeros(E, N, 32, device=dev)
    greedy, _ = net.sample_action_fused(obs, hidden, 0.0)
    a1, _ = net.sample_action_fused(obs, hidden, 0.3, step=5, seed=77)
    a2, _ = net.sample_action_fused(obs, hidden, 0.3, step=5, seed=77)
    a3, _ = net.sample_action_fused(obs, hidden, 0.3, step=6, seed=77)
    allr, _ = net.sample_action_fused(obs, hidden, 1.0, step=5, seed=77)
    torch.cuda.synchronize()
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    assert ((a1 >= 0) & (a1 < A) & (a1 == a1.round())).all()
    # one decision per env (net.py:54): an env is either entirely greedy or entirely the random draw
    env_greedy = (a1 == greedy).all(dim=1)
    env_random = (a1 == allr).all(dim=1)
    assert (env_greedy | env_random).all()
    frac = 1.0 - env_greedy.float().mean().item()
    assert 0.2 < frac < 0.4                                        # explores w.p. epsilon = 0.3 (minus coincidences)
    counts = torch.bincount(allr.flatten().long(), minlength=A).float() / allr.numel()
    assert (counts - 1.0 / A).abs().max() < 0.02                   # uniform ids
    # the second half of the envs evaluated as its own shard gives the same actions
    half = E // 2
    b, _ = net.sample_action_fused(obs[half:].contiguous(), hidden[half:].contiguous(), 0.3, step=5, seed=77, env_offset=half)
    torch.cuda.synchronize()
    assert torch.equal(b, a1[half:])


def test_fused_qnet_drives_the_discrete_env():
    from marl_range_flocking_b200 import VecEnv
    dev = torch.device("cuda:0")
    E, N, k = 512, 16, 4
    env = VecEnv("uwd", E, N, k, 0.5, range_start=(0, 100), sensor_range=7.0, seed=2, device="cuda:0")
    obs = env.reset()
    net = _net(N, k, k, True, 1, dev)
    hidden = net.init_hidden(E)
    for t in range(5):
        act, hidden = net.sample_action_fused(obs, hidden, 0.1, step=t, seed=4)
        obs, reward, dones, _ = env.step(act, 0.1)
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(hidden).all()
