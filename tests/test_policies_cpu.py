"""CPU: the batched policy helpers ("next" row 8f-2) against per-agent torch modules shaped like the
reference's ActorNetwork (ddpg_network.py:85-141) and QNet (vdn/net.py:11-37)."""
import torch
import torch.nn as nn

from marl_range_flocking_b200.policies import BatchedActors, BatchedQNet


class _Actor(nn.Module):                                   # layer names = the reference's state_dict keys
    def __init__(self, i, f1, f2, a):
        super().__init__()
        self.fc1, self.bn1 = nn.Linear(i, f1), nn.LayerNorm(f1)
        self.fc2, self.bn2 = nn.Linear(f1, f2), nn.LayerNorm(f2)
        self.mu = nn.Linear(f2, a)

    def forward(self, s):
        x = torch.relu(self.bn1(self.fc1(s)))
        x = torch.relu(self.bn2(self.fc2(x)))
        return torch.tanh(self.mu(x))


def test_batched_actors_match_per_agent_modules():
    torch.manual_seed(0)
    N, E, k = 5, 7, 3
    actors = [_Actor(4 * k, 40, 30, 2) for _ in range(N)]
    for a in actors:                                        # non-trivial LayerNorm affine parameters
        nn.init.normal_(a.bn1.weight, 1, 0.1); nn.init.normal_(a.bn1.bias, 0, 0.1)
    batched = BatchedActors.from_state_dicts([a.state_dict() for a in actors])
    obs = torch.rand(E, N, 4, k) * 7
    want = torch.stack([torch.stack([actors[i](obs[e, i].reshape(-1)) for i in range(N)]) for e in range(E)])
    got = batched(obs)
    assert got.shape == (E, N, 2)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    assert BatchedActors(N, 12).forward(obs).abs().max() <= 1.0


class _QNet(nn.Module):
    def __init__(self, N, n_obs, A, recurrent):
        super().__init__()
        self.N, self.recurrent = N, recurrent
        for i in range(N):
            setattr(self, f"agent_feature_{i}", nn.Sequential(nn.Linear(n_obs, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU()))
            if recurrent:
                setattr(self, f"agent_gru_{i}", nn.GRUCell(32, 32))
            setattr(self, f"agent_q_{i}", nn.Linear(32, A))

    def forward(self, obs, hidden):
        qs, hs = [], []
        for i in range(self.N):
            x = getattr(self, f"agent_feature_{i}")(obs[:, i, :])
            if self.recurrent:
                x = getattr(self, f"agent_gru_{i}")(x, hidden[:, i, :])
                hs.append(x.unsqueeze(1))
            qs.append(getattr(self, f"agent_q_{i}")(x).unsqueeze(1))
        return torch.cat(qs, 1), (torch.cat(hs, 1) if hs else None)


def test_batched_qnet_matches_per_agent_modules_and_samples_actions():
    torch.manual_seed(1)
    N, E, k, A = 6, 9, 4, 4
    for rec in (False, True):
        ref = _QNet(N, k, A, rec)
        net = BatchedQNet.from_state_dict(ref.state_dict(), N)
        obs, hid = torch.rand(E, N, k) * 7, torch.randn(E, N, 32)
        q_ref, h_ref = ref(obs, hid)
        q, h = net(obs, hid)
        assert torch.allclose(q, q_ref, atol=1e-5, rtol=1e-5)
        if rec:
            assert torch.allclose(h, h_ref, atol=1e-5, rtol=1e-5)
        greedy, _ = net.sample_action(obs, hid, epsilon=-1.0)
        assert torch.equal(greedy, q_ref.argmax(2).float()) and greedy.dtype == torch.float32
        rnd, _ = net.sample_action(obs, hid, epsilon=2.0)
        assert rnd.shape == (E, N) and bool(((rnd >= 0) & (rnd < A)).all())


def test_device_replay_stores_slices_and_samples_chunks():
    from marl_range_flocking_b200.replay import DeviceReplay
    E, N, k, A, T = 4, 3, 2, 2, 16
    rb = DeviceReplay(E, N, k, A, capacity_steps=T, chunk_size=5)
    for t in range(12):
        base = torch.full((E, N, k), float(t)) + torch.arange(E)[:, None, None] * 100
        rb.add(dict(obs=base, next_obs=base + 1, actions=torch.full((E, N, A), float(t)), reward=torch.full((E, N, 1), -float(t)),
                    agent_done=torch.zeros(E, N, dtype=torch.bool), episode_end=torch.zeros(E, dtype=torch.bool)))
    assert len(rb) == 12 * E
    s, r, ns, d, a_s, a_ns, a_a = rb.get_minibatch(batch_size=8, generator=torch.Generator().manual_seed(0))
    assert s.shape == (8, 5, N, k) and r.shape == (8, 5, N, 1) and d.shape == (8, 5, N, 1)
    assert a_s.shape == (N, 8, 5, k) and a_a.shape == (N, 8, 5, A)
    # every chunk is 5 consecutive steps of ONE env, and the fields stay aligned
    step = s[..., 0, 0] % 100
    env = (s[..., 0, 0] // 100)
    assert bool((step[:, 1:] - step[:, :-1] == 1).all()) and bool((env == env[:, :1]).all())
    assert torch.equal(ns, s + 1) and torch.equal(r[..., 0, 0], -step) and torch.equal(a_a[0, ..., 0], step)
    assert torch.equal(a_s, s.permute(2, 0, 1, 3))


def test_device_replay_ring_never_straddles_the_write_head():
    from marl_range_flocking_b200.replay import DeviceReplay
    E, N, T, C = 2, 2, 8, 3
    rb = DeviceReplay(E, N, 1, 1, capacity_steps=T, chunk_size=C)
    for t in range(21):                                  # wraps twice; slices 13..20 are alive
        x = torch.full((E, N, 1), float(t))
        rb.add(dict(obs=x, next_obs=x, actions=x, reward=x, agent_done=torch.zeros(E, N, dtype=torch.bool),
                    episode_end=torch.zeros(E, dtype=torch.bool)))
    g = torch.Generator().manual_seed(1)
    for _ in range(20):
        s = rb.get_minibatch(16, generator=g)[0][..., 0, 0]                     # (B, C) step ids
        assert bool((s[:, 1:] - s[:, :-1] == 1).all()) and float(s.min()) >= 13 and float(s.max()) <= 20


def test_fused_policy_paths_have_no_cpu_fallback():
    """The fused kernels (flock_actor_forward, flock_qnet_forward) are CUDA only: with CPU parameters the fused
    entry points raise instead of silently running the PyTorch path."""
    import pytest
    from marl_range_flocking_b200.policies import BatchedActors, BatchedQNet
    a = BatchedActors(2, 12, 400, 300, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        a.forward_fused(torch.zeros(4, 2, 12))
    q = BatchedQNet(2, 4, 4, recurrent=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        q.sample_action_fused(torch.zeros(4, 2, 4), q_hidden := torch.zeros(4, 2, 32), 0.1)


def test_actor_and_qnet_entry_points_validate_arguments_without_a_gpu():
    """Argument validation of the new C entry points happens before any CUDA call."""
    import ctypes
    from marl_range_flocking_b200 import _lib
    lib = _lib.load_library()
    assert lib.flock_actor_packed_bytes(3) == 3 * lib.flock_actor_packed_bytes(1) > 0
    null10 = (ctypes.c_void_p * 10)()
    assert lib.flock_actor_pack(2, 12, 256, 128, 2, null10, None, None) == -1          # FLOCK_E_INVALID: dims
    assert b"400-300-2" in lib.flock_last_error()
    assert lib.flock_actor_pack(2, 15, 400, 300, 2, null10, None, None) == -1          # input width
    assert lib.flock_actor_forward(None, None, None, 128, 2, 12, None) == -1
    assert lib.flock_qnet_forward(null10, 1, None, None, None, None, None, 8, 2, 4, 4, 0.1, 0, 0, 0, None, None) == -1
    assert lib.flock_wait_host(None) == -1
    # tensor-core GRU front ends (flock_gru_tc.cu)
    assert lib.flock_gru_tc_packed_bytes(0, 5) == 5 * lib.flock_gru_tc_packed_bytes(0, 1) > 0
    assert lib.flock_gru_tc_packed_bytes(1, 5) > lib.flock_gru_tc_packed_bytes(0, 5)         # the VDN image also holds W2 and the head
    assert lib.flock_gru_tc_packed_bytes(2, 5) == 0 and lib.flock_gru_tc_packed_bytes(1, 0) == 0
    assert lib.flock_gru_tc_pack(2, 4, 4, 4, null10, None, None) == -1 and b"mode" in lib.flock_last_error()
    assert lib.flock_gru_tc_pack(1, 4, 17, 4, null10, None, None) == -1 and b"n_obs" in lib.flock_last_error()
    assert lib.flock_gru_tc_pack(1, 4, 4, 17, null10, None, None) == -1 and b"n_actions" in lib.flock_last_error()
    assert lib.flock_gru_tc_pack(0, 4, 4, 0, null10, None, None) == -1                       # NULL parameters / image
    assert lib.flock_qnet_forward_tc(None, None, None, None, None, None, 8, 2, 4, 4, 0.1, 0, 0, 0, None, None) == -1
    assert lib.flock_rnn_actor_forward_tc(None, None, None, None, None, None, 8, 2, 4, None, 0.15, 0.0, 0.2, 0.01, 0, 0, 0,
                                          None, None) == -1


class _RnnActor(nn.Module):                                # layer names = the reference's state_dict keys (net.py:32-38)
    def __init__(self, i, a=2, h1=40, h2=30, hr=32):
        super().__init__()
        self.fce, self.gru = nn.Linear(i, hr), nn.GRUCell(hr, hr)
        self.fc1, self.fc2 = nn.Linear(hr, h1), nn.Linear(h1, h2)
        self.linear_speed, self.angular_speed = nn.Linear(h2, a // 2), nn.Linear(h2, a // 2)

    def forward(self, x, hidden):                          # net.py:53-72
        out = self.gru(self.fce(x), hidden)
        nxt = out.clone()
        out = torch.relu(self.fc2(torch.relu(self.fc1(out))))
        lin = (torch.tanh(self.linear_speed(out)) + 1) / 2
        ang = torch.tanh(self.angular_speed(out)) * 1.5
        return torch.cat([lin, ang], dim=1), nxt


def test_batched_rnn_actors_match_per_agent_modules():
    from marl_range_flocking_b200.policies import BatchedRnnActors
    torch.manual_seed(3)
    N, E, k = 4, 6, 4
    actors = [_RnnActor(k) for _ in range(N)]
    batched = BatchedRnnActors.from_state_dicts([a.state_dict() for a in actors])
    obs, hid = torch.rand(E, N, k) * 14, torch.randn(E, N, 32) * 0.5
    want_a = torch.stack([actors[i](obs[:, i], hid[:, i])[0] for i in range(N)], dim=1)
    want_h = torch.stack([actors[i](obs[:, i], hid[:, i])[1] for i in range(N)], dim=1)
    got_a, got_h = batched(obs, hid)
    assert got_a.shape == (E, N, 2) and got_h.shape == (E, N, 32)
    assert torch.allclose(got_a, want_a, atol=1e-5, rtol=1e-5) and torch.allclose(got_h, want_h, atol=1e-5, rtol=1e-5)
    assert bool(((got_a[..., 0] >= 0) & (got_a[..., 0] <= 1)).all()) and bool((got_a[..., 1].abs() <= 1.5).all())
    assert batched.init_hidden(E).shape == (E, N, 32)
