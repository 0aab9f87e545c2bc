"""Pin the batched policies (policies.py) and the replay buffers (replay.py) to the reference's OWN classes.

* fixture tests (always run): tests/golden/policy_golden.npz holds inputs and the outputs of the unmodified
  reference modules (ActorNetwork, Actor, QNet, ReplayBufferMaddpg, ReplayBufferVDN) recorded by
  tests/golden/make_policy_golden.py; the batched modules, loaded with the same seeded weights through
  `from_state_dict(s)`, must reproduce them (fp32, 1e-5) and the replay buffers must return the same
  minibatches for the same chunk starts, exactly.
* live tests (skipped when /root/reference is absent): the same against freshly constructed reference modules
  with their own default initialisation, plus the state_dict layouts this test suite restates.
* GPU tests: the fused kernels against the recorded REFERENCE outputs (not against policies.py).
"""
import numpy as np
import pytest
import torch

from tests.policy_golden_util import (actor_layout, build_policies, load_golden, qnet_layout, rnn_actor_layout)

try:
    from oracle import ref_shim
    HAVE_REF = ref_shim.reference_available()
except Exception:        # pragma: no cover
    HAVE_REF = False
live = pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree (build container only)")


def test_batched_policies_reproduce_the_recorded_reference_outputs():
    g = load_golden()
    actors, rnn, qff, qrec = build_policies(g)
    with torch.no_grad():
        out = actors(torch.from_numpy(g["actor_obs"]))
        assert torch.allclose(out, torch.from_numpy(g["actor_out"]), atol=2e-6, rtol=1e-5)
        h = rnn.init_hidden(g["rnn_obs"].shape[1])
        for t in range(2):
            a, h = rnn(torch.from_numpy(g["rnn_obs"][t]), h)
            assert torch.allclose(a, torch.from_numpy(g["rnn_act"][t]), atol=2e-6, rtol=1e-5), t
            assert torch.allclose(h, torch.from_numpy(g["rnn_hidden"][t]), atol=2e-6, rtol=1e-5), t
        obs, hid = torch.from_numpy(g["qnet_obs"]), torch.from_numpy(g["qnet_hid"])
        q, _ = qff(obs, hid)
        assert torch.allclose(q, torch.from_numpy(g["qnet_ff_q"]), atol=2e-6, rtol=1e-5)
        q, nh = qrec(obs, hid)
        assert torch.allclose(q, torch.from_numpy(g["qnet_rec_q"]), atol=2e-6, rtol=1e-5)
        assert torch.allclose(nh, torch.from_numpy(g["qnet_rec_hidden"]), atol=2e-6, rtol=1e-5)
        # greedy action selection = the reference's argmax branch (net.py:57)
        act, _ = qrec.sample_action(obs, hid, 0.0)
        assert torch.equal(act, torch.from_numpy(g["qnet_rec_q"]).argmax(dim=2).float())


def _fill(rb_add, g):
    T = g["replay_act"].shape[0]
    for t in range(T):
        rb_add(dict(obs=torch.from_numpy(g["replay_obs"][t])[None], next_obs=torch.from_numpy(g["replay_obs"][t + 1])[None],
                    actions=torch.from_numpy(g["replay_act"][t])[None], reward=torch.from_numpy(g["replay_rew"][t])[None],
                    agent_done=torch.from_numpy(g["replay_done"][t])[None], env_done=torch.from_numpy(g["replay_env_done"][t:t + 1]),
                    episode_end=torch.from_numpy(g["replay_env_done"][t:t + 1])))


def test_device_replay_returns_the_reference_minibatches():
    """E = 1: DeviceReplay holds what ReplayBufferMaddpg / ReplayBufferVDN hold; for the same chunk starts the
    seven (memory_rnn.py:95-101) resp. five (vdn/utils.py:54-58) tensors are identical."""
    from marl_range_flocking_b200.replay import DeviceReplay
    g = load_golden()
    rb = DeviceReplay(1, 3, 4, 2, capacity_steps=64, chunk_size=6)
    _fill(rb.add, g)
    mb = rb.get_minibatch(5, starts=g["maddpg_starts"], envs=np.zeros(5, np.int64))
    for got, name in zip(mb, ("states", "rewards", "next_states", "dones", "a_states", "a_next_states", "a_actions")):
        want = torch.from_numpy(g["maddpg_mb_" + name])
        assert got.shape == want.shape and torch.equal(got, want), name
    vc = rb.sample_chunk(5, 6, starts=g["vdn_starts"], envs=np.zeros(5, np.int64))
    # the VDN buffer stores only the first action column (ids); reward[:, 0] (vdn/utils.py:42)
    for got, name in zip(vc, ("state", "action", "reward", "new_state", "terminal")):
        want = torch.from_numpy(g["vdn_chunk_" + name])
        assert got.shape == want.shape and torch.equal(got, want), name


def test_trajectory_replay_matches_device_replay_on_the_same_stream():
    """TrajectoryReplay (the zero-copy target of flock_rollout_n) keeps one stream of range rows; for v2-like
    data (window 1) its gathers must equal DeviceReplay's on the same transitions -- hence the reference's."""
    from marl_range_flocking_b200.replay import DeviceReplay, TrajectoryReplay
    g = load_golden()
    T = g["replay_act"].shape[0]
    rb = DeviceReplay(1, 3, 4, 2, capacity_steps=T, chunk_size=6)
    _fill(rb.add, g)
    tr = TrajectoryReplay(1, 3, 4, 2, capacity_steps=T, chunk_size=6)
    tr.begin(torch.from_numpy(g["replay_obs"][0])[None])
    v = tr.next_views(T)                       # what VecEnv.rollout_n would write into
    v.obs.copy_(torch.from_numpy(g["replay_obs"][1:])[:, None])
    v.reward.copy_(torch.from_numpy(g["replay_rew"])[:, None])
    v.agent_done.copy_(torch.from_numpy(g["replay_done"])[:, None])
    v.env_done.copy_(torch.from_numpy(g["replay_env_done"])[:, None])
    tr.commit(T, torch.from_numpy(g["replay_act"])[:, None])
    starts, envs = g["maddpg_starts"], np.zeros(5, np.int64)
    for a, b in zip(rb.get_minibatch(5, starts=starts, envs=envs), tr.get_minibatch(5, starts=starts, envs=envs)):
        assert torch.equal(a, b)
    for a, b in zip(rb.sample_chunk(5, 6, starts=starts, envs=envs), tr.sample_chunk(5, 6, starts=starts, envs=envs)):
        assert torch.equal(a, b)


def test_trajectory_replay_rebuilds_the_uw_window_from_the_row_stream():
    """uw: the (4, k) newest-first window of step t is rows t, t-1, t-2, t-3 of the stream, zero before the episode's
    first observation (gym_flock_uw.py:100-102,120-123). Checked against an explicit roll-and-insert simulation."""
    from marl_range_flocking_b200.replay import TrajectoryReplay
    rng = np.random.default_rng(0)
    T, E, N, k = 24, 3, 2, 3
    rows = rng.uniform(0, 7, (T + 1, E, N, k)).astype(np.float32)
    end = rng.uniform(0, 1, (T, E)) < 0.2
    tr = TrajectoryReplay(E, N, k, 2, capacity_steps=T, chunk_size=5, window=4)
    tr.begin(torch.from_numpy(rows[0]))
    v = tr.next_views(T)
    v.obs.copy_(torch.from_numpy(rows[1:]))
    v.env_done.copy_(torch.from_numpy(end))
    tr.commit(T, torch.zeros(T, E, N, 2))
    # reference semantics: memory rolled by one and the newest row inserted; reset zero-fills then inserts
    win = np.zeros((T + 1, E, N, 4, k), np.float32)
    win[0, :, :, 0] = rows[0]
    for t in range(T):
        for e in range(E):
            if end[t, e]:
                win[t + 1, e] = 0
            else:
                win[t + 1, e, :, 1:] = win[t, e, :, :3]
            win[t + 1, e, :, 0] = rows[t + 1, e]
    starts, envs = np.array([0, 3, 11, 19]), np.array([0, 1, 2, 1])
    s, a, r, s2, term = tr.sample_chunk(4, 5, starts=starts, envs=envs)
    for b in range(4):
        for c in range(5):
            t, e = starts[b] + c, envs[b]
            assert np.array_equal(s[b, c].numpy(), win[t, e].reshape(N, 4 * k)), (b, c)
            assert np.array_equal(s2[b, c].numpy(), win[t + 1, e].reshape(N, 4 * k)), (b, c)
            assert float(term[b, c, 0]) == float(end[t, e])


def test_stats_logger_writes_the_reference_tags():
    """tb_bridge.StatsLogger: the reference's per-game scalars (main.py:61-65) fed from the device statistics."""
    from marl_range_flocking_b200.tb_bridge import StatsLogger

    class Env:
        num_particles = 10

        def __init__(self):
            self.s = torch.zeros(8, dtype=torch.int64)

        def stats_tensor(self):
            return self.s

    class Writer:
        def __init__(self):
            self.rows = []

        def add_scalar(self, tag, value, step):
            self.rows.append((tag, value, step))

    env, w = Env(), Writer()
    log = StatsLogger(w, env)
    env.s[:3] = torch.tensor([4, 100, int(-2.5 * 10 * 4 * 2**32)])       # 4 episodes, mean return -2.5 per agent
    out = log.log(0, extra={"Buffer size": 1234})
    assert abs(out["Train/Score"] + 2.5) < 1e-9 and out["Train/Mean episode length"] == 25.0 and out["Train/Buffer size"] == 1234.0
    env.s[:3] += torch.tensor([2, 10, int(1.0 * 10 * 2 * 2**32)])
    out = log.log(1)
    assert abs(out["Train/Score"] - 1.0) < 1e-9 and abs(out["Train/Average score"] - (-0.75)) < 1e-9
    tags = {t for t, _, _ in w.rows}
    assert {"Train/Score", "Train/Average score", "Train/Time taken", "Train/Buffer size"} <= tags


# ---------------------------------------------------------------------------------------------------
# live reference (build container only)
# ---------------------------------------------------------------------------------------------------
class _Space:
    def __init__(self, n):
        self.shape, self.n = (n,), n


@live
def test_restated_state_dict_layouts_match_the_reference_modules():
    m = ref_shim.load_learner_module("learners/maddpg_shared_critic/ddpg_network.py")
    with ref_shim.scratch_cwd(m):
        net = m.ActorNetwork(1e-3, (12,), 400, 300, 2, "a")
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == actor_layout()
    m = ref_shim.load_learner_module("learners/maddpg_official_rnn/net.py")
    with ref_shim.scratch_cwd(m):
        net = m.Actor(4, 2)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == rnn_actor_layout()
    m = ref_shim.load_learner_module("learners/vdn/net.py")
    for rec in (False, True):
        q = m.QNet([_Space(4)] * 4, [_Space(4)] * 4, rec)
        assert [(k, tuple(v.shape)) for k, v in q.state_dict().items()] == qnet_layout(recurrent=rec)


@live
def test_batched_policies_equal_live_reference_modules_with_default_init():
    from marl_range_flocking_b200.policies import BatchedActors, BatchedQNet, BatchedRnnActors
    torch.manual_seed(0)
    E, N = 33, 5
    m = ref_shim.load_learner_module("learners/maddpg_shared_critic/ddpg_network.py")
    with ref_shim.scratch_cwd(m):
        nets = [m.ActorNetwork(1e-3, (12,), 400, 300, 2, f"a{i}") for i in range(N)]
    obs = torch.rand(E, N, 4, 3) * 7
    with torch.no_grad():
        want = torch.stack([nets[i](obs[:, i].reshape(E, -1)) for i in range(N)], dim=1)     # train_flock.py:114-115
        got = BatchedActors.from_state_dicts([n.state_dict() for n in nets])(obs)
    assert torch.allclose(got, want, atol=2e-6, rtol=1e-5)
    m = ref_shim.load_learner_module("learners/maddpg_official_rnn/net.py")
    with ref_shim.scratch_cwd(m):
        nets = [m.Actor(4, 2, name=f"r{i}") for i in range(N)]
    obs, hid = torch.rand(E, N, 4) * 14, torch.randn(E, N, 32) * 0.3
    with torch.no_grad():
        outs = [nets[i](obs[:, i], hid[:, i]) for i in range(N)]                                # MADDPG.py:24-33
        got_a, got_h = BatchedRnnActors.from_state_dicts([n.state_dict() for n in nets])(obs, hid)
    assert torch.allclose(got_a, torch.stack([o[0] for o in outs], dim=1), atol=2e-6, rtol=1e-5)
    assert torch.allclose(got_h, torch.stack([o[1] for o in outs], dim=1), atol=2e-6, rtol=1e-5)
    m = ref_shim.load_learner_module("learners/vdn/net.py")
    for rec in (False, True):
        q = m.QNet([_Space(4)] * N, [_Space(4)] * N, rec)
        obs, hid = torch.rand(E, N, 4) * 7, torch.randn(E, N, 32) * 0.3
        with torch.no_grad():
            wq, wh = q(obs, hid)
            gq, gh = BatchedQNet.from_state_dict(q.state_dict(), N)(obs, hid)
        assert torch.allclose(gq, wq, atol=2e-6, rtol=1e-5)
        if rec:
            assert torch.allclose(gh, wh, atol=2e-6, rtol=1e-5)


@live
def test_policy_fixture_regenerates_bit_identically(tmp_path, monkeypatch):
    """The committed fixture IS what the unmodified reference produces (same check the env goldens get)."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("_make_policy_golden", os.path.join(here, "golden", "make_policy_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "OUT", str(tmp_path))
    mod.main()
    new = np.load(os.path.join(str(tmp_path), "policy_golden.npz"))
    old = load_golden()
    assert set(new.files) == set(old)
    for k in new.files:
        assert np.array_equal(new[k], old[k]), k


# ---------------------------------------------------------------------------------------------------
# GPU: the fused kernels against the recorded reference outputs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_fused_kernels_against_the_recorded_reference_outputs():
    g = load_golden()
    actors, rnn, qff, qrec = build_policies(g, device="cuda")
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(a).to(dev)
    # shared-critic actor: bf16 tensor-core operands -> 4e-2 absolute on a tanh output
    out = actors.forward_fused(t(g["actor_obs"]))
    assert (out - t(g["actor_out"])).abs().max().item() <= 4e-2
    # recurrent actor: fp32 GRU front end (hidden state 2e-5), bf16 MLP (actions 4e-2)
    h = rnn.init_hidden(g["rnn_obs"].shape[1])
    for step in range(2):
        a, h = rnn.forward_fused(t(g["rnn_obs"][step]), h)
        assert (h - t(g["rnn_hidden"][step])).abs().max().item() <= 2e-5, step
        assert (a - t(g["rnn_act"][step])).abs().max().item() <= 4e-2, step
    # VDN Q networks: fp32 kernel, 5e-5; greedy actions equal to the reference's argmax wherever the gap is clear
    obs, hid = t(g["qnet_obs"]), t(g["qnet_hid"])
    q, _ = qff.forward_fused(obs, hid)
    assert torch.allclose(q, t(g["qnet_ff_q"]), atol=5e-5, rtol=5e-5)
    q, nh = qrec.forward_fused(obs, hid)
    assert torch.allclose(q, t(g["qnet_rec_q"]), atol=5e-5, rtol=5e-5) and torch.allclose(nh, t(g["qnet_rec_hidden"]), atol=5e-5, rtol=5e-5)
    act, _ = qrec.sample_action_fused(obs, hid, 0.0)
    ref_q = t(g["qnet_rec_q"])
    top2 = ref_q.topk(2, dim=2).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-4
    assert torch.equal(act[clear], ref_q.argmax(dim=2).float()[clear])
