"""Generate the golden vectors that pin policies.py / replay.py to the reference's OWN learner classes.

Runs the UNMODIFIED reference modules on CPU under oracle/ref_shim.py:
  learners/maddpg_shared_critic/ddpg_network.py:85-141  ActorNetwork  (shared-critic DDPG actor)
  learners/maddpg_official_rnn/net.py:14-72             Actor         (recurrent MADDPG actor)
  learners/vdn/net.py:11-58                             QNet          (VDN per-agent Q networks)
  learners/maddpg_official_rnn/memory_rnn.py:8-103      ReplayBufferMaddpg
  learners/vdn/utils.py:7-69                            ReplayBufferVDN
Network weights are drawn from a numpy generator with a recorded seed and loaded into the reference modules, so
the fixture only has to carry inputs and the reference's outputs (tests rebuild the same weights from the seed).
Only runnable in the build container; tests/golden/policy_golden.npz is committed.

    python tests/golden/make_policy_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def seeded_state_dict(module, seed, scale=None):
    """Deterministic weights: every parameter ~ U(-b, b) with b = 1/sqrt(fan_in) (or `scale`), LayerNorm affine
    parameters non-trivial; drawn parameter by parameter, in state_dict order, from numpy's PCG64(seed)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, t in module.state_dict().items():
        shape = tuple(t.shape)
        if name.startswith("bn") and name.endswith("weight"):
            a = rng.uniform(0.5, 1.5, shape)
        elif name.startswith("bn"):
            a = rng.uniform(-0.3, 0.3, shape)
        else:
            b = scale.get(name.split(".")[0]) if scale and name.split(".")[0] in scale else shape[-1] ** -0.5 if len(shape) > 1 else 0.1
            a = rng.uniform(-b, b, shape)
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


class _Space:
    def __init__(self, n):
        self.shape, self.n = (n,), n


def main():
    out = {}
    rng = np.random.default_rng(123)
    # ---- ActorNetwork: N = 3 agents, input 12 (uw: 4 x k=3), 400-300-2 ----
    m = ref_shim.load_learner_module("learners/maddpg_shared_critic/ddpg_network.py")
    N, E = 3, 64
    obs = (rng.uniform(0, 7, (E, N, 12))).astype(np.float32)
    outs = []
    with ref_shim.scratch_cwd(m):
        for i in range(N):
            net = m.ActorNetwork(1e-3, (12,), 400, 300, 2, f"a{i}")
            net.load_state_dict(seeded_state_dict(net, 1000 + i, scale={"mu": 0.1}))
            with torch.no_grad():
                outs.append(net(torch.from_numpy(obs[:, i])).numpy())
    out["actor_obs"], out["actor_out"] = obs, np.stack(outs, axis=1)          # (E, N, 2)
    out["actor_seeds"] = np.arange(1000, 1000 + N)
    # ---- recurrent Actor: N = 3 agents, input 4 (v2 k=4), two consecutive steps ----
    m = ref_shim.load_learner_module("learners/maddpg_official_rnn/net.py")
    obs = rng.uniform(0, 14, (2, E, N, 4)).astype(np.float32)
    acts, hids = [], []
    with ref_shim.scratch_cwd(m):
        for i in range(N):
            net = m.Actor(4, 2, name=f"r{i}")
            net.load_state_dict(seeded_state_dict(net, 2000 + i, scale={"linear_speed": 0.1, "angular_speed": 0.1}))
            h = net.init_hidden(E)
            a_t, h_t = [], []
            with torch.no_grad():
                for t in range(2):
                    a, h = net(torch.from_numpy(obs[t, :, i]), h)
                    a_t.append(a.numpy()); h_t.append(h.numpy())
            acts.append(np.stack(a_t)); hids.append(np.stack(h_t))
    out["rnn_obs"], out["rnn_act"], out["rnn_hidden"] = obs, np.stack(acts, axis=2), np.stack(hids, axis=2)   # (2,E,N,.)
    out["rnn_seeds"] = np.arange(2000, 2000 + N)
    # ---- QNet (recurrent and not): N = 4 agents, 4 obs, 4 actions ----
    m = ref_shim.load_learner_module("learners/vdn/net.py")
    N = 4
    obs = rng.uniform(0, 7, (E, N, 4)).astype(np.float32)
    hid = (rng.standard_normal((E, N, 32)) * 0.3).astype(np.float32)
    for rec in (False, True):
        q = m.QNet([_Space(4)] * N, [_Space(4)] * N, rec)
        q.load_state_dict(seeded_state_dict(q, 3000 + int(rec)))
        with torch.no_grad():
            qv, nh = q(torch.from_numpy(obs), torch.from_numpy(hid))
        tag = "rec" if rec else "ff"
        out[f"qnet_{tag}_q"] = qv.numpy()
        if rec:
            out["qnet_rec_hidden"] = nh.numpy()
    out["qnet_obs"], out["qnet_hid"] = obs, hid
    # ---- replay buffers: one env, T transitions, fixed chunk starts ----
    T, N, k, A = 40, 3, 4, 2
    tr = dict(obs=rng.uniform(0, 14, (T + 1, N, k)).astype(np.float32), act=rng.uniform(-1, 1, (T, N, A)).astype(np.float32),
              rew=rng.uniform(-5, 1, (T, N, 1)).astype(np.float32), done=(rng.uniform(0, 1, (T, N)) < 0.1),
              env_done=(rng.uniform(0, 1, (T,)) < 0.15))
    m = ref_shim.load_learner_module("learners/maddpg_official_rnn/memory_rnn.py")

    class _Env:
        num_particles = N
        observation_space = [type("B", (), {"shape": (N, k)})(), [_Space(k)] * N]
        action_space = [_Space(A)] * N
    rb = m.ReplayBufferMaddpg(_Env(), buffer_capacity=64, batch_size=5, chunk_size=6, min_size_buffer=1)
    for t in range(T):
        o, o2 = torch.from_numpy(tr["obs"][t]), torch.from_numpy(tr["obs"][t + 1])
        rb.add_record(o, o2, torch.from_numpy(tr["act"][t]), o, o2, torch.from_numpy(tr["rew"][t]),
                      torch.from_numpy(tr["done"][t]))
    np.random.seed(7)
    starts = np.random.choice(min(rb.buffer_counter, rb.buffer_capacity) - rb.chunk_size, rb.batch_size, replace=False)
    np.random.seed(7)
    mb = rb.get_minibatch()
    for i, name in enumerate(("states", "rewards", "next_states", "dones", "a_states", "a_next_states", "a_actions")):
        out["maddpg_mb_" + name] = mb[i].numpy()
    out["maddpg_starts"] = starts
    m = ref_shim.load_learner_module("learners/vdn/utils.py")
    vb = m.ReplayBufferVDN(64, chunk_size=6, n_agents=N, input_shape=[k], batch_size=5)
    for t in range(T):
        vb.put((torch.from_numpy(tr["obs"][t]), torch.from_numpy(tr["act"][t, :, 0]), torch.from_numpy(tr["rew"][t]),
                torch.from_numpy(tr["obs"][t + 1]), [int(tr["env_done"][t])]))
    np.random.seed(9)
    vstarts = np.random.randint(0, vb.size() - 6, 5)
    np.random.seed(9)
    vc = vb.sample_chunk(5, 6)
    for i, name in enumerate(("state", "action", "reward", "new_state", "terminal")):
        out["vdn_chunk_" + name] = vc[i].numpy()
    out["vdn_starts"] = vstarts
    for key, v in tr.items():
        out["replay_" + key] = v
    np.savez_compressed(os.path.join(OUT, "policy_golden.npz"), **out)
    print("policy_golden.npz", {k_: v_.shape for k_, v_ in out.items()})


if __name__ == "__main__":
    main()
