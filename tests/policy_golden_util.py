"""Rebuild the seeded weights of tests/golden/make_policy_golden.py without the reference tree: the state_dict
layouts (key order and shapes) of the reference's learner networks, restated. `test_reference_pinning.py` checks
these layouts against the live reference modules whenever /root/reference is present."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_golden.npz")


def actor_layout(in_dims=12, fc1=400, fc2=300, na=2):
    # ActorNetwork, learners/maddpg_shared_critic/ddpg_network.py:104-123
    return [("fc1.weight", (fc1, in_dims)), ("fc1.bias", (fc1,)), ("bn1.weight", (fc1,)), ("bn1.bias", (fc1,)),
            ("fc2.weight", (fc2, fc1)), ("fc2.bias", (fc2,)), ("bn2.weight", (fc2,)), ("bn2.bias", (fc2,)),
            ("mu.weight", (na, fc2)), ("mu.bias", (na,))]


def rnn_actor_layout(in_dims=4, h1=400, h2=300, hr=32):
    # Actor, learners/maddpg_official_rnn/net.py:32-38
    return [("fce.weight", (hr, in_dims)), ("fce.bias", (hr,)), ("gru.weight_ih", (3 * hr, hr)), ("gru.weight_hh", (3 * hr, hr)),
            ("gru.bias_ih", (3 * hr,)), ("gru.bias_hh", (3 * hr,)), ("fc1.weight", (h1, hr)), ("fc1.bias", (h1,)),
            ("fc2.weight", (h2, h1)), ("fc2.bias", (h2,)), ("linear_speed.weight", (1, h2)), ("linear_speed.bias", (1,)),
            ("angular_speed.weight", (1, h2)), ("angular_speed.bias", (1,))]


def qnet_layout(num_agents=4, n_obs=4, n_act=4, recurrent=False, hx=32):
    # QNet, learners/vdn/net.py:17-25 (per agent: feature Sequential, [GRUCell], q head)
    out = []
    for i in range(num_agents):
        out += [(f"agent_feature_{i}.0.weight", (64, n_obs)), (f"agent_feature_{i}.0.bias", (64,)),
                (f"agent_feature_{i}.2.weight", (hx, 64)), (f"agent_feature_{i}.2.bias", (hx,))]
        if recurrent:
            out += [(f"agent_gru_{i}.weight_ih", (3 * hx, hx)), (f"agent_gru_{i}.weight_hh", (3 * hx, hx)),
                    (f"agent_gru_{i}.bias_ih", (3 * hx,)), (f"agent_gru_{i}.bias_hh", (3 * hx,))]
        out += [(f"agent_q_{i}.weight", (n_act, hx)), (f"agent_q_{i}.bias", (n_act,))]
    return out


def seeded_state_dict(layout, seed, scale=None):
    """Same draws as make_policy_golden.seeded_state_dict, from the restated layout."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in layout:
        if name.startswith("bn") and name.endswith("weight"):
            a = rng.uniform(0.5, 1.5, shape)
        elif name.startswith("bn"):
            a = rng.uniform(-0.3, 0.3, shape)
        else:
            head = name.split(".")[0]
            b = scale[head] if scale and head in scale else (shape[-1] ** -0.5 if len(shape) > 1 else 0.1)
            a = rng.uniform(-b, b, shape)
        sd[name] = torch.from_numpy(a.astype(np.float32))
    return sd


def load_golden():
    z = np.load(GOLDEN, allow_pickle=False)
    return {k: z[k] for k in z.files}


def build_policies(g, device="cpu"):
    """The three batched policy modules of policies.py loaded with the fixture's seeded reference weights."""
    from marl_range_flocking_b200.policies import BatchedActors, BatchedQNet, BatchedRnnActors
    actors = BatchedActors.from_state_dicts(
        [seeded_state_dict(actor_layout(), int(s), scale={"mu": 0.1}) for s in g["actor_seeds"]], device=device)
    rnn = BatchedRnnActors.from_state_dicts(
        [seeded_state_dict(rnn_actor_layout(), int(s), scale={"linear_speed": 0.1, "angular_speed": 0.1}) for s in g["rnn_seeds"]],
        device=device)
    qff = BatchedQNet.from_state_dict(seeded_state_dict(qnet_layout(recurrent=False), 3000), 4, device=device)
    qrec = BatchedQNet.from_state_dict(seeded_state_dict(qnet_layout(recurrent=True), 3001), 4, device=device)
    return actors, rnn, qff, qrec
