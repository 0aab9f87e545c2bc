"""Batched policy inference next to the env step ("next" scope, plain PyTorch).

The reference evaluates one small network per agent in a Python loop
(`learners/maddpg_shared_critic/train_flock.py:114-115`, `learners/vdn/net.py:27-37`). With E envs
on the device that loop becomes one batched matrix multiply per layer over the agent dimension:
weights of the N per-agent networks are stacked to `(N, in, out)` and applied to observations laid
out `(N, E, in)` with `torch.baddbmm` (cuBLAS; library GEMMs are fine here -- this is not the
env hot path). Parameters load from the reference's own `state_dict`s.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def _stack(sds: Sequence[Dict[str, torch.Tensor]], key: str) -> torch.Tensor:
    return torch.stack([sd[key] for sd in sds])


class BatchedActors(torch.nn.Module):
    """N per-agent DDPG actors (`ActorNetwork`, learners/maddpg_shared_critic/ddpg_network.py:85-141):
    fc1 -> LayerNorm -> ReLU -> fc2 -> LayerNorm -> ReLU -> mu -> tanh, evaluated for all envs and
    agents at once. `forward(obs)` takes `(E, N, ...)` observations (trailing dims are flattened, as
    `obs.reshape(N, -1)` does upstream, train_flock.py:115) and returns `(E, N, n_actions)`."""

    def __init__(self, num_agents: int, input_dims: int, fc1_dims: int = 400, fc2_dims: int = 300, n_actions: int = 2,
                 device=None, dtype=torch.float32):
        super().__init__()
        kw = dict(device=device, dtype=dtype)
        n = num_agents

        def uni(shape, bound):
            return torch.nn.Parameter(torch.empty(*shape, **kw).uniform_(-bound, bound))

        f1, f2, f3 = fc1_dims ** -0.5, fc2_dims ** -0.5, 0.003          # ddpg_network.py:108-126
        self.w1, self.b1 = uni((n, input_dims, fc1_dims), f1), uni((n, 1, fc1_dims), f1)
        self.g1 = torch.nn.Parameter(torch.ones(n, 1, fc1_dims, **kw))
        self.be1 = torch.nn.Parameter(torch.zeros(n, 1, fc1_dims, **kw))
        self.w2, self.b2 = uni((n, fc1_dims, fc2_dims), f2), uni((n, 1, fc2_dims), f2)
        self.g2 = torch.nn.Parameter(torch.ones(n, 1, fc2_dims, **kw))
        self.be2 = torch.nn.Parameter(torch.zeros(n, 1, fc2_dims, **kw))
        self.w3, self.b3 = uni((n, fc2_dims, n_actions), f3), uni((n, 1, n_actions), f3)
        self.num_agents, self.input_dims = n, input_dims
        self._packed = None

    @classmethod
    def from_state_dicts(cls, sds: Sequence[Dict[str, torch.Tensor]], device=None, dtype=torch.float32):
        """Build from N reference `ActorNetwork.state_dict()`s (keys fc1/bn1/fc2/bn2/mu .weight/.bias)."""
        w1 = _stack(sds, "fc1.weight")
        self = cls(len(sds), w1.shape[2], w1.shape[1], sds[0]["fc2.weight"].shape[0], sds[0]["mu.weight"].shape[0],
                   device=device, dtype=dtype)
        with torch.no_grad():
            for wt, b, name in ((self.w1, self.b1, "fc1"), (self.w2, self.b2, "fc2"), (self.w3, self.b3, "mu")):
                wt.copy_(_stack(sds, name + ".weight").transpose(1, 2))
                b.copy_(_stack(sds, name + ".bias").unsqueeze(1))
            for g, be, name in ((self.g1, self.be1, "bn1"), (self.g2, self.be2, "bn2")):
                g.copy_(_stack(sds, name + ".weight").unsqueeze(1))
                be.copy_(_stack(sds, name + ".bias").unsqueeze(1))
        return self

    # ---- fused tensor-core path (csrc/flock_actor.cu through the C ABI) ----
    def pack_fused(self) -> torch.Tensor:
        """(Re)build the packed parameter image of `flock_actor_forward` from the current parameters
        (call after every optimiser step). No fallback: raises if the CUDA library is missing."""
        import ctypes

        from . import _lib
        lib = _lib.load_library()
        if self.w1.device.type != "cuda":
            raise RuntimeError("the fused actor kernel needs CUDA parameters (there is no CPU path)")
        fc1, fc2, na = self.w1.shape[2], self.w2.shape[2], self.w3.shape[2]
        srcs = [t.detach().float().contiguous() for t in (self.w1, self.b1, self.g1, self.be1, self.w2, self.b2, self.g2,
                                                          self.be2, self.w3, self.b3)]
        ptrs = (ctypes.c_void_p * 10)(*[t.data_ptr() for t in srcs])
        packed = torch.empty(lib.flock_actor_packed_bytes(self.num_agents), dtype=torch.uint8, device=self.w1.device)
        with torch.cuda.device(self.w1.device):
            _lib.check(lib.flock_actor_pack(self.num_agents, self.input_dims, fc1, fc2, na, ptrs, packed.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
        self._packed = packed
        self._packed_srcs = srcs      # keep the fp32 sources alive until the pack kernel has run
        return packed

    @torch.no_grad()
    def forward_fused(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None, ou_state: Optional[torch.Tensor] = None,
                      ou_theta: float = 0.2, ou_mu: float = 0.0, ou_sigma: float = 0.15, ou_dt: float = 1e-2, seed: int = 0,
                      step: int = 0, env_offset: int = 0, counters=None) -> torch.Tensor:
        """Same function as `forward` in ONE kernel launch on the tensor cores (bf16 operands, fp32
        accumulation / LayerNorm): `(E, N, ...)` float32 CUDA observations -> `(E, N, 2)` actions.

        With `ou_state` (an `(E, N, 2)` float32 CUDA tensor, zeros after a reset) the learner's exploration noise is
        fused into the same launch: one Ornstein-Uhlenbeck process per (env, agent, action) as in
        `OUActionNoiseGPU` (learners/maddpg_shared_critic/utils.py:6-21), `actions = mu + x` with
        `x <- x + theta (mu_ou - x) dt + sigma sqrt(dt) N(0,1)` updated in place; normals from Philox(seed; env, agent, step).

        `step` is a HOST scalar: captured in a CUDA graph (e.g. `rollout.GraphedRollout`) it would be frozen and every
        replay would redraw the same normals. Pass `counters=env.noise_counters` (per-env device counters, added to
        `step` inside the kernel) whenever the launch may be captured."""
        from . import _lib
        lib = _lib.load_library()
        if getattr(self, "_packed", None) is None:
            self.pack_fused()
        ring = hasattr(obs, "ring") and hasattr(obs, "head")        # vec_env.ObsRing: the env's history, read in place
        if ring:
            E, H, N, k = obs.ring.shape
            x, in_dims = obs.ring, H * k
        else:
            E, N = obs.shape[0], obs.shape[1]
            x = obs.reshape(E, N, -1)
            in_dims = x.shape[2]
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        if out is None:
            out = torch.empty(E, N, 2, dtype=torch.float32, device=x.device)
        if ou_state is not None and (ou_state.shape != (E, N, 2) or ou_state.dtype != torch.float32 or not ou_state.is_contiguous()):
            raise ValueError("ou_state must be a contiguous float32 (E, N, 2) tensor")
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            if ring:
                _lib.check(lib.flock_actor_forward_ring(
                    self._packed.data_ptr(), x.data_ptr(), obs.head.data_ptr(), out.data_ptr(), E, N, H, k,
                    None if ou_state is None else ou_state.data_ptr(), float(ou_theta), float(ou_mu), float(ou_sigma),
                    float(ou_dt), int(seed), int(step) & 0xFFFFFFFF, int(env_offset), _lib.noise_counters(counters), stream))
            elif ou_state is None:
                _lib.check(lib.flock_actor_forward(self._packed.data_ptr(), x.data_ptr(), out.data_ptr(), E, N, in_dims, stream))
            else:
                _lib.check(lib.flock_actor_forward_ou(self._packed.data_ptr(), x.data_ptr(), out.data_ptr(), E, N, in_dims,
                                                      ou_state.data_ptr(), float(ou_theta), float(ou_mu), float(ou_sigma),
                                                      float(ou_dt), int(seed), int(step) & 0xFFFFFFFF, int(env_offset),
                                                      _lib.noise_counters(counters), stream))
        return out

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        if hasattr(obs, "window"):            # vec_env.ObsRing: the PyTorch path works on the materialised window
            obs = obs.window()
        E, N = obs.shape[0], obs.shape[1]
        x = obs.reshape(E, N, -1).transpose(0, 1).to(self.w1.dtype)              # (N, E, in)
        x = torch.baddbmm(self.b1, x, self.w1)
        x = F.relu(F.layer_norm(x, x.shape[-1:]) * self.g1 + self.be1)
        x = torch.baddbmm(self.b2, x, self.w2)
        x = F.relu(F.layer_norm(x, x.shape[-1:]) * self.g2 + self.be2)
        mu = torch.tanh(torch.baddbmm(self.b3, x, self.w3))
        return mu.transpose(0, 1).float().contiguous()                           # (E, N, n_actions)


class BatchedRnnActors(torch.nn.Module):
    """N per-agent recurrent MADDPG actors (`Actor`, learners/maddpg_official_rnn/net.py:14-72 -- the policy of the
    reference's default `main.py` loop): fce(in -> 32) -> GRUCell(32, 32) -> fc1(32 -> 400) -> ReLU -> fc2(400 -> 300)
    -> ReLU -> [linear_speed -> (tanh + 1) / 2 | angular_speed -> 1.5 tanh], evaluated for all envs and agents at once
    (one `baddbmm` per layer over the agent dimension; plain PyTorch -- the fused tensor-core kernel covers the
    shared-critic actor, this one is the batched reference path for the recurrent learner).
    `forward(obs (E, N, in), hidden (E, N, 32)) -> (actions (E, N, 2), next_hidden (E, N, 32))`."""

    def __init__(self, num_agents: int, input_dims: int, hidden1: int = 400, hidden2: int = 300, hidden_rnn: int = 32,
                 init_w: float = 3e-3, device=None, dtype=torch.float32):
        super().__init__()
        kw = dict(device=device, dtype=dtype)
        n = num_agents

        def lin(i, o, bound=None):
            b = (i ** -0.5) if bound is None else bound
            return (torch.nn.Parameter(torch.empty(n, i, o, **kw).uniform_(-b, b)),
                    torch.nn.Parameter(torch.empty(n, 1, o, **kw).uniform_(-(i ** -0.5), i ** -0.5)))

        self.we, self.be = lin(input_dims, hidden_rnn)
        self.w_ih, self.b_ih = lin(hidden_rnn, 3 * hidden_rnn)
        self.w_hh, self.b_hh = lin(hidden_rnn, 3 * hidden_rnn)
        self.w1, self.b1 = lin(hidden_rnn, hidden1)
        self.w2, self.b2 = lin(hidden1, hidden2)
        self.wl, self.bl = lin(hidden2, 1, init_w)         # linear_speed, net.py:47
        self.wa, self.ba = lin(hidden2, 1, init_w)         # angular_speed, net.py:48
        self.num_agents, self.input_dims, self.hidden_rnn = n, input_dims, hidden_rnn
        self._packed = None

    @classmethod
    def from_state_dicts(cls, sds: Sequence[Dict[str, torch.Tensor]], device=None, dtype=torch.float32):
        """Build from N reference `Actor.state_dict()`s (keys fce / gru / fc1 / fc2 / linear_speed / angular_speed)."""
        self = cls(len(sds), sds[0]["fce.weight"].shape[1], sds[0]["fc1.weight"].shape[0], sds[0]["fc2.weight"].shape[0],
                   sds[0]["fce.weight"].shape[0], device=device, dtype=dtype)
        pairs = ((self.we, self.be, "fce.weight", "fce.bias"), (self.w_ih, self.b_ih, "gru.weight_ih", "gru.bias_ih"),
                 (self.w_hh, self.b_hh, "gru.weight_hh", "gru.bias_hh"), (self.w1, self.b1, "fc1.weight", "fc1.bias"),
                 (self.w2, self.b2, "fc2.weight", "fc2.bias"), (self.wl, self.bl, "linear_speed.weight", "linear_speed.bias"),
                 (self.wa, self.ba, "angular_speed.weight", "angular_speed.bias"))
        with torch.no_grad():
            for wt, b, kw_, kb_ in pairs:
                wt.copy_(_stack(sds, kw_).transpose(1, 2))
                b.copy_(_stack(sds, kb_).unsqueeze(1))
        return self

    def init_hidden(self, batch_size: int = 1) -> torch.Tensor:
        return torch.zeros(batch_size, self.num_agents, self.hidden_rnn, device=self.we.device)

    # ---- fused path (csrc/flock_rnn_actor.cu through the C ABI): fp32 GRU front end + tensor-core MLP ----
    def pack_fused(self) -> torch.Tensor:
        """(Re)build the packed MLP parameter image (call after every optimiser step). No fallback."""
        import ctypes

        from . import _lib
        lib = _lib.load_library()
        if self.we.device.type != "cuda":
            raise RuntimeError("the fused recurrent actor kernel needs CUDA parameters (there is no CPU path)")
        srcs = [t.detach().float().contiguous() for t in (self.w1, self.b1, self.w2, self.b2, self.wl, self.bl, self.wa, self.ba)]
        ptrs = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in srcs])
        packed = torch.empty(lib.flock_rnn_actor_packed_bytes(self.num_agents), dtype=torch.uint8, device=self.we.device)
        with torch.cuda.device(self.we.device):
            _lib.check(lib.flock_rnn_actor_pack(self.num_agents, self.hidden_rnn, self.w1.shape[2], self.w2.shape[2], 2, ptrs,
                                                packed.data_ptr(), torch.cuda.current_stream().cuda_stream))
        front = [t.detach().float().contiguous() for t in (self.we, self.be, self.w_ih, self.b_ih, self.w_hh, self.b_hh)]
        self._packed, self._packed_srcs = packed, srcs
        self._front, self._front_ptrs = front, (ctypes.c_void_p * 6)(*[t.data_ptr() for t in front])
        # the same front end for the tensor-core kernel (flock_gru_tc.cu): split-fp16 weight images
        self._front_packed = None
        if self.hidden_rnn == 32 and self.input_dims <= 16:
            fp = torch.empty(lib.flock_gru_tc_packed_bytes(0, self.num_agents), dtype=torch.uint8, device=self.we.device)
            with torch.cuda.device(self.we.device):
                _lib.check(lib.flock_gru_tc_pack(0, self.num_agents, self.input_dims, 0, self._front_ptrs, fp.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
            self._front_packed = fp
        return packed

    @torch.no_grad()
    def forward_fused(self, obs: torch.Tensor, hidden: torch.Tensor, out: Optional[torch.Tensor] = None,
                      hidden_out: Optional[torch.Tensor] = None, ou_state: Optional[torch.Tensor] = None,
                      ou_theta: float = 0.15, ou_mu: float = 0.0, ou_sigma: float = 0.2, ou_dt: float = 1e-2, seed: int = 0,
                      step: int = 0, env_offset: int = 0, counters=None, impl: str = "tc") -> Tuple[torch.Tensor, torch.Tensor]:
        """`forward` in two kernel launches: fce + GRUCell, then the 32-400-300-2 MLP on the tensor cores (bf16
        operands, fp32 accumulation). `impl="tc"` (default): the front end runs on the tensor cores too, with split
        fp16 operands (fp32-level accuracy, hidden state within 1e-5 of the fp32 module; csrc/flock_gru_tc.cu);
        `impl="fp32"`: the CUDA-core front kernel. `hidden_out` may be `hidden`.
        With `ou_state` ((E, N, 2) float32, zeros after a reset) the learner's Ornstein-Uhlenbeck exploration noise
        (agent.py:61, utils.py:43-47) is added in the same launch, one process per (env, agent, action). Pass
        `counters=env.noise_counters` when the launch may be captured in a CUDA graph (see BatchedActors.forward_fused)."""
        from . import _lib
        lib = _lib.load_library()
        if getattr(self, "_packed", None) is None:
            self.pack_fused()
        E, N = obs.shape[0], obs.shape[1]
        x = obs.reshape(E, N, -1)
        x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
        h = hidden if (hidden.dtype == torch.float32 and hidden.is_contiguous()) else hidden.float().contiguous()
        if out is None:
            out = torch.empty(E, N, 2, dtype=torch.float32, device=x.device)
        if hidden_out is None:
            hidden_out = torch.empty(E, N, self.hidden_rnn, dtype=torch.float32, device=x.device)
        if ou_state is not None and (ou_state.shape != (E, N, 2) or ou_state.dtype != torch.float32 or not ou_state.is_contiguous()):
            raise ValueError("ou_state must be a contiguous float32 (E, N, 2) tensor")
        if impl not in ("tc", "fp32"):
            raise ValueError("impl must be 'tc' or 'fp32'")
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            if impl == "tc" and self._front_packed is not None:
                _lib.check(lib.flock_rnn_actor_forward_tc(
                    self._packed.data_ptr(), self._front_packed.data_ptr(), x.data_ptr(), h.data_ptr(), hidden_out.data_ptr(),
                    out.data_ptr(), E, N, x.shape[2], None if ou_state is None else ou_state.data_ptr(), float(ou_theta),
                    float(ou_mu), float(ou_sigma), float(ou_dt), int(seed), int(step) & 0xFFFFFFFF, int(env_offset),
                    _lib.noise_counters(counters), stream))
            elif ou_state is None:
                _lib.check(lib.flock_rnn_actor_forward(self._packed.data_ptr(), self._front_ptrs, x.data_ptr(), h.data_ptr(),
                                                       hidden_out.data_ptr(), out.data_ptr(), E, N, x.shape[2], stream))
            else:
                _lib.check(lib.flock_rnn_actor_forward_ou(self._packed.data_ptr(), self._front_ptrs, x.data_ptr(), h.data_ptr(),
                                                          hidden_out.data_ptr(), out.data_ptr(), E, N, x.shape[2],
                                                          ou_state.data_ptr(), float(ou_theta), float(ou_mu), float(ou_sigma),
                                                          float(ou_dt), int(seed), int(step) & 0xFFFFFFFF, int(env_offset),
                                                          _lib.noise_counters(counters), stream))
        return out, hidden_out

    def forward(self, obs: torch.Tensor, hidden: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        E, N = obs.shape[0], obs.shape[1]
        x = obs.reshape(E, N, -1).transpose(0, 1).to(self.we.dtype)               # (N, E, in)
        h = hidden.transpose(0, 1).to(x.dtype)
        x = torch.baddbmm(self.be, x, self.we)                                     # fce: no activation (net.py:57)
        gi = torch.baddbmm(self.b_ih, x, self.w_ih)
        gh = torch.baddbmm(self.b_hh, h, self.w_hh)
        i_r, i_z, i_n = gi.chunk(3, dim=-1)
        h_r, h_z, h_n = gh.chunk(3, dim=-1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        x = (1 - z) * n + z * h                                                    # torch.nn.GRUCell
        next_hidden = x.transpose(0, 1).float().contiguous()
        x = F.relu(torch.baddbmm(self.b1, x, self.w1))
        x = F.relu(torch.baddbmm(self.b2, x, self.w2))
        linear = (torch.tanh(torch.baddbmm(self.bl, x, self.wl)) + 1) / 2          # net.py:66-67
        angular = torch.tanh(torch.baddbmm(self.ba, x, self.wa)) * 1.5             # net.py:70-71
        return torch.cat([linear, angular], dim=-1).transpose(0, 1).float().contiguous(), next_hidden


class BatchedQNet(torch.nn.Module):
    """The VDN per-agent Q networks (`QNet`, learners/vdn/net.py:11-37): Linear(n_obs,64)-ReLU-
    Linear(64,32)-ReLU [-GRUCell(32,32)] -Linear(32,n_actions) per agent, batched over agents.
    `forward(obs (E,N,n_obs), hidden (E,N,32)) -> (q (E,N,A), hidden)`; `sample_action` is the
    epsilon-greedy of net.py:52-58 with the exploration decision per env, on the device."""

    def __init__(self, num_agents: int, n_obs: int, n_actions: int, recurrent: bool = False, hx_size: int = 32,
                 device=None, dtype=torch.float32):
        super().__init__()
        kw = dict(device=device, dtype=dtype)
        n = num_agents

        def lin(i, o):
            bound = i ** -0.5
            return (torch.nn.Parameter(torch.empty(n, i, o, **kw).uniform_(-bound, bound)),
                    torch.nn.Parameter(torch.empty(n, 1, o, **kw).uniform_(-bound, bound)))

        self.w1, self.b1 = lin(n_obs, 64)
        self.w2, self.b2 = lin(64, hx_size)
        self.wq, self.bq = lin(hx_size, n_actions)
        self.recurrent, self.hx_size, self.num_agents = recurrent, hx_size, n
        if recurrent:
            self.w_ih, self.b_ih = lin(hx_size, 3 * hx_size)
            self.w_hh, self.b_hh = lin(hx_size, 3 * hx_size)

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], num_agents: int, device=None, dtype=torch.float32):
        """Load a reference `QNet.state_dict()` (keys agent_feature_i.{0,2}, agent_gru_i, agent_q_i)."""
        rec = any(k.startswith("agent_gru_0") for k in sd)
        n_obs = sd["agent_feature_0.0.weight"].shape[1]
        self = cls(num_agents, n_obs, sd["agent_q_0.weight"].shape[0], rec, sd["agent_feature_0.2.weight"].shape[0],
                   device=device, dtype=dtype)
        per = lambda fmt: torch.stack([sd[fmt.format(i)] for i in range(num_agents)])
        with torch.no_grad():
            self.w1.copy_(per("agent_feature_{}.0.weight").transpose(1, 2)); self.b1.copy_(per("agent_feature_{}.0.bias").unsqueeze(1))
            self.w2.copy_(per("agent_feature_{}.2.weight").transpose(1, 2)); self.b2.copy_(per("agent_feature_{}.2.bias").unsqueeze(1))
            self.wq.copy_(per("agent_q_{}.weight").transpose(1, 2)); self.bq.copy_(per("agent_q_{}.bias").unsqueeze(1))
            if rec:
                self.w_ih.copy_(per("agent_gru_{}.weight_ih").transpose(1, 2)); self.b_ih.copy_(per("agent_gru_{}.bias_ih").unsqueeze(1))
                self.w_hh.copy_(per("agent_gru_{}.weight_hh").transpose(1, 2)); self.b_hh.copy_(per("agent_gru_{}.bias_hh").unsqueeze(1))
        return self

    def init_hidden(self, batch_size: int = 1) -> torch.Tensor:
        return torch.zeros(batch_size, self.num_agents, self.hx_size, device=self.w1.device)

    def forward(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        x = obs.transpose(0, 1).to(self.w1.dtype)                                 # (N, E, n_obs)
        x = F.relu(torch.baddbmm(self.b1, x, self.w1))
        x = F.relu(torch.baddbmm(self.b2, x, self.w2))
        if self.recurrent:
            h = hidden.transpose(0, 1).to(x.dtype)
            gi = torch.baddbmm(self.b_ih, x, self.w_ih)
            gh = torch.baddbmm(self.b_hh, h, self.w_hh)
            i_r, i_z, i_n = gi.chunk(3, dim=-1)
            h_r, h_z, h_n = gh.chunk(3, dim=-1)
            r = torch.sigmoid(i_r + h_r)
            z = torch.sigmoid(i_z + h_z)
            n = torch.tanh(i_n + r * h_n)
            x = (1 - z) * n + z * h                                                # torch.nn.GRUCell
            next_hidden = x.transpose(0, 1).float()
        else:
            next_hidden = torch.empty(obs.shape[0], self.num_agents, self.hx_size, device=obs.device)
        q = torch.baddbmm(self.bq, x, self.wq).transpose(0, 1).float()             # (E, N, A)
        return q, next_hidden

    # ---- fused path (csrc/flock_qnet.cu through the C ABI): one fp32 launch for all envs and agents ----
    @torch.no_grad()
    def _fused(self, obs, hidden, want_q: bool, want_actions: bool, epsilon: float, step: int, seed: int, env_offset: int,
               out=None, hidden_out=None, counters=None, impl: str = "tc"):
        import ctypes

        from . import _lib
        lib = _lib.load_library()
        if self.w1.device.type != "cuda":
            raise RuntimeError("the fused Q-network kernel needs CUDA parameters (there is no CPU path)")
        E, N, n_obs = obs.shape
        A = self.wq.shape[2]
        x = obs if (obs.dtype == torch.float32 and obs.is_contiguous()) else obs.float().contiguous()
        names = ["w1", "b1", "w2", "b2", "wq", "bq"] + (["w_ih", "b_ih", "w_hh", "b_hh"] if self.recurrent else [])
        if impl not in ("tc", "fp32"):
            raise ValueError("impl must be 'tc' or 'fp32'")
        # parameters moved or were updated in place (optimiser step): rebuild the pointer table / the packed image
        key = tuple((getattr(self, n).data_ptr(), getattr(self, n)._version) for n in names)
        cache = getattr(self, "_fused_ptrs", None)
        if cache is None or cache[0] != key:
            srcs = [getattr(self, n).detach() for n in names]
            srcs = [t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous() for t in srcs]
            ptrs = (ctypes.c_void_p * 10)(*([t.data_ptr() for t in srcs] + [None] * (10 - len(srcs))))
            packed = None
            if self.recurrent and self.hx_size == 32 and n_obs <= 16 and A <= 16 and self.w1.shape[2] == 64:
                # tensor-core path (csrc/flock_gru_tc.cu): split-fp16 weight images, re-packed after every update
                packed = torch.empty(lib.flock_gru_tc_packed_bytes(1, N), dtype=torch.uint8, device=x.device)
                with torch.cuda.device(x.device):
                    _lib.check(lib.flock_gru_tc_pack(1, N, n_obs, A, ptrs, packed.data_ptr(), torch.cuda.current_stream().cuda_stream))
            cache = self._fused_ptrs = (key, ptrs, srcs, packed)
        ptrs, packed = cache[1], cache[3]
        dev = x.device
        q = torch.empty(E, N, A, dtype=torch.float32, device=dev) if want_q else None
        act = (out if out is not None else torch.empty(E, N, dtype=torch.float32, device=dev)) if want_actions else None
        h_in = h_out = None
        if self.recurrent:
            h_in = hidden if (hidden.dtype == torch.float32 and hidden.is_contiguous()) else hidden.float().contiguous()
            h_out = hidden_out if hidden_out is not None else torch.empty(E, N, self.hx_size, dtype=torch.float32, device=dev)
        p = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            if impl == "tc" and packed is not None:
                _lib.check(lib.flock_qnet_forward_tc(packed.data_ptr(), x.data_ptr(), p(h_in), p(q), p(h_out), p(act), E, N, n_obs, A,
                                                     float(epsilon), int(seed), int(step) & 0xFFFFFFFF, int(env_offset),
                                                     _lib.noise_counters(counters), torch.cuda.current_stream().cuda_stream))
            else:
                _lib.check(lib.flock_qnet_forward(ptrs, int(self.recurrent), x.data_ptr(), p(h_in), p(q), p(h_out), p(act), E, N,
                                                  n_obs, A, float(epsilon), int(seed), int(step) & 0xFFFFFFFF, int(env_offset),
                                                  _lib.noise_counters(counters), torch.cuda.current_stream().cuda_stream))
        if h_out is None:
            h_out = torch.empty(E, N, self.hx_size, device=dev)      # like `forward`: unused without the GRU
        return q, h_out, act

    def forward_fused(self, obs: torch.Tensor, hidden: Optional[torch.Tensor] = None, impl: str = "tc") -> Tuple[torch.Tensor, torch.Tensor]:
        """`forward` in one kernel launch. `impl="tc"` (default, recurrent nets): tcgen05 MMAs with split fp16 operands
        (fp32-level accuracy: within 1e-5 of the fp32 module); `impl="fp32"`: the CUDA-core kernel (same arithmetic as
        PyTorch up to summation order; also what non-recurrent nets use)."""
        q, h, _ = self._fused(obs, hidden, True, False, 0.0, 0, 0, 0, impl=impl)
        return q, h

    def sample_action_fused(self, obs: torch.Tensor, hidden: Optional[torch.Tensor], epsilon: float, step: int = 0,
                            seed: int = 0, env_offset: int = 0, out: Optional[torch.Tensor] = None,
                            hidden_out: Optional[torch.Tensor] = None, counters=None, impl: str = "tc") -> Tuple[torch.Tensor, torch.Tensor]:
        """`sample_action` (net.py:52-58) in one kernel launch: Q-values, argmax and the per-env epsilon-greedy
        decision never leave the SM. Exploration draws are Philox(seed; env_offset + env, agent, step), i.e.
        reproducible and invariant under env sharding -- pass the rollout step as `step`, or, when the launch may be
        captured in a CUDA graph (a host `step` is frozen there), `counters=env.noise_counters`. `out` (E, N) and
        `hidden_out` (E, N, 32; may be `hidden`) receive the results in place when given."""
        _, h, act = self._fused(obs, hidden, False, True, epsilon, step, seed, env_offset, out=out, hidden_out=hidden_out,
                                counters=counters, impl=impl)
        return act, h

    @torch.no_grad()
    def sample_action(self, obs: torch.Tensor, hidden: Optional[torch.Tensor], epsilon: float,
                      generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        q, hidden = self.forward(obs, hidden)
        E, N, A = q.shape
        explore = torch.rand(E, device=q.device, generator=generator) <= epsilon   # one decision per env, net.py:54
        rand_a = torch.randint(0, A, (E, N), device=q.device, generator=generator)
        action = torch.where(explore[:, None], rand_a, q.argmax(dim=2)).float()    # float-coded ids, net.py:56-57
        return action, hidden
