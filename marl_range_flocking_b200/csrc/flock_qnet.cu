// flock_qnet.cu -- fused VDN action selection for the batched rollout (BASELINE configs[3]
// "gym_flock_uw_discrete with VDN action selection"): QNet.forward + QNet.sample_action of
// learners/vdn/net.py:11-58 for all envs and agents in ONE launch:
//     Linear(n_obs, 64) - ReLU - Linear(64, 32) - ReLU - [GRUCell(32, 32)] - Linear(32, n_actions),
//     then epsilon-greedy with ONE exploration decision per env (net.py:54) and float-coded action ids.
// One weight set per agent (net.py:17-25), evaluated upstream in a Python loop over agents.
//
// This is 8.6 kFLOP-pairs per agent-step with K = 4 / 64 / 32 -- far too small and too argmax-sensitive
// for reduced-precision tensor-core operands (greedy actions must equal the fp32 module's), so it is an
// fp32 CUDA-core kernel: CTA = 256 envs x one agent, the agent's 35 KB of weights staged once in shared
// memory and read back as warp-wide broadcast 128-bit loads (4 weights per LDS, 4 FMAs per LDS), all
// activations in registers, no intermediate ever written to global memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include "flock_device.cuh"
#include "flock_launch.h"

namespace flock {
namespace qnet {

constexpr int kHid1 = 64, kHx = 32, kMaxObs = 16, kMaxAct = 16;
constexpr int kThreads = 256;
constexpr uint32_t kTagExplore = 4u, kTagRandAct = 5u;   // Philox stream tags (0-3 belong to the env kernels)

struct Args {
    const float *w1, *b1, *w2, *b2, *wq, *bq, *w_ih, *b_ih, *w_hh, *b_hh;   // [A][in][out] / [A][out], fp32
    const float* obs;        // [E][A][n_obs]
    const float* hidden_in;  // [E][A][32] (recurrent)
    float* q_out;            // [E][A][n_actions] or null
    float* hidden_out;       // [E][A][32] or null
    float* actions;          // [E][A] float-coded ids or null
    int E, A, n_obs, n_act, env_offset;
    float epsilon;
    uint32_t seed_lo, seed_hi, step;
    const int32_t* env_step;      // [E] device step counters added to `step` (nullable; replay-safe under CUDA graphs)
    const uint32_t* env_epoch;    // [E] device epoch counters folded into the tag words (nullable)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// out[j] = bias[j] + sum_i in[i] * W[i][j], j < NOUT (NOUT % 4 == 0), W row-major [NIN][NOUT] in shared memory
template <int NIN, int NOUT>
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ bias, const float (&in)[NIN],
                                      float (&out)[NOUT], int nin) {
#pragma unroll
    for (int j = 0; j < NOUT; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias + j);
        out[j] = b.x; out[j + 1] = b.y; out[j + 2] = b.z; out[j + 3] = b.w;
    }
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
        if (i < nin) {
#pragma unroll
            for (int j = 0; j < NOUT; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W + i * NOUT + j);
                out[j] = fmaf(in[i], w.x, out[j]);
                out[j + 1] = fmaf(in[i], w.y, out[j + 1]);
                out[j + 2] = fmaf(in[i], w.z, out[j + 2]);
                out[j + 3] = fmaf(in[i], w.w, out[j + 3]);
            }
        }
    }
}

template <bool REC>
__global__ void __launch_bounds__(kThreads, 2) flock_qnet_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) float sw[];
    const int agent = blockIdx.x;
    const int n_obs = a.n_obs, n_act = a.n_act;
    const int actp = (n_act + 3) & ~3;                    // q head padded to a float4 multiple (zero weights)
    // shared layout: W1 [n_obs][64] | b1 [64] | W2 [64][32] | b2 [32] | Wq [32][actp] | bq [actp] | (GRU) Wih [32][96] | bih | Whh | bhh
    float* sW1 = sw;
    float* sb1 = sW1 + n_obs * kHid1;
    float* sW2 = sb1 + kHid1;
    float* sb2 = sW2 + kHid1 * kHx;
    float* sWq = sb2 + kHx;
    float* sbq = sWq + kHx * actp;
    float* sWih = sbq + actp;
    float* sbih = sWih + kHx * 3 * kHx;
    float* sWhh = sbih + 3 * kHx;
    float* sbhh = sWhh + kHx * 3 * kHx;
    // stage the agent's weights: 128-bit copies (every block is a 16-byte multiple; the API checks alignment)
    auto copy4 = [&](float* dst, const float* src, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int t = threadIdx.x; t < (n >> 2); t += kThreads) d4[t] = s4[t];
    };
    copy4(sW1, a.w1 + (size_t)agent * n_obs * kHid1, n_obs * kHid1);
    copy4(sb1, a.b1 + (size_t)agent * kHid1, kHid1);
    copy4(sW2, a.w2 + (size_t)agent * kHid1 * kHx, kHid1 * kHx);
    copy4(sb2, a.b2 + (size_t)agent * kHx, kHx);
    for (int t = threadIdx.x; t < kHx * actp; t += kThreads) {
        const int i = t / actp, j = t % actp;
        sWq[t] = j < n_act ? a.wq[((size_t)agent * kHx + i) * n_act + j] : 0.0f;
    }
    for (int t = threadIdx.x; t < actp; t += kThreads) sbq[t] = t < n_act ? a.bq[(size_t)agent * n_act + t] : 0.0f;
    if (REC) {
        copy4(sWih, a.w_ih + (size_t)agent * kHx * 3 * kHx, kHx * 3 * kHx);
        copy4(sWhh, a.w_hh + (size_t)agent * kHx * 3 * kHx, kHx * 3 * kHx);
        copy4(sbih, a.b_ih + (size_t)agent * 3 * kHx, 3 * kHx);
        copy4(sbhh, a.b_hh + (size_t)agent * 3 * kHx, 3 * kHx);
    }
    __syncthreads();

    const int env = blockIdx.y * kThreads + threadIdx.x;
    if (env >= a.E) return;
    const size_t ea = (size_t)env * a.A + agent;

    float x[kMaxObs];
#pragma unroll
    for (int i = 0; i < kMaxObs; ++i) x[i] = i < n_obs ? a.obs[ea * n_obs + i] : 0.0f;
    float h1[kHid1];
    dense<kMaxObs, kHid1>(sW1, sb1, x, h1, n_obs);
#pragma unroll
    for (int j = 0; j < kHid1; ++j) h1[j] = fmaxf(h1[j], 0.0f);
    float f[kHx];
    dense<kHid1, kHx>(sW2, sb2, h1, f, kHid1);
#pragma unroll
    for (int j = 0; j < kHx; ++j) f[j] = fmaxf(f[j], 0.0f);

    float q[kMaxAct];
#pragma unroll
    for (int j = 0; j < kMaxAct; ++j) q[j] = j < actp ? sbq[j] : 0.0f;
    // q += v * Wq[row][:] (row may be a runtime value: the accumulators keep static indices)
    auto head_row = [&](float v, int rowi) {
#pragma unroll
        for (int j = 0; j < kMaxAct; j += 4) {
            if (j < actp) {
                const float4 w = *reinterpret_cast<const float4*>(sWq + rowi * actp + j);
                q[j] = fmaf(v, w.x, q[j]); q[j + 1] = fmaf(v, w.y, q[j + 1]);
                q[j + 2] = fmaf(v, w.z, q[j + 2]); q[j + 3] = fmaf(v, w.w, q[j + 3]);
            }
        }
    };
    if (REC) {   // torch.nn.GRUCell: gates ordered r | z | n along the 96 outputs
        float hp[kHx];
        const float4* hin = reinterpret_cast<const float4*>(a.hidden_in + ea * kHx);
#pragma unroll
        for (int j = 0; j < kHx; j += 4) {
            const float4 v = hin[j >> 2];
            hp[j] = v.x; hp[j + 1] = v.y; hp[j + 2] = v.z; hp[j + 3] = v.w;
        }
        // Four hidden units per iteration (6 float4 accumulators). The loop over the unit chunks is NOT unrolled:
        // fully unrolled the kernel was 250 KB of code and instruction fetch (stall_no_inst 58 %) set the pace.
        // The new state goes straight to global memory and into the Q head, so nothing is indexed by `u`.
#pragma unroll 1
        for (int u = 0; u < kHx; u += 4) {
            float gi[3][4], gh[3][4];
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const float4 bi = *reinterpret_cast<const float4*>(sbih + g * kHx + u);
                const float4 bh = *reinterpret_cast<const float4*>(sbhh + g * kHx + u);
                gi[g][0] = bi.x; gi[g][1] = bi.y; gi[g][2] = bi.z; gi[g][3] = bi.w;
                gh[g][0] = bh.x; gh[g][1] = bh.y; gh[g][2] = bh.z; gh[g][3] = bh.w;
            }
#pragma unroll
            for (int i = 0; i < kHx; ++i) {
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const float4 wi = *reinterpret_cast<const float4*>(sWih + i * 3 * kHx + g * kHx + u);
                    const float4 wh = *reinterpret_cast<const float4*>(sWhh + i * 3 * kHx + g * kHx + u);
                    gi[g][0] = fmaf(f[i], wi.x, gi[g][0]); gi[g][1] = fmaf(f[i], wi.y, gi[g][1]);
                    gi[g][2] = fmaf(f[i], wi.z, gi[g][2]); gi[g][3] = fmaf(f[i], wi.w, gi[g][3]);
                    gh[g][0] = fmaf(hp[i], wh.x, gh[g][0]); gh[g][1] = fmaf(hp[i], wh.y, gh[g][1]);
                    gh[g][2] = fmaf(hp[i], wh.z, gh[g][2]); gh[g][3] = fmaf(hp[i], wh.w, gh[g][3]);
                }
            }
            const float4 hpu4 = hin[u >> 2];           // hp[u..u+3] again (L1 hit) instead of a runtime register index
            const float hpu[4] = {hpu4.x, hpu4.y, hpu4.z, hpu4.w};
            float hn[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float r = sigmoidf_(gi[0][c] + gh[0][c]);
                const float z = sigmoidf_(gi[1][c] + gh[1][c]);
                const float n = tanhf(gi[2][c] + r * gh[2][c]);
                hn[c] = (1.0f - z) * n + z * hpu[c];
                head_row(hn[c], u + c);
            }
            if (a.hidden_out != nullptr)
                reinterpret_cast<float4*>(a.hidden_out + ea * kHx)[u >> 2] = make_float4(hn[0], hn[1], hn[2], hn[3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kHx; ++i) head_row(f[i], i);
    }
    if (a.q_out != nullptr) {
#pragma unroll
        for (int j = 0; j < kMaxAct; ++j)
            if (j < n_act) a.q_out[ea * n_act + j] = q[j];
    }
    if (a.actions != nullptr) {
        // greedy: first maximum, like torch.argmax; exploration: one decision per env (net.py:54), then a
        // uniform action id per agent (net.py:56)
        int best = 0;
        float bv = q[0];
#pragma unroll
        for (int j = 1; j < kMaxAct; ++j)
            if (j < n_act && q[j] > bv) {
                bv = q[j];
                best = j;
            }
        const uint32_t ge = (uint32_t)(a.env_offset + env);
        const uint32_t c2 = a.step + (a.env_step != nullptr ? (uint32_t)a.env_step[env] : 0u);
        const uint32_t ep = a.env_epoch != nullptr ? (a.env_epoch[env] << 4) : 0u;
        const uint4 re = philox4x32_10(ge, 0xffffffffu, c2, kTagExplore + ep, a.seed_lo, a.seed_hi);
        if (u24(re.x) <= a.epsilon && a.epsilon > 0.0f) {
            const uint4 ra = philox4x32_10(ge, (uint32_t)agent, c2, kTagRandAct + ep, a.seed_lo, a.seed_hi);
            best = (int)(((unsigned long long)ra.x * (unsigned long long)n_act) >> 32);
        }
        a.actions[ea] = (float)best;
    }
}

}  // namespace qnet

int qnet_max_obs() { return qnet::kMaxObs; }
int qnet_max_actions() { return qnet::kMaxAct; }

cudaError_t launch_qnet(const float* const* params, int recurrent, const float* obs, const float* hidden_in, float* q_out,
                        float* hidden_out, float* actions, int E, int A, int n_obs, int n_act, float epsilon, uint64_t seed,
                        uint32_t step, int env_offset, NoiseCounters ctr, cudaStream_t s) {
    qnet::Args a;
    a.env_step = ctr.env_step;
    a.env_epoch = ctr.env_epoch;
    a.w1 = params[0]; a.b1 = params[1]; a.w2 = params[2]; a.b2 = params[3]; a.wq = params[4]; a.bq = params[5];
    a.w_ih = recurrent ? params[6] : nullptr; a.b_ih = recurrent ? params[7] : nullptr;
    a.w_hh = recurrent ? params[8] : nullptr; a.b_hh = recurrent ? params[9] : nullptr;
    a.obs = obs; a.hidden_in = hidden_in; a.q_out = q_out; a.hidden_out = hidden_out; a.actions = actions;
    a.E = E; a.A = A; a.n_obs = n_obs; a.n_act = n_act; a.env_offset = env_offset; a.epsilon = epsilon;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.step = step;
    const int actp = (n_act + 3) & ~3;
    size_t floats = (size_t)n_obs * qnet::kHid1 + qnet::kHid1 + qnet::kHid1 * qnet::kHx + qnet::kHx + qnet::kHx * actp + actp;
    if (recurrent) floats += 2 * (qnet::kHx * 3 * qnet::kHx + 3 * qnet::kHx);
    const size_t bytes = floats * sizeof(float);
    const dim3 grid((unsigned)A, (unsigned)((E + qnet::kThreads - 1) / qnet::kThreads));
    if (recurrent) {
        static DeviceOnce once;
        int sms = 0;
        const cudaError_t cfg = once.get(
            [] { return cudaFuncSetAttribute(qnet::flock_qnet_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); },
            &sms);
        if (cfg != cudaSuccess) return cfg;
        qnet::flock_qnet_kernel<true><<<grid, qnet::kThreads, bytes, s>>>(a);
    } else {
        qnet::flock_qnet_kernel<false><<<grid, qnet::kThreads, bytes, s>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace flock
