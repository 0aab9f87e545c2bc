// flock_gru_tc.cu -- the recurrent front ends of the fused policies on the tensor cores (tcgen05 / TMEM):
//   mode A  recurrent MADDPG actor front (learners/maddpg_official_rnn/net.py:53-58):
//           e = fce(obs) ; h' = GRUCell(e, h)                                     -> hidden_out (the MLP kernel's input)
//   mode B  VDN QNet.forward + sample_action (learners/vdn/net.py:27-58), recurrent:
//           x = ReLU(Linear(64, 32)(ReLU(Linear(n_obs, 64)(obs)))) ; h' = GRUCell(x, h) ; q = Linear(32, A)(h')
//           + per-env epsilon-greedy                                              -> hidden_out, q_out, actions
// one weight set per agent, all envs and agents in ONE launch. These are chains of tiny GEMMs (K = 16..64,
// N = 16..96) whose fp32 CUDA-core form (flock_qnet.cu, flock_rnn_front_kernel) is bound by broadcast shared-memory
// weight reads at ~25 % of the FP32 peak. Here the K >= 32 layers are tcgen05.mma over a tile of 128 env rows:
//   * fp32-level accuracy from the 16-bit tensor cores by SPLIT OPERANDS: v = hi + lo with hi = fp16(v),
//     lo = fp16(v - hi) (22 significant bits; fp16 rather than bf16 pieces, whose two pieces keep only 16 bits and needed
//     a third piece for the recurrent actor's unsquashed GRU input), and every product is the three MMAs
//     hi*hi + hi*lo + lo*hi into the same TMEM accumulator (fp32). Measured against the fp32 modules: hidden state and
//     Q-values within 1e-5 (tests: 2e-5 / 5e-5), greedy actions equal wherever the top-two Q gap exceeds 1e-4 -- the
//     recurrent state stays fp32 in memory. Range: activations and weights below 6.5e4 in magnitude (fp16).
//   * the observation layer (K = n_obs <= 16) runs on the CUDA cores, in fp32, while the row is staged: it is ~100 FMAs
//     per thread, and as an MMA it cost a full MMA -> epilogue round trip per item -- the kernel is bound by those
//     round trips (four epilogue warps or eight: same time), not by arithmetic.
//   * weights are pre-packed once per parameter update (flock_gru_tc_pack) into the K-major core-matrix images the
//     MMA reads (hi and lo), loaded per agent with one 1-D TMA bulk copy; activations are written by the epilogue
//     threads as the next layer's A operand (hi and lo images) straight from TMEM: nothing intermediate leaves the SM.
//   * CTA = 8 epilogue warps (two threads per env row = TMEM lane, one column half each) + 1 MMA / TMA warp, 256 TMEM
//     columns, 75-90 KB of shared memory: two CTAs per SM, so one CTA's MMA -> epilogue round trips overlap the other's.
//     Work item = (agent, tile of 128 envs), dealt out in contiguous runs (weights reloaded only when the agent changes).
// UMMA descriptors, TMEM loads, mbarrier plumbing: flock_tc.cuh (shared with flock_actor.cu / flock_rnn_actor.cu).
#include <cmath>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "flock_device.cuh"
#include "flock_launch.h"
#include "flock_tc.cuh"

namespace flock {
namespace grutc {

using namespace tc;

constexpr int kRows = 128;
constexpr int kHx = 32;           // GRU width (hx_size / hidden_rnn)
constexpr int kG = 3 * kHx;       // gate columns r | z | n
constexpr int kObsPad = 16;       // largest n_obs
constexpr int kH1 = 64;           // VDN feature layer width
constexpr int kActPad = 16;       // VDN head N (UMMA N is a multiple of 16)
constexpr uint32_t kTagExplore = 4u, kTagRandAct = 5u;   // same Philox tags as flock_qnet.cu

template <int MODE>
struct Layout {
    static constexpr bool kVdn = MODE == 1;
    static constexpr int kN0 = kVdn ? kH1 : kHx;                 // output width of the observation layer
    // fp16 weight images, bytes per piece (hi or lo): [K step][k-group][n] x 16 B
    static constexpr int kW2 = kVdn ? kH1 * kHx * 2 : 0;
    static constexpr int kWg = kHx * kG * 2;
    static constexpr int kWq = kVdn ? kHx * kActPad * 2 : 0;
    static constexpr int kOffW2 = 0, kOffWih = kOffW2 + kW2, kOffWhh = kOffWih + kWg, kOffWq = kOffWhh + kWg,
                         kPiece = kOffWq + kWq;
    static constexpr int kOffW0f = 2 * kPiece;                    // fp32 observation-layer weights [kObsPad][kN0]
    static constexpr int kW0f = kObsPad * kN0 * 4;
    // fp32 parameters: b0 [kN0] | b2 [32] (vdn) | b_ih [96] | b_hh [96] | bq [16] (vdn)
    static constexpr int kPb0 = 0, kPb2 = kPb0 + kN0, kPbih = kPb2 + (kVdn ? kHx : 0), kPbhh = kPbih + kG,
                         kPbq = kPbhh + kG, kParamFloats = kPbq + (kVdn ? kActPad : 0);
    static constexpr int kOffPar = kOffW0f + kW0f;
    static constexpr int kBlobBytes = kOffPar + kParamFloats * 4;
    // shared memory: blob | A operand images | barriers
    static constexpr int kAx1 = (kVdn ? kH1 : kHx) * kRows * 2;   // H1 (vdn, K = 64), later X (K = 32) in the same place
    static constexpr int kAh = kHx * kRows * 2;                   // h in, later h' (vdn head input)
    static constexpr int kOffBlob = 0;
    static constexpr int kOffX1 = (kBlobBytes + 127) & ~127;
    static constexpr int kOffH = kOffX1 + 2 * kAx1;
    static constexpr int kOffBar = kOffH + 2 * kAh;
    static constexpr int kNumBars = 8;
    static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 128;
    static constexpr int kTmemCols = 256;
    // TMEM columns: gi [0, 96), gh [96, 192), vdn layer 2 [192, 224), q [224, 240)
    static constexpr int kColGi = 0, kColGh = 96, kColH2 = 192, kColQ = 224;
};
static_assert(Layout<1>::kBlobBytes % 16 == 0 && Layout<0>::kBlobBytes % 16 == 0, "bulk copies move 16-byte units");
static_assert(Layout<1>::kSmemBytes <= 113 * 1024, "two CTAs per SM");

constexpr int kEpiThreads = 2 * kRows;       // two threads per env row: column halves
constexpr int kMmaWarp = kEpiThreads / 32;
constexpr int kThreads = kEpiThreads + 32;

struct Args {
    const uint8_t* blobs;      // [A] packed parameter blobs (flock_gru_tc_pack)
    const float* obs;          // [E][A][n_obs]
    const float* hidden_in;    // [E][A][32]
    float* hidden_out;         // [E][A][32] (may alias hidden_in)
    float* q_out;              // vdn: [E][A][n_act] or null
    float* actions;            // vdn: [E][A] float-coded ids or null
    int E, A, n_obs, n_act, env_offset, tiles, items_base, items_rem;   // CTA b: items_base (+ 1 if b < items_rem) items
    float epsilon;
    uint32_t seed_lo, seed_hi, step;
    const int32_t* env_step;
    const uint32_t* env_epoch;
    long long* dbg;            // FLOCK_GRU_TIMING=1: 32 clock64 slots per CTA (phases of the CTA's first two items)
};

// v = hi + lo with hi = fp16(v), lo = fp16(v - hi): 22 significant bits in two tensor-core operands
__device__ __forceinline__ uint32_t split_f16x2(float v0, float v1, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);          // one packed conversion (F2FP) per pair, not two F2F
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
    lo = *reinterpret_cast<const uint32_t*>(&l);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// eight consecutive K elements of one row -> one 16-byte core-matrix row in the hi image and one in the lo image
__device__ __forceinline__ void store_k8(uint32_t hi_addr, uint32_t lo_addr, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = split_f16x2(v[2 * i], v[2 * i + 1], l[i]);
    sts128(hi_addr, h[0], h[1], h[2], h[3]);
    sts128(lo_addr, l[0], l[1], l[2], l[3]);
}
// gate non-linearities on the MUFU path: ex2.approx + rcp (relative error ~1e-6, far below the operand split's 4e-6);
// tanh x = 1 - 2 / (1 + e^{2x}), which saturates correctly at +-1 for large |x|
__device__ __forceinline__ float rcp_(float x) {      // MUFU.RCP alone (1 ulp); __frcp_rn is a ~10-instruction IEEE sequence
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_(float x) { return rcp_(1.0f + ex2_(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_(float x) { return fmaf(-2.0f, rcp_(1.0f + ex2_(2.8853900817779268f * x)), 1.0f); }

// instruction descriptor as in flock_tc.cuh, with FP16 operands (A / B format fields [7,10) / [10,13) = 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[128 x N] (+)= A[128 x 16*ksteps] . B[16*ksteps x N] with split operands: hi*hi + hi*lo + lo*hi per K step (lo*lo is
// 2^-22 of the product, below the 2^-21 the split itself keeps)
__device__ __forceinline__ void mma_split(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int ksteps,
                                          int N, bool leader) {
    if (!leader) return;
    const uint32_t idesc = umma_idesc_f16(kRows, N);
    const uint32_t a_step = 2 * kRows * 16, b_step = 2 * (uint32_t)N * 16;
    for (int s = 0; s < ksteps; ++s) {
        const uint64_t dah = umma_desc(a_hi + s * a_step, kRows * 16, 128), dal = umma_desc(a_lo + s * a_step, kRows * 16, 128);
        const uint64_t dbh = umma_desc(b_hi + s * b_step, (uint32_t)N * 16, 128), dbl = umma_desc(b_lo + s * b_step, (uint32_t)N * 16, 128);
        umma_bf16(tmem_d, dah, dbh, idesc, s == 0 ? 0u : 1u);
        umma_bf16(tmem_d, dah, dbl, idesc, 1u);
        umma_bf16(tmem_d, dal, dbh, idesc, 1u);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) flock_gru_tc_kernel(const __grid_constant__ Args a) {
    using L = Layout<MODE>;
    constexpr bool kVdn = L::kVdn;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;
    uint8_t* sm = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* dbg_cta = a.dbg != nullptr ? a.dbg + (size_t)blockIdx.x * 32 : nullptr;
    auto stamp = [&](int it, int slot) {        // thread 0 only, first two items of the CTA
        if (dbg_cta != nullptr && threadIdx.x == 0 && it < 2) dbg_cta[it * 16 + slot] = clock64();
    };
    stamp(0, 0);
    const int total_items = a.A * a.tiles;
    // contiguous runs, the longer ones first: with two CTAs per SM placed round-robin, CTA b shares its SM with CTA
    // b + #SMs, so a long run is paired with a short one
    const int item0 = blockIdx.x * a.items_base + min((int)blockIdx.x, a.items_rem);
    const int item1 = min(item0 + a.items_base + ((int)blockIdx.x < a.items_rem ? 1 : 0), total_items);

    const uint32_t sBlob = base + L::kOffBlob;
    const uint32_t sWhi = sBlob, sWlo = sBlob + L::kPiece;
    const float* w0f = reinterpret_cast<const float*>(sm + L::kOffBlob + L::kOffW0f);
    const float* par = reinterpret_cast<const float*>(sm + L::kOffBlob + L::kOffPar);
    const uint32_t sX1h = base + L::kOffX1, sX1l = sX1h + L::kAx1;
    const uint32_t sHh = base + L::kOffH, sHl = sHh + L::kAh;
    const uint32_t sBar = base + L::kOffBar;
    const uint32_t bar_w = sBar, bar_a0 = sBar + 8, bar_m2 = sBar + 16, bar_ax = sBar + 24, bar_m3 = sBar + 32, bar_ah = sBar + 40,
                   bar_m4 = sBar + 48, bar_done = sBar + 56;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L::kOffBar + L::kNumBars * 8);

    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(L::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (lane == 0) {
            mbar_init(bar_w, 1);
            mbar_init(bar_a0, kEpiThreads);
            mbar_init(bar_m2, 1);
            mbar_init(bar_ax, kEpiThreads);
            mbar_init(bar_m3, 1);
            mbar_init(bar_ah, kEpiThreads);
            mbar_init(bar_m4, 1);
            mbar_init(bar_done, kEpiThreads);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // launched as a programmatic dependent (launch_pdl): TMEM allocation and barrier set-up above overlap the previous
    // kernel of the stream (usually the env step); nothing an earlier kernel wrote is read before this point
    pdl_wait_prior_grid();

    if (warp == kMmaWarp) {
        // ---- weight loader + MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const bool leader = elect_one();
        int prev_agent = -1;
        uint32_t w_loads = 0;
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / a.tiles;
            const uint32_t ph = (uint32_t)it & 1u;
            if (agent != prev_agent) {
                // the previous item's epilogue still reads the old parameters, its MMAs the old weights
                if (it > 0) mbar_wait(bar_done, (uint32_t)(it - 1) & 1u);
                if (leader) {
                    mbar_expect_tx(bar_w, L::kBlobBytes);
                    bulk_g2s(sBlob, a.blobs + (size_t)agent * L::kBlobBytes, L::kBlobBytes, bar_w);
                }
                __syncwarp();
                mbar_wait(bar_w, w_loads & 1u);
                ++w_loads;
                prev_agent = agent;
            }
            // the staged row: observation-layer output (X1) and hidden state (H)
            mbar_wait(bar_a0, ph);
            tc_fence_after();
            if (kVdn) {
                mma_split(tmem + L::kColH2, sX1h, sX1l, sWhi + L::kOffW2, sWlo + L::kOffW2, kH1 / 16, kHx, leader);
                if (leader) umma_commit(bar_m2);
                // hidden half of the GRU under the layer-2 epilogue
                mma_split(tmem + L::kColGh, sHh, sHl, sWhi + L::kOffWhh, sWlo + L::kOffWhh, kHx / 16, kG, leader);
                __syncwarp();
                mbar_wait(bar_ax, ph);           // x = ReLU(layer 2) staged in place of H1
                tc_fence_after();
                mma_split(tmem + L::kColGi, sX1h, sX1l, sWhi + L::kOffWih, sWlo + L::kOffWih, kHx / 16, kG, leader);
            } else {
                mma_split(tmem + L::kColGh, sHh, sHl, sWhi + L::kOffWhh, sWlo + L::kOffWhh, kHx / 16, kG, leader);
                mma_split(tmem + L::kColGi, sX1h, sX1l, sWhi + L::kOffWih, sWlo + L::kOffWih, kHx / 16, kG, leader);
            }
            if (leader) umma_commit(bar_m3);          // gi and gh
            __syncwarp();
            if (kVdn) {
                mbar_wait(bar_ah, ph);
                tc_fence_after();
                mma_split(tmem + L::kColQ, sHh, sHl, sWhi + L::kOffWq, sWlo + L::kOffWq, kHx / 16, kActPad, leader);
                if (leader) umma_commit(bar_m4);
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue warps: two threads per env row (= TMEM lane); `half` selects the columns / K groups a thread owns.
        // A warp may only touch TMEM lanes 32 * (warp % 4) .. + 31, so warps w and w + 4 share a lane group.
        const int row = threadIdx.x & (kRows - 1);
        const int half = threadIdx.x >> 7;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int kN0h = L::kN0 / 2;       // observation-layer outputs per thread
        int prev_agent = -1;
        uint32_t w_loads = 0;
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / a.tiles, tile = item - agent * a.tiles;
            const uint32_t ph = (uint32_t)it & 1u;
            const int env = tile * kRows + row;
            const bool valid = env < a.E;
            const size_t ea = (size_t)(valid ? env : 0) * a.A + agent;
            stamp(it, 1);
            // stage 0: the row's observation (both halves read all of it) and this half's 16 hidden units
            float ob[kObsPad];
#pragma unroll
            for (int c = 0; c < kObsPad; ++c) ob[c] = (valid && c < a.n_obs) ? a.obs[ea * a.n_obs + c] : 0.0f;
            float hold[kHx / 2];
            {
                const float4* h4 = reinterpret_cast<const float4*>(a.hidden_in + ea * kHx + half * (kHx / 2));
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = p;
                    if (valid) {
                        p = h4[2 * g];
                        q = h4[2 * g + 1];
                    }
                    const float w[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) hold[g * 8 + i] = w[i];
                    const int kg = half * 2 + g;
                    store_k8(sHh + (kg * kRows + row) * 16, sHl + (kg * kRows + row) * 16, w);
                }
            }
            stamp(it, 2);
            if (agent != prev_agent) {        // this agent's weights and biases have landed
                mbar_wait(bar_w, w_loads & 1u);
                ++w_loads;
                prev_agent = agent;
            }
            stamp(it, 3);
            // observation layer on the CUDA cores, fp32: outputs half * kN0h .. + kN0h - 1 (+ ReLU for the VDN feature
            // layer), written as the next MMA's A operand
#pragma unroll 1
            for (int c0 = half * kN0h; c0 < (half + 1) * kN0h; c0 += 16) {
                float acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = par[L::kPb0 + c0 + i];
#pragma unroll
                for (int c = 0; c < kObsPad; ++c) {
                    if (c >= a.n_obs) break;             // warp-uniform; unrolled so that ob[] stays in registers
                    const float4* wr = reinterpret_cast<const float4*>(w0f + c * L::kN0 + c0);
                    const float o = ob[c];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 w = wr[i4];
                        acc[4 * i4] = fmaf(o, w.x, acc[4 * i4]);
                        acc[4 * i4 + 1] = fmaf(o, w.y, acc[4 * i4 + 1]);
                        acc[4 * i4 + 2] = fmaf(o, w.z, acc[4 * i4 + 2]);
                        acc[4 * i4 + 3] = fmaf(o, w.w, acc[4 * i4 + 3]);
                    }
                }
                float v0[8], v1[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v0[i] = kVdn ? fmaxf(acc[i], 0.0f) : acc[i];
                    v1[i] = kVdn ? fmaxf(acc[8 + i], 0.0f) : acc[8 + i];
                }
                const uint32_t o0 = (uint32_t)((c0 >> 3) * kRows + row) * 16, o1 = o0 + kRows * 16;
                store_k8(sX1h + o0, sX1l + o0, v0);
                store_k8(sX1h + o1, sX1l + o1, v1);
            }
            fence_proxy_async();
            mbar_arrive(bar_a0);
            stamp(it, 4);
            if (kVdn) {
                // second feature layer: bias + ReLU -> x (same place: the layer-2 MMA has consumed H1)
                mbar_wait(bar_m2, ph);
                tc_fence_after();
                stamp(it, 5);
                {
                    const int c0 = half * 16;
                    uint32_t r[16];
                    tmem_ld16_issue(trow + L::kColH2 + c0, r);
                    tmem_ld16_wait(r);
                    float v0[8], v1[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v0[i] = fmaxf(__uint_as_float(r[i]) + par[L::kPb2 + c0 + i], 0.0f);
                        v1[i] = fmaxf(__uint_as_float(r[8 + i]) + par[L::kPb2 + c0 + 8 + i], 0.0f);
                    }
                    const int kg = c0 >> 3;
                    store_k8(sX1h + (kg * kRows + row) * 16, sX1l + (kg * kRows + row) * 16, v0);
                    store_k8(sX1h + ((kg + 1) * kRows + row) * 16, sX1l + ((kg + 1) * kRows + row) * 16, v1);
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_ax);
                stamp(it, 6);
            }
            // GRU gates (torch.nn.GRUCell, gate order r | z | n): h' = (1 - z) n + z h, units 16 * half .. + 15
            mbar_wait(bar_m3, ph);
            tc_fence_after();
            stamp(it, 7);
            float hnew[kHx / 2];
            {
                const int u0 = half * 16;
                float rg[16], zg[16];
                uint32_t gi[16], gh[16];
                tmem_ld16_issue(trow + L::kColGi + u0, gi);
                tmem_ld16_issue(trow + L::kColGh + u0, gh);
                tmem_ld16_wait(gi);
                tmem_ld16_wait(gh);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    rg[i] = sigmoid_((__uint_as_float(gi[i]) + par[L::kPbih + u0 + i]) + (__uint_as_float(gh[i]) + par[L::kPbhh + u0 + i]));
                tmem_ld16_issue(trow + L::kColGi + kHx + u0, gi);
                tmem_ld16_issue(trow + L::kColGh + kHx + u0, gh);
                tmem_ld16_wait(gi);
                tmem_ld16_wait(gh);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    zg[i] = sigmoid_((__uint_as_float(gi[i]) + par[L::kPbih + kHx + u0 + i]) +
                                     (__uint_as_float(gh[i]) + par[L::kPbhh + kHx + u0 + i]));
                tmem_ld16_issue(trow + L::kColGi + 2 * kHx + u0, gi);
                tmem_ld16_issue(trow + L::kColGh + 2 * kHx + u0, gh);
                tmem_ld16_wait(gi);
                tmem_ld16_wait(gh);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float n = tanh_((__uint_as_float(gi[i]) + par[L::kPbih + 2 * kHx + u0 + i]) +
                                          rg[i] * (__uint_as_float(gh[i]) + par[L::kPbhh + 2 * kHx + u0 + i]));
                    hnew[i] = (1.0f - zg[i]) * n + zg[i] * hold[i];
                }
            }
            stamp(it, 8);
            if (valid && a.hidden_out != nullptr) {
                float4* o4 = reinterpret_cast<float4*>(a.hidden_out + ea * kHx + half * (kHx / 2));
#pragma unroll
                for (int j = 0; j < kHx / 8; ++j) o4[j] = make_float4(hnew[4 * j], hnew[4 * j + 1], hnew[4 * j + 2], hnew[4 * j + 3]);
            }
            if (kVdn) {
                // h' as the head's A operand (same place as h: the gh MMAs completed with m3)
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = hnew[g * 8 + i];
                    const int kg = half * 2 + g;
                    store_k8(sHh + (kg * kRows + row) * 16, sHl + (kg * kRows + row) * 16, v);
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_ah);
                stamp(it, 9);
                mbar_wait(bar_m4, ph);
                stamp(it, 10);     // every thread: the head MMA reads h', which the next item's stage 0 overwrites
                if (half == 0) {      // the head is 16 columns: one thread per row finishes it
                    tc_fence_after();
                    uint32_t qr[16];
                    tmem_ld16_issue(trow + L::kColQ, qr);
                    tmem_ld16_wait(qr);
                    float q[kActPad];
#pragma unroll
                    for (int j = 0; j < kActPad; ++j) q[j] = __uint_as_float(qr[j]) + par[L::kPbq + j];
                    if (valid) {
                        if (a.q_out != nullptr) {
#pragma unroll
                            for (int j = 0; j < kActPad; ++j)
                                if (j < a.n_act) a.q_out[ea * a.n_act + j] = q[j];
                        }
                        if (a.actions != nullptr) {
                            // greedy: first maximum, like torch.argmax; exploration: one decision per env (net.py:54), then a
                            // uniform action id per agent (net.py:56) -- same Philox stream as flock_qnet_kernel
                            int best = 0;
                            float bv = q[0];
#pragma unroll
                            for (int j = 1; j < kActPad; ++j)
                                if (j < a.n_act && q[j] > bv) {
                                    bv = q[j];
                                    best = j;
                                }
                            const uint32_t ge = (uint32_t)(a.env_offset + env);
                            const uint32_t c2 = a.step + (a.env_step != nullptr ? (uint32_t)a.env_step[env] : 0u);
                            const uint32_t ep = a.env_epoch != nullptr ? (a.env_epoch[env] << 4) : 0u;
                            const uint4 re = philox4x32_10(ge, 0xffffffffu, c2, kTagExplore + ep, a.seed_lo, a.seed_hi);
                            if (u24(re.x) <= a.epsilon && a.epsilon > 0.0f) {
                                const uint4 ra = philox4x32_10(ge, (uint32_t)agent, c2, kTagRandAct + ep, a.seed_lo, a.seed_hi);
                                best = (int)(((unsigned long long)ra.x * (unsigned long long)a.n_act) >> 32);
                            }
                            a.actions[ea] = (float)best;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_done);
            stamp(it, 11);
        }
    }
    pdl_launch_dependents();     // the next kernel of the stream (recurrent actor: the MLP kernel; the env step) may start launching
    tc_fence_before();
    __syncthreads();
    stamp(1, 12);
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Layout<MODE>::kTmemCols) : "memory");
    }
}

// ---- pack: fp32 parameters (weights [A][in][out], as policies.py keeps them) -> hi / lo fp16 images + fp32 weights of the observation layer + fp32 biases ----
struct PackArgs {
    const float *w0, *b0, *w2, *b2, *w_ih, *b_ih, *w_hh, *b_hh, *wq, *bq;
    int n_obs, n_act;
};

__device__ __forceinline__ void pack_unit(uint8_t* hi_img, uint8_t* lo_img, const float* W, int K, int N, int Npad, int u) {
    // unit u = 16 bytes = 8 consecutive K elements of column n: offset ((s * 2 + g) * Npad + n) * 16
    const int n = u % Npad, sg = u / Npad;
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = sg * 8 + 2 * i + j;
            v[j] = (k < K && n < N) ? W[(size_t)k * N + n] : 0.0f;
        }
        h[i] = split_f16x2(v[0], v[1], l[i]);
    }
    reinterpret_cast<uint4*>(hi_img)[u] = make_uint4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<uint4*>(lo_img)[u] = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int MODE>
__global__ void flock_gru_tc_pack_kernel(PackArgs p, uint8_t* __restrict__ blobs) {
    using L = Layout<MODE>;
    const int agent = blockIdx.x;
    uint8_t* blob = blobs + (size_t)agent * L::kBlobBytes;
    uint8_t *hi = blob, *lo = blob + L::kPiece;
    auto pack = [&](int off, const float* W, int K, int Kpad, int N, int Npad) {
        const int units = (Kpad / 8) * Npad;
        for (int u = threadIdx.x; u < units; u += blockDim.x) pack_unit(hi + off, lo + off, W, K, N, Npad, u);
    };
    if (L::kVdn) pack(L::kOffW2, p.w2 + (size_t)agent * kH1 * kHx, kH1, kH1, kHx, kHx);
    pack(L::kOffWih, p.w_ih + (size_t)agent * kHx * kG, kHx, kHx, kG, kG);
    pack(L::kOffWhh, p.w_hh + (size_t)agent * kHx * kG, kHx, kHx, kG, kG);
    if (L::kVdn) pack(L::kOffWq, p.wq + (size_t)agent * kHx * p.n_act, kHx, kHx, p.n_act, kActPad);
    float* w0f = reinterpret_cast<float*>(blob + L::kOffW0f);      // observation layer: fp32, rows beyond n_obs zero
    for (int i = threadIdx.x; i < kObsPad * L::kN0; i += blockDim.x)
        w0f[i] = (i / L::kN0) < p.n_obs ? p.w0[(size_t)agent * p.n_obs * L::kN0 + i] : 0.0f;
    float* par = reinterpret_cast<float*>(blob + L::kOffPar);
    for (int i = threadIdx.x; i < L::kParamFloats; i += blockDim.x) {
        float v;
        if (i < L::kPb2) v = p.b0[(size_t)agent * L::kN0 + i];
        else if (i < L::kPbih) v = p.b2[(size_t)agent * kHx + (i - L::kPb2)];
        else if (i < L::kPbhh) v = p.b_ih[(size_t)agent * kG + (i - L::kPbih)];
        else if (i < L::kPbq) v = p.b_hh[(size_t)agent * kG + (i - L::kPbhh)];
        else v = (i - L::kPbq) < p.n_act ? p.bq[(size_t)agent * p.n_act + (i - L::kPbq)] : 0.0f;
        par[i] = v;
    }
}

}  // namespace grutc

size_t gru_tc_blob_bytes(int mode) { return mode == 1 ? grutc::Layout<1>::kBlobBytes : grutc::Layout<0>::kBlobBytes; }
int gru_tc_max_obs() { return grutc::kObsPad; }
int gru_tc_max_actions() { return grutc::kActPad; }

cudaError_t launch_gru_tc_pack(int mode, int agents, int n_obs, int n_act, const float* const* p, void* blobs, cudaStream_t s) {
    grutc::PackArgs a;
    if (mode == 1) {   // {w1, b1, w2, b2, wq, bq, w_ih, b_ih, w_hh, b_hh}: the parameter order of flock_qnet_forward
        a.w0 = p[0]; a.b0 = p[1]; a.w2 = p[2]; a.b2 = p[3]; a.wq = p[4]; a.bq = p[5];
        a.w_ih = p[6]; a.b_ih = p[7]; a.w_hh = p[8]; a.b_hh = p[9];
    } else {           // {w_e, b_e, w_ih, b_ih, w_hh, b_hh}: the front parameters of flock_rnn_actor_forward
        a.w0 = p[0]; a.b0 = p[1]; a.w_ih = p[2]; a.b_ih = p[3]; a.w_hh = p[4]; a.b_hh = p[5];
        a.w2 = a.b2 = a.wq = a.bq = nullptr;
    }
    a.n_obs = n_obs;
    a.n_act = n_act;
    if (mode == 1) grutc::flock_gru_tc_pack_kernel<1><<<agents, 256, 0, s>>>(a, static_cast<uint8_t*>(blobs));
    else grutc::flock_gru_tc_pack_kernel<0><<<agents, 256, 0, s>>>(a, static_cast<uint8_t*>(blobs));
    return cudaGetLastError();
}

cudaError_t launch_gru_tc_forward(int mode, const void* blobs, const float* obs, const float* hidden_in, float* hidden_out,
                                  float* q_out, float* actions, int E, int A, int n_obs, int n_act, float epsilon, uint64_t seed,
                                  uint32_t step, int env_offset, NoiseCounters ctr, cudaStream_t s) {
    grutc::Args a;
    a.blobs = static_cast<const uint8_t*>(blobs);
    a.obs = obs; a.hidden_in = hidden_in; a.hidden_out = hidden_out; a.q_out = q_out; a.actions = actions;
    a.E = E; a.A = A; a.n_obs = n_obs; a.n_act = n_act; a.env_offset = env_offset;
    a.epsilon = epsilon;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.step = step;
    a.env_step = ctr.env_step; a.env_epoch = ctr.env_epoch;
    a.tiles = (E + grutc::kRows - 1) / grutc::kRows;
    static DeviceOnce once;
    int sm_count = 148;
    const cudaError_t cfg = once.get(
        [] {
            cudaError_t e = cudaFuncSetAttribute(grutc::flock_gru_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 grutc::Layout<0>::kSmemBytes);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(grutc::flock_gru_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         grutc::Layout<1>::kSmemBytes);
            return e;
        },
        &sm_count);
    if (cfg != cudaSuccess) return cfg;
    const int total = a.tiles * A;
    const int slots = 2 * sm_count;                                  // two resident CTAs per SM
    const int grid = total < slots ? total : slots;
    a.items_base = total / grid;
    a.items_rem = total % grid;
    a.dbg = nullptr;
    // FLOCK_GRU_TIMING=1 (developer knob): the first launch records clock64 at the phase boundaries of the first two items
    // of every CTA (epilogue thread 0), synchronises and prints the mean phase lengths to stderr.
    static bool timing = getenv("FLOCK_GRU_TIMING") != nullptr;
    long long* dbg = nullptr;
    if (timing && cudaMalloc(&dbg, (size_t)grid * 32 * sizeof(long long)) == cudaSuccess) {
        cudaMemsetAsync(dbg, 0, (size_t)grid * 32 * sizeof(long long), s);
        a.dbg = dbg;
    }
    const cudaError_t launched =
        mode == 1 ? launch_pdl(grutc::flock_gru_tc_kernel<1>, dim3(grid), dim3(grutc::kThreads), grutc::Layout<1>::kSmemBytes, s, a)
                  : launch_pdl(grutc::flock_gru_tc_kernel<0>, dim3(grid), dim3(grutc::kThreads), grutc::Layout<0>::kSmemBytes, s, a);
    if (dbg != nullptr) {
        timing = false;
        cudaStreamSynchronize(s);
        long long* h = static_cast<long long*>(malloc((size_t)grid * 32 * sizeof(long long)));
        cudaMemcpy(h, dbg, (size_t)grid * 32 * sizeof(long long), cudaMemcpyDeviceToHost);
        static const char* names[12] = {"setup (alloc, barrier init, sync)", "global loads + h split", "wait weights", "obs layer (fp32) -> X1",
                                        "wait MMA layer 2", "layer-2 epilogue -> X", "wait MMA gates", "gates", "hidden store + h' split",
                                        "wait MMA head", "head epilogue", "item total"};
        static const int from[12] = {0, 1, 2, 3, 4, 5, 4, 7, 8, 9, 10, 1}, to[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 11};
        fprintf(stderr, "flock_gru_tc_kernel<%d>: %d CTAs x %d(+1) items, phase means (SM clocks), first item | second item:\n", mode, grid,
                a.items_base);
        for (int ph = 0; ph < 12; ++ph) {
            double acc[2] = {0, 0};
            int cnt[2] = {0, 0};
            for (int it = 0; it < 2; ++it)
                for (int c = 0; c < grid; ++c) {
                    const long long t1 = h[c * 32 + it * 16 + to[ph]], t0 = h[c * 32 + it * 16 + from[ph]];
                    if (t1 != 0 && t0 != 0) {
                        acc[it] += (double)(t1 - t0);
                        cnt[it] += 1;
                    }
                }
            fprintf(stderr, "  %-36s %9.0f | %9.0f\n", names[ph], cnt[0] ? acc[0] / cnt[0] : 0.0, cnt[1] ? acc[1] / cnt[1] : 0.0);
        }
        double tot = 0;
        for (int c = 0; c < grid; ++c) tot += (double)(h[c * 32 + 16 + 12] - h[c * 32]);
        fprintf(stderr, "  %-36s %9.0f\n", "CTA total (all items)", tot / grid);
        free(h);
        cudaFree(dbg);
    }
    return launched;
}

}  // namespace flock
