// flock_rnn_actor.cu -- fused recurrent MADDPG actor (the policy of the reference's default main.py loop):
// `Actor` of learners/maddpg_official_rnn/net.py:14-72, one weight set per agent,
//     fce(in -> 32) -> GRUCell(32, 32) -> fc1(32 -> 400) -> ReLU -> fc2(400 -> 300) -> ReLU ->
//     [linear_speed(300 -> 1): (tanh + 1) / 2 | angular_speed(300 -> 1): 1.5 tanh]
// evaluated for all envs and agents in TWO launches:
//   1. flock_rnn_front_kernel: fce + GRUCell in fp32 on CUDA cores (the recurrent state stays exact fp32; same
//      structure as the GRU of flock_qnet.cu), writes the new hidden state -- which is also the MLP's input;
//   2. flock_rnn_mlp_kernel: the 32-400-300-2 MLP on the tensor cores with the machinery of flock_actor.cu
//      (persistent CTAs, tcgen05.mma with hand-built descriptors, weights pre-packed into shared-memory images and
//      streamed by 1-D TMA bulk copies, accumulators in TMEM, two-phase epilogue 1 under the first layer-2 MMAs).
//      No LayerNorm here, so the epilogues are one pass each: bias + ReLU -> bf16 (layer 1; the layer-2 bias rides
//      in the MMA as hi + lo bf16 rows), ReLU + the two heads (layer 2).
#include <cmath>
#include <cstdlib>

#include "flock_device.cuh"
#include "flock_launch.h"
#include "flock_tc.cuh"

namespace flock {
namespace rnn {

using namespace tc;

constexpr int kRows = 128;                   // env rows per work item = UMMA M
constexpr int kHid = 32;                     // hidden_rnn: GRU state = layer-1 K (two UMMA K steps)
constexpr int kK1Steps = kHid / 16;
constexpr int kFc1 = 400, kFc2 = 300, kFc2Pad = 304;
constexpr int kK2 = kFc1 + 16;               // layer-2 K: 400 activations + one K step whose first two slots are the bias
constexpr int kSteps2 = kK2 / 16;
constexpr int kW1StepBytes = 2 * kFc1 * 16;  // one K step of W1: 2 k-groups x 400 rows x 16 B
constexpr int kW1Bytes = kK1Steps * kW1StepBytes;
constexpr int kStepBytes = 2 * kFc2Pad * 16; // one K step of W2
constexpr int kStepsPerChunk = 2;
constexpr int kChunks = kSteps2 / kStepsPerChunk;
constexpr int kChunkBytes = kStepsPerChunk * kStepBytes;
constexpr int kStages = 4;
// fp32 parameters: b1 [400] | w_linear [304] | w_angular [304] | b_linear, b_angular + 2 pad
constexpr int kParamFloats = kFc1 + 2 * kFc2Pad + 4;
constexpr int kParamBytes = kParamFloats * 4;
constexpr int kBlobBytes = kW1Bytes + kParamBytes + kSteps2 * kStepBytes;
static_assert(kParamBytes % 16 == 0 && kBlobBytes % 16 == 0, "bulk copies move 16-byte units");

constexpr int kOffA2 = 0;
constexpr int kA2Bytes = (kK2 / 8) * kRows * 16;
constexpr int kOffA1 = kOffA2 + kA2Bytes;                  // layer-1 A: 4 k-groups x 128 rows x 16 B
constexpr int kA1Bytes = 2 * kK1Steps * kRows * 16;
constexpr int kOffW1 = kOffA1 + kA1Bytes;
constexpr int kOffPar = kOffW1 + kW1Bytes;
constexpr int kOffRing = (kOffPar + kParamBytes + 127) & ~127;
constexpr int kOffBar = kOffRing + kStages * kChunkBytes;
constexpr int kNumBars = 2 * kStages + 7;
constexpr int kOffRed = kOffBar + kNumBars * 8 + 16;
constexpr int kColGroups = 4;
constexpr int kRedBytes = kColGroups * kRows * 8;          // float2 [4][128] head partials
constexpr int kSmemBytes = kOffRed + kRedBytes + 128;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr int kEpiWarps = 4 * kColGroups;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kTmaWarp = kEpiWarps + 1;
constexpr int kThreads = kEpiThreads + 64;
constexpr int kTmemCols = 512;
constexpr int kL2Col = 208;                  // layer-1 accumulators [0, 400), layer-2 accumulators [208, 512): see flock_actor.cu
constexpr int kPhaseAUnit = 12;
constexpr int kFirstChunk = kPhaseAUnit / kStepsPerChunk;

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ int unit_begin(int units, int g) { return (units * g + kColGroups - 1) / kColGroups; }

// Optional exploration noise fused into the output stage (mu' = mu + noise.sample(), agent.py:61): one
// Ornstein-Uhlenbeck process per (env, agent, action) -- OrnsteinUhlenbeckProcess.sample,
// learners/maddpg_official_rnn/utils.py:43-47 -- x <- x + theta (mu - x) dt + sigma sqrt(dt) N(0, 1), with Philox
// normals (counter = env_offset + env, agent, step, tag 7; key = seed). The reference shares ONE process object
// between all agents of its single env; the batched form keeps them independent.
struct OuArgs {
    float2* state;            // [E][A] (x0, x1), updated in place; nullptr = no noise
    float theta_dt, mu, sigma_sqrt_dt;
    uint32_t seed_lo, seed_hi, step;
    int env_offset;
    const int32_t* env_step;      // [E] device step counters added to `step` (nullable; replay-safe under CUDA graphs)
    const uint32_t* env_epoch;    // [E] device epoch counters folded into the tag word (nullable)
};
constexpr uint32_t kTagOu = 7u;

// hin [E][N][32] fp32 (the GRU state the front kernel just wrote), out [E][N][2] fp32
__global__ void __launch_bounds__(kThreads, 1)
flock_rnn_mlp_kernel(const uint8_t* __restrict__ blobs, const float* __restrict__ hin, float* __restrict__ out, int E, int N,
                     int tiles, int items_per_cta, OuArgs ou) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;
    uint8_t* sm = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_items = N * tiles;
    const int item0 = blockIdx.x * items_per_cta;
    const int item1 = min(item0 + items_per_cta, total_items);

    const uint32_t sA2 = base + kOffA2, sA1 = base + kOffA1, sW1 = base + kOffW1, sPar = base + kOffPar;
    const uint32_t sRing = base + kOffRing, sBar = base + kOffBar;
    auto bar_full = [&](int s) { return sBar + 8u * s; };
    auto bar_empty = [&](int s) { return sBar + 8u * (kStages + s); };
    const uint32_t bar_w1 = sBar + 8u * (2 * kStages), bar_a1 = bar_w1 + 8u, bar_mma1 = bar_w1 + 16u, bar_a2 = bar_w1 + 24u,
                   bar_mma2 = bar_w1 + 32u, bar_done = bar_w1 + 40u, bar_a2b = bar_w1 + 48u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + kNumBars * 8);
    const float* par = reinterpret_cast<const float*>(sm + kOffPar);
    float2* red_head = reinterpret_cast<float2*>(sm + kOffRed);

    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp == kTmaWarp && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_w1, 1);
        mbar_init(bar_a1, kRows);
        mbar_init(bar_mma1, 1);
        mbar_init(bar_a2, kEpiThreads);
        mbar_init(bar_mma2, 1);
        mbar_init(bar_done, kEpiThreads);
        mbar_init(bar_a2b, kEpiThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // launched as a programmatic dependent (launch_pdl): TMEM allocation and barrier set-up above overlap the previous
    // kernel of the stream (usually the env step); nothing an earlier kernel wrote is read before this point
    pdl_wait_prior_grid();

    // roles, barriers and parities exactly as in flock_actor_kernel
    if (warp == kTmaWarp) {
        const bool leader = elect_one();
        int prev_agent = -1;
        uint32_t g = 0;
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / tiles;
            const uint8_t* blob = blobs + (size_t)agent * kBlobBytes;
            if (agent != prev_agent) {
                if (it > 0) mbar_wait(bar_done, (uint32_t)(it - 1) & 1u);   // the epilogues still read the old parameters
                if (leader) {
                    mbar_expect_tx(bar_w1, kW1Bytes + kParamBytes);
                    bulk_g2s(sW1, blob, kW1Bytes, bar_w1);
                    bulk_g2s(sPar, blob + kW1Bytes, kParamBytes, bar_w1);
                }
                prev_agent = agent;
            }
            const uint8_t* w2 = blob + kW1Bytes + kParamBytes;
#pragma unroll 1
            for (int idx = 0; idx < kChunks; ++idx, ++g) {
                const int c = (idx + kFirstChunk) % kChunks;
                const uint32_t slot = g % kStages;
                if (g >= kStages) mbar_wait(bar_empty(slot), (g / kStages - 1u) & 1u);
                if (leader) {
                    mbar_expect_tx(bar_full(slot), kChunkBytes);
                    bulk_g2s(sRing + slot * kChunkBytes, w2 + (size_t)c * kChunkBytes, kChunkBytes, bar_full(slot));
                }
                __syncwarp();
            }
        }
    } else if (warp == kMmaWarp) {
        const bool leader = elect_one();
        int prev_agent = -1;
        uint32_t w1_loads = 0, g = 0;
        const uint64_t da1 = umma_desc(sA1, kRows * 16, 128);
        const uint64_t db1 = umma_desc(sW1, kFc1 * 16, 128);
        const uint64_t da0 = umma_desc(sA2, kRows * 16, 128);
        const uint64_t db0 = umma_desc(sRing, kFc2Pad * 16, 128);
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / tiles;
            const uint32_t ph = (uint32_t)it & 1u;
            // layer 1: [128 x 32] x [32 x 400] -> TMEM columns [0, 400); the previous item's epilogue 2 still reads
            // columns [208, 512): wait until every epilogue thread has left it
            if (it > 0) mbar_wait(bar_done, (uint32_t)(it - 1) & 1u);
            mbar_wait(bar_a1, ph);
            if (agent != prev_agent) {
                mbar_wait(bar_w1, w1_loads & 1u);
                ++w1_loads;
                prev_agent = agent;
            }
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int s = 0; s < kK1Steps; ++s) {
                    const uint64_t da = da1 + (uint64_t)(s * ((2 * kRows * 16) >> 4));
                    const uint64_t db = db1 + (uint64_t)(s * (kW1StepBytes >> 4));
                    umma_bf16(tmem + 0, da, db, umma_idesc(kRows, 256), s > 0 ? 1u : 0u);
                    umma_bf16(tmem + 256, da, db + (256 * 16 >> 4), umma_idesc(kRows, kFc1 - 256), s > 0 ? 1u : 0u);
                }
                umma_commit(bar_mma1);
            }
            __syncwarp();
            // layer 2: K steps 12..25 first (their A operand is phase A of epilogue 1), then 0..11
            mbar_wait(bar_a2, ph);
            tc_fence_after();
#pragma unroll
            for (int idx = 0; idx < kChunks; ++idx, ++g) {
                constexpr int kAStep = (2 * kRows * 16) >> 4, kBSlot = kChunkBytes >> 4, kBStep = kStepBytes >> 4;
                const int c = (idx + kFirstChunk) % kChunks;
                if (c == 0) {
                    mbar_wait(bar_a2b, ph);
                    tc_fence_after();
                }
                const uint32_t st = g % kStages;
                mbar_wait(bar_full(st), (g / kStages) & 1u);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int j = 0; j < kStepsPerChunk; ++j) {
                        const int s = c * kStepsPerChunk + j;
                        const uint64_t da = da0 + (uint64_t)(s * kAStep);
                        const uint64_t db = db0 + (uint64_t)(st * kBSlot + j * kBStep);
                        const uint32_t acc = (idx > 0 || j > 0) ? 1u : 0u;
                        umma_bf16(tmem + kL2Col, da, db, umma_idesc(kRows, 256), acc);
                        umma_bf16(tmem + kL2Col + 256, da, db + (256 * 16 >> 4), umma_idesc(kRows, kFc2Pad - 256), acc);
                    }
                    umma_commit(bar_empty(st));
                }
                __syncwarp();
            }
            if (leader) umma_commit(bar_mma2);
            __syncwarp();
        }
    } else {
        const int q = warp & 3, cg = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t trow2 = trow + kL2Col;
        const float* b1 = par;
        const float* wl = par + kFc1;
        const float* wa = wl + kFc2Pad;
        const float* bh = wa + kFc2Pad;                    // (b_linear, b_angular)
        const int cab = (kPhaseAUnit + unit_begin(kFc1 / 16 - kPhaseAUnit, cg)) * 16,
                  cae = (kPhaseAUnit + unit_begin(kFc1 / 16 - kPhaseAUnit, cg + 1)) * 16;
        const int cbb = unit_begin(kPhaseAUnit, cg) * 16, cbe = unit_begin(kPhaseAUnit, cg + 1) * 16;
        const int c2b = unit_begin(kFc2Pad / 16, cg) * 16, c2e = unit_begin(kFc2Pad / 16, cg + 1) * 16;
        int prev_agent = -1;
        uint32_t w1_loads = 0;
        if (cg == 1) {   // the bias K step of layer 2 never changes: A2[:, 400] = A2[:, 401] = 1, A2[:, 402..415] = 0
            sts128(sA2 + (kFc1 / 8) * (kRows * 16) + row * 16, pack_bf16(1.0f, 1.0f), 0u, 0u, 0u);
            sts128(sA2 + (kFc1 / 8 + 1) * (kRows * 16) + row * 16, 0u, 0u, 0u, 0u);
        }
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / tiles, tile = item - agent * tiles;
            const uint32_t ph = (uint32_t)it & 1u;
            const int env = tile * kRows + row;
            const bool valid = env < E;
            if (cg == 0) {   // this row's hidden state = the layer-1 A operand (four 8-element K groups)
                const float4* s4 = reinterpret_cast<const float4*>(hin + ((size_t)(valid ? env : 0) * N + agent) * kHid);
#pragma unroll
                for (int kg = 0; kg < kHid / 8; ++kg) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                    if (valid) {
                        a = s4[2 * kg];
                        b = s4[2 * kg + 1];
                    }
                    sts128(sA1 + kg * (kRows * 16) + row * 16, pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y),
                           pack_bf16(b.z, b.w));
                }
                fence_proxy_async();
                mbar_arrive(bar_a1);
            }
            if (agent != prev_agent) {
                mbar_wait(bar_w1, w1_loads & 1u);
                ++w1_loads;
                prev_agent = agent;
            }
            // ---- epilogue 1: bias + ReLU -> bf16 A operand of layer 2, phase A then phase B ----
            mbar_wait(bar_mma1, ph);
            tc_fence_after();
            auto relu_unit = [&](int c0, const uint32_t (&r)[16]) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(r[i]) + b1[c0 + i], 0.0f);
                const uint32_t dst = sA2 + (c0 >> 3) * (kRows * 16) + row * 16;
                sts128(dst, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                sts128(dst + kRows * 16, pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                       pack_bf16(v[14], v[15]));
            };
            for_each_unit(trow, cab, cae, relu_unit);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_a2);
            for_each_unit(trow, cbb, cbe, relu_unit);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_a2b);
            // ---- epilogue 2: ReLU + the two heads (the bias came with the MMA; padded columns have zero head weights) ----
            mbar_wait(bar_mma2, ph);
            tc_fence_after();
            float o0 = 0.0f, o1 = 0.0f;
            for_each_unit(trow2, c2b, c2e, [&](int c0, const uint32_t (&r)[16]) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float yv = fmaxf(__uint_as_float(r[i]), 0.0f);
                    o0 = fmaf(yv, wl[c0 + i], o0);
                    o1 = fmaf(yv, wa[c0 + i], o1);
                }
            });
            red_head[cg * kRows + row] = make_float2(o0, o1);
            epi_sync();
            if (cg == 0 && valid) {
                o0 = 0.0f;
                o1 = 0.0f;
#pragma unroll
                for (int g = 0; g < kColGroups; ++g) {
                    const float2 pr = red_head[g * kRows + row];
                    o0 += pr.x;
                    o1 += pr.y;
                }
                float2 a;
                a.x = (tanhf(o0 + bh[0]) + 1.0f) * 0.5f;      // net.py:66-67
                a.y = tanhf(o1 + bh[1]) * 1.5f;              // net.py:70-71
                if (ou.state != nullptr) {
                    float2 x = ou.state[(size_t)env * N + agent];
                    const uint32_t c2 = ou.step + (ou.env_step != nullptr ? (uint32_t)ou.env_step[env] : 0u);
                    const uint32_t c3 = kTagOu + (ou.env_epoch != nullptr ? (ou.env_epoch[env] << 4) : 0u);
                    const uint4 rnd = philox4x32_10((uint32_t)(ou.env_offset + env), (uint32_t)agent, c2, c3, ou.seed_lo,
                                                    ou.seed_hi);
                    float z0, z1;
                    normal2(rnd.x, rnd.y, z0, z1);
                    x.x = x.x + ou.theta_dt * (ou.mu - x.x) + ou.sigma_sqrt_dt * z0;
                    x.y = x.y + ou.theta_dt * (ou.mu - x.y) + ou.sigma_sqrt_dt * z1;
                    ou.state[(size_t)env * N + agent] = x;
                    a.x += x.x;
                    a.y += x.y;
                }
                reinterpret_cast<float2*>(out)[(size_t)env * N + agent] = a;
            }
            tc_fence_before();
            mbar_arrive(bar_done);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ---- pack: fp32 parameters (layout of policies.BatchedRnnActors, weights [A][in][out]) -> shared-memory images ----
struct PackArgs {
    const float *w1, *b1, *w2, *b2, *wl, *bl, *wa, *ba;
    int agents;
};

__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void flock_rnn_pack_kernel(PackArgs a, uint8_t* __restrict__ blobs) {
    const int units = kBlobBytes / 16;
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (gid >= (size_t)units * a.agents) return;
    const int ag = (int)(gid / units);
    int u = (int)(gid % units);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (u < kW1Bytes / 16) {                       // W1 images: [K step][k-group][n] x 8 bf16
        const int s = u / (2 * kFc1), rem = u % (2 * kFc1);
        const int kg = rem / kFc1, n = rem % kFc1;
        float v[8];
        for (int j = 0; j < 8; ++j) v[j] = a.w1[((size_t)ag * kHid + (s * 16 + kg * 8 + j)) * kFc1 + n];
        o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    } else if (u < (kW1Bytes + kParamBytes) / 16) {   // fp32 parameters
        const int f0 = (u - kW1Bytes / 16) * 4;
        float v[4];
        for (int j = 0; j < 4; ++j) {
            int f = f0 + j;
            float x = 0.0f;
            if (f < kFc1) {
                x = a.b1[(size_t)ag * kFc1 + f];
            } else if (f < kFc1 + 2 * kFc2Pad) {
                f -= kFc1;
                const int which = f / kFc2Pad, c = f % kFc2Pad;
                x = c < kFc2 ? (which == 0 ? a.wl : a.wa)[(size_t)ag * kFc2 + c] : 0.0f;
            } else {
                f -= kFc1 + 2 * kFc2Pad;
                x = f == 0 ? a.bl[ag] : (f == 1 ? a.ba[ag] : 0.0f);
            }
            v[j] = x;
        }
        o = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    } else {                                        // W2 images: [K step][k-group][n] x 8 bf16; K slots 400, 401 = bias hi, lo
        u -= (kW1Bytes + kParamBytes) / 16;
        const int s = u / (2 * kFc2Pad), rem = u % (2 * kFc2Pad);
        const int kg = rem / kFc2Pad, n = rem % kFc2Pad;
        float v[8];
        for (int j = 0; j < 8; ++j) {
            const int k = s * 16 + kg * 8 + j;
            float x = 0.0f;
            if (n < kFc2) {
                const float bias = a.b2[(size_t)ag * kFc2 + n];
                if (k < kFc1) x = a.w2[((size_t)ag * kFc1 + k) * kFc2 + n];
                else if (k == kFc1) x = bias;
                else if (k == kFc1 + 1) x = bias - bf16_hi(bias);
            }
            v[j] = x;
        }
        o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    reinterpret_cast<uint4*>(blobs)[gid] = o;
}

// ---- front end: fce + GRUCell, fp32, thread per env, the agent's weights in shared memory ----
constexpr int kFrontThreads = 256, kMaxObs = 16;

struct FrontArgs {
    const float *we, *be, *w_ih, *b_ih, *w_hh, *b_hh;   // [A][in][32] | [A][32] | [A][32][96] | [A][96] | ...
    const float* obs;        // [E][A][n_obs]
    const float* hidden_in;  // [E][A][32]
    float* hidden_out;       // [E][A][32]
    int E, A, n_obs;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kFrontThreads, 2) flock_rnn_front_kernel(const __grid_constant__ FrontArgs a) {
    extern __shared__ __align__(16) float sw[];
    const int agent = blockIdx.x, n_obs = a.n_obs;
    float* sWe = sw;
    float* sbe = sWe + n_obs * kHid;
    float* sWih = sbe + kHid;
    float* sbih = sWih + kHid * 3 * kHid;
    float* sWhh = sbih + 3 * kHid;
    float* sbhh = sWhh + kHid * 3 * kHid;
    auto copy4 = [&](float* dst, const float* src, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int t = threadIdx.x; t < (n >> 2); t += kFrontThreads) d4[t] = s4[t];
    };
    copy4(sWe, a.we + (size_t)agent * n_obs * kHid, n_obs * kHid);
    copy4(sbe, a.be + (size_t)agent * kHid, kHid);
    copy4(sWih, a.w_ih + (size_t)agent * kHid * 3 * kHid, kHid * 3 * kHid);
    copy4(sbih, a.b_ih + (size_t)agent * 3 * kHid, 3 * kHid);
    copy4(sWhh, a.w_hh + (size_t)agent * kHid * 3 * kHid, kHid * 3 * kHid);
    copy4(sbhh, a.b_hh + (size_t)agent * 3 * kHid, 3 * kHid);
    __syncthreads();
    const int env = blockIdx.y * kFrontThreads + threadIdx.x;
    if (env >= a.E) return;
    const size_t ea = (size_t)env * a.A + agent;
    float x[kMaxObs];
#pragma unroll
    for (int i = 0; i < kMaxObs; ++i) x[i] = i < n_obs ? a.obs[ea * n_obs + i] : 0.0f;
    float f[kHid];                                        // fce output (no activation, net.py:57)
#pragma unroll
    for (int j = 0; j < kHid; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(sbe + j);
        f[j] = b.x; f[j + 1] = b.y; f[j + 2] = b.z; f[j + 3] = b.w;
    }
#pragma unroll
    for (int i = 0; i < kMaxObs; ++i) {
        if (i < n_obs) {
#pragma unroll
            for (int j = 0; j < kHid; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(sWe + i * kHid + j);
                f[j] = fmaf(x[i], w.x, f[j]); f[j + 1] = fmaf(x[i], w.y, f[j + 1]);
                f[j + 2] = fmaf(x[i], w.z, f[j + 2]); f[j + 3] = fmaf(x[i], w.w, f[j + 3]);
            }
        }
    }
    float hp[kHid];
    const float4* hin = reinterpret_cast<const float4*>(a.hidden_in + ea * kHid);
#pragma unroll
    for (int j = 0; j < kHid; j += 4) {
        const float4 v = hin[j >> 2];
        hp[j] = v.x; hp[j + 1] = v.y; hp[j + 2] = v.z; hp[j + 3] = v.w;
    }
    // torch.nn.GRUCell, gates r | z | n; four hidden units per (rolled) iteration, as in flock_qnet.cu
#pragma unroll 1
    for (int u = 0; u < kHid; u += 4) {
        float gi[3][4], gh[3][4];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float4 bi = *reinterpret_cast<const float4*>(sbih + g * kHid + u);
            const float4 bh = *reinterpret_cast<const float4*>(sbhh + g * kHid + u);
            gi[g][0] = bi.x; gi[g][1] = bi.y; gi[g][2] = bi.z; gi[g][3] = bi.w;
            gh[g][0] = bh.x; gh[g][1] = bh.y; gh[g][2] = bh.z; gh[g][3] = bh.w;
        }
#pragma unroll
        for (int i = 0; i < kHid; ++i) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                const float4 wi = *reinterpret_cast<const float4*>(sWih + i * 3 * kHid + g * kHid + u);
                const float4 wh = *reinterpret_cast<const float4*>(sWhh + i * 3 * kHid + g * kHid + u);
                gi[g][0] = fmaf(f[i], wi.x, gi[g][0]); gi[g][1] = fmaf(f[i], wi.y, gi[g][1]);
                gi[g][2] = fmaf(f[i], wi.z, gi[g][2]); gi[g][3] = fmaf(f[i], wi.w, gi[g][3]);
                gh[g][0] = fmaf(hp[i], wh.x, gh[g][0]); gh[g][1] = fmaf(hp[i], wh.y, gh[g][1]);
                gh[g][2] = fmaf(hp[i], wh.z, gh[g][2]); gh[g][3] = fmaf(hp[i], wh.w, gh[g][3]);
            }
        }
        const float4 hpu4 = hin[u >> 2];
        const float hpu[4] = {hpu4.x, hpu4.y, hpu4.z, hpu4.w};
        float hn[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float r = sigmoidf_(gi[0][c] + gh[0][c]);
            const float z = sigmoidf_(gi[1][c] + gh[1][c]);
            const float n = tanhf(gi[2][c] + r * gh[2][c]);
            hn[c] = (1.0f - z) * n + z * hpu[c];
        }
        reinterpret_cast<float4*>(a.hidden_out + ea * kHid)[u >> 2] = make_float4(hn[0], hn[1], hn[2], hn[3]);
    }
}

}  // namespace rnn

size_t rnn_actor_blob_bytes() { return (size_t)rnn::kBlobBytes; }
int rnn_actor_max_obs() { return rnn::kMaxObs; }

cudaError_t launch_rnn_actor_pack(int agents, const float* const* ptrs, void* blobs, cudaStream_t s) {
    rnn::PackArgs a{ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], ptrs[6], ptrs[7], agents};
    const size_t total = (size_t)(rnn::kBlobBytes / 16) * agents;
    rnn::flock_rnn_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, static_cast<uint8_t*>(blobs));
    return cudaGetLastError();
}

cudaError_t launch_rnn_actor_forward(const void* blobs, const float* const* front, const float* obs, const float* hidden_in,
                                     float* hidden_out, float* actions, int E, int N, int n_obs, float* ou_state,
                                     float ou_theta, float ou_mu, float ou_sigma, float ou_dt, uint64_t seed, uint32_t step,
                                     int env_offset, NoiseCounters ctr, cudaStream_t s) {
    rnn::OuArgs ou;
    ou.env_step = ctr.env_step;
    ou.env_epoch = ctr.env_epoch;
    ou.state = reinterpret_cast<float2*>(ou_state);
    ou.theta_dt = ou_theta * ou_dt;
    ou.mu = ou_mu;
    ou.sigma_sqrt_dt = ou_sigma * sqrtf(ou_dt);
    ou.seed_lo = (uint32_t)seed;
    ou.seed_hi = (uint32_t)(seed >> 32);
    ou.step = step;
    ou.env_offset = env_offset;
    static DeviceOnce once;
    int sm_count = 148;
    const cudaError_t configured = once.get(
        [] { return cudaFuncSetAttribute(rnn::flock_rnn_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rnn::kSmemBytes); },
        &sm_count);
    if (configured != cudaSuccess) return configured;
    // launch 1 of 2: fce + GRUCell (fp32); the hidden state may be updated in place (each thread reads its own
    // row before it writes it)
    cudaError_t err = cudaSuccess;
    if (front != nullptr) {   // nullptr: the caller has already run the front end (flock_gru_tc.cu) into hidden_out
        rnn::FrontArgs fa{front[0], front[1], front[2], front[3], front[4], front[5], obs, hidden_in, hidden_out, E, N, n_obs};
        const size_t fbytes = ((size_t)n_obs * rnn::kHid + rnn::kHid + 2 * (rnn::kHid * 3 * rnn::kHid + 3 * rnn::kHid)) * sizeof(float);
        const dim3 fgrid((unsigned)N, (unsigned)((E + rnn::kFrontThreads - 1) / rnn::kFrontThreads));
        rnn::flock_rnn_front_kernel<<<fgrid, rnn::kFrontThreads, fbytes, s>>>(fa);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    // launch 2 of 2: the MLP on the tensor cores, input = the new hidden state
    const int tiles = (E + rnn::kRows - 1) / rnn::kRows;
    const int total = tiles * N;
    const int per_cta = (total + sm_count - 1) / sm_count;
    const int grid = (total + per_cta - 1) / per_cta;
    return launch_pdl(rnn::flock_rnn_mlp_kernel, dim3(grid), dim3(rnn::kThreads), rnn::kSmemBytes, s, static_cast<const uint8_t*>(blobs),
                      static_cast<const float*>(hidden_out), actions, E, N, tiles, per_cta, ou);
}

}  // namespace flock
