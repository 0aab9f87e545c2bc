// flock_actor.cu -- fused per-agent actor MLP for the batched rollout (BASELINE configs[2]
// "MADDPG actor rollout"): the N per-agent ActorNetworks of
// learners/maddpg_shared_critic/ddpg_network.py:85-141
//     fc1(in -> 400) -> LayerNorm -> ReLU -> fc2(400 -> 300) -> LayerNorm -> ReLU -> mu(300 -> 2) -> tanh
// evaluated for all envs in ONE launch, straight from the env's observation buffer into the action
// buffer flock_step consumes. This IS GEMM-shaped work (250 kFLOP per agent-step, 40x the env step),
// so it runs on the 5th-generation tensor cores:
//   * persistent kernel, one CTA per SM; work item = 128 envs x one agent, dealt out in contiguous runs so a
//     CTA mostly stays on one agent; both matrix layers are tcgen05.mma (kind::f16, bf16 operands, fp32
//     accumulators in TMEM, M = 128), issued by one thread; operands sit in shared memory in the
//     no-swizzle K-major canonical layout (8-row x 16-byte core matrices);
//   * the per-agent weights are pre-packed once (flock_actor_pack) into exactly the shared-memory
//     image the MMA wants, so they stream from L2 with plain 1-D TMA bulk copies (cp.async.bulk +
//     mbarrier complete_tx) through a 4-slot ring of two-K-step chunks -- no tensor maps;
//   * one warp issues the TMA copies and one the MMAs, both warp-uniformly with a single elected lane
//     (descriptors stay in uniform registers); biases ride in the MMAs as two extra K slots (hi + lo
//     bf16 parts against constant-1 inputs);
//   * sixteen epilogue warps read the accumulators back with tcgen05.ld (warp w owns TMEM lanes
//     32(w%4)..+31 = 32 env rows and the column group w/4), apply bias + LayerNorm + ReLU in fp32
//     (row statistics combined across the four column groups through shared memory) and write the
//     layer-2 A operand back to shared memory as bf16; the 300 -> 2 head and tanh run in the second
//     epilogue on CUDA cores, so the activations never touch global memory.
// Accumulation and LayerNorm are fp32; operands are bf16 (tested against the fp32 PyTorch module).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "flock_device.cuh"
#include "flock_launch.h"
#include <type_traits>

#include "flock_tc.cuh"

namespace flock {
namespace actor {

constexpr int kRows = 128;                   // env rows per CTA = UMMA M
constexpr int kInPad = 16;                   // layer-1 K, one UMMA K step of bf16
constexpr int kMaxIn = kInPad - 2;           // two K slots carry the bias (hi + lo bf16 parts) against constant-1 inputs
constexpr int kFc1 = 400, kFc2 = 300, kFc2Pad = 304, kAct = 2;
constexpr int kK2 = kFc1 + 16;               // layer-2 K: 400 activations + one K step whose first two slots are the bias
constexpr int kSteps2 = kK2 / 16;            // 26 K steps of layer 2
constexpr int kW1Bytes = 2 * kFc1 * 16;      // 2 k-groups x 400 rows x 16 B
constexpr int kStepBytes = 2 * kFc2Pad * 16; // one K step of W2: 2 k-groups x 304 rows x 16 B
constexpr int kStepsPerChunk = 2;            // K steps per ring slot: one barrier round trip + commit per 4 MMAs
constexpr int kChunks = kSteps2 / kStepsPerChunk;
constexpr int kChunkBytes = kStepsPerChunk * kStepBytes;
constexpr int kStages = 4;                   // W2 ring depth (chunks)
static_assert(kSteps2 % kStepsPerChunk == 0, "whole chunks");
// fp32 parameters: g1 be1 [400] | g2 be2 [304] | w3[:,0] w3[:,1] [304] | b3[2] + 2 pad | Gram matrix of W1 [16][16] | row sums [16]
constexpr int kGramOff = 2 * kFc1 + 4 * kFc2Pad + 4;
constexpr int kSumOff = kGramOff + kInPad * kInPad;          // + row sums of the packed W1 over the outputs [16]
constexpr int kParamFloats = kSumOff + kInPad;
constexpr int kParamBytes = kParamFloats * 4;
constexpr int kBlobBytes = kW1Bytes + kParamBytes + kSteps2 * kStepBytes;
static_assert(kParamBytes % 16 == 0 && kBlobBytes % 16 == 0, "bulk copies move 16-byte units");

// shared-memory carve-up (offsets from a 128-byte aligned base)
constexpr int kOffA2 = 0;                                  // layer-2 A: 52 k-groups x 128 rows x 16 B
constexpr int kA2Bytes = (kK2 / 8) * kRows * 16;
constexpr int kOffA1 = kOffA2 + kA2Bytes;                  // layer-1 A: 2 k-groups x 128 rows x 16 B
constexpr int kA1Bytes = 2 * kRows * 16;
constexpr int kOffW1 = kOffA1 + kA1Bytes;
constexpr int kOffPar = kOffW1 + kW1Bytes;
constexpr int kOffRing = (kOffPar + kParamBytes + 127) & ~127;
constexpr int kOffBar = kOffRing + kStages * kChunkBytes;
constexpr int kNumBars = 2 * kStages + 7;
constexpr int kOffRed = kOffBar + kNumBars * 8 + 16;       // + tmem pointer; then row-statistic / head partials
constexpr int kColGroups = 4;
constexpr int kRedBytes = 2 * kColGroups * kRows * 8 + kRows * 8;   // two float2 [4][128] buffers + (rstd, -mean rstd) of layer 1 [128]
constexpr int kSmemBytes = kOffRed + kRedBytes + 128;       // + alignment slack
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr int kEpiWarps = 4 * kColGroups;     // warps 0-15: epilogues (TMEM lane quarter = w % 4, column group = w / 4)
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kMmaWarp = kEpiWarps;           // warp 16: TMEM allocation + MMA issue (one thread)
constexpr int kTmaWarp = kEpiWarps + 1;       // warp 17: TMA producer (one thread)
constexpr int kThreads = kEpiThreads + 64;
constexpr int kTmemCols = 512;
// Layer-1 accumulators live in TMEM columns [0, 400), layer-2 accumulators in [208, 512). Epilogue 1 first consumes
// the layer-1 columns [192, 400) (phase A), which frees everything the layer-2 accumulators overlap, so the layer-2
// MMAs over K steps 12..25 run while phase B turns columns [0, 192) into K steps 0..11.
constexpr int kL2Col = 208;
constexpr int kPhaseAUnit = 12;                       // first layer-1 column unit (16 columns) of phase A
constexpr int kFirstChunk = kPhaseAUnit / kStepsPerChunk;   // W2 chunks are streamed / multiplied in the order 6..12, 0..5
static_assert(kPhaseAUnit % kStepsPerChunk == 0 && kPhaseAUnit * 16 <= kL2Col && kL2Col + kFc2Pad <= kTmemCols, "phase split");

using namespace tc;

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// column units (16 accumulator columns each) owned by column group g: layer 1 has 25 units, layer 2 has 19
__device__ __forceinline__ int unit_begin(int units, int g) { return (units * g + kColGroups - 1) / kColGroups; }

// Optional exploration noise fused into the output stage: the Ornstein-Uhlenbeck process of the shared-critic
// learner (OUActionNoiseGPU, learners/maddpg_shared_critic/utils.py:6-21; mu' = mu + noise(),
// agent_simple_shared_critic.py:104), one independent process per (env, agent, action):
//     x <- x + theta (mu - x) dt + sigma sqrt(dt) N(0, 1),   action = tanh(...) + x
// with the normals from Philox4x32-10, counter (env_offset + env, agent, step, tag 6), key = seed.
struct OuArgs {
    float2* state;            // [E][A] (x0, x1), updated in place; nullptr = no noise
    float theta_dt, mu, sigma_sqrt_dt;
    uint32_t seed_lo, seed_hi, step;
    int env_offset;
    const int32_t* env_step;      // [E] device step counters added to `step` (nullable; replay-safe under CUDA graphs)
    const uint32_t* env_epoch;    // [E] device epoch counters folded into the tag word (nullable)
    // uw ring layout of the env's observation history (flock_buffers_t.obs_head): obs is [E][ring_h][N][ring_k] and
    // feature r * ring_k + c of an agent is row r of its newest-first window = slot (obs_head[env] + r) % ring_h
    const int32_t* obs_head;      // nullptr = obs is the plain [E][N][in_dims] matrix
    int ring_h, ring_k;
};
constexpr uint32_t kTagOu = 6u;

// Persistent kernel: grid = min(#SMs, work items) CTAs, one per SM; work item = (agent, tile of 128 envs),
// items are dealt out in contiguous runs so that a CTA mostly stays on one agent (W1 and the fp32 parameters
// are reloaded only when the agent changes; TMEM, the barriers and the W2 ring live across items).
// obs [E][N][in_dims] fp32, out [E][N][2] fp32
__global__ void __launch_bounds__(kThreads, 1)
flock_actor_kernel(const uint8_t* __restrict__ blobs, const float* __restrict__ obs, float* __restrict__ out, int E, int N,
                   int in_dims, int tiles, int items_per_cta, OuArgs ou, long long* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    // phase timestamps of the CTA's FIRST item (FLOCK_ACTOR_TIMING=1, see launch_actor_forward): 16 clock64 slots
    long long* dbg_cta = dbg != nullptr ? dbg + (size_t)blockIdx.x * 16 : nullptr;
    auto stamp = [&](int slot) {
        if (dbg_cta != nullptr) dbg_cta[slot] = clock64();
    };
    if (threadIdx.x == 0) stamp(0);
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
    float2* red_stat = reinterpret_cast<float2*>(sm + kOffRed);               // [4][128] (sum, sum of squares)
    float2* red_head = red_stat + kColGroups * kRows;                          // [4][128] (o0, o1)
    float2* row_stat = red_head + kColGroups * kRows;                          // [128] layer-1 (1/sigma, -mean/sigma), from the Gram matrix

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
    if (threadIdx.x == 0) stamp(1);

    // The two single-thread roles run warp-uniformly (all 32 lanes execute the control flow and the
    // address arithmetic, so it stays in uniform registers); only the issuing instructions are
    // predicated on one elected lane. Under `if (lane == 0)` every descriptor went through a
    // per-thread -> uniform register waterfall and the issue loop, not the tensor core, set the pace
    // (316 cycles per K step instead of 152).
    // Per-item barriers (a1, mma1, a2, mma2, done) complete once per item: parity = local item index & 1.
    // bar_w1 completes once per agent change: every role counts the changes the same way.
    if (warp == kTmaWarp) {
        // ---- TMA producer: W1 + parameters on an agent change, the W2 chunks through the ring ----
        const bool leader = elect_one();
        int prev_agent = -1;
        uint32_t g = 0;                                  // chunks issued so far (ring position)
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / tiles;
            const uint8_t* blob = blobs + (size_t)agent * kBlobBytes;
            if (agent != prev_agent) {
                // the previous item's epilogues still read the old parameters: wait until they are done
                if (it > 0) mbar_wait(bar_done, (uint32_t)(it - 1) & 1u);
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
                const int c = (idx + kFirstChunk) % kChunks;     // the order the MMA warp consumes them in
                const uint32_t slot = g % kStages;
                if (g >= kStages) mbar_wait(bar_empty(slot), (g / kStages - 1u) & 1u);   // MMAs of the chunk that used the slot are done
                if (leader) {
                    mbar_expect_tx(bar_full(slot), kChunkBytes);
                    bulk_g2s(sRing + slot * kChunkBytes, w2 + (size_t)c * kChunkBytes, kChunkBytes, bar_full(slot));
                }
                __syncwarp();
            }
        }
    } else if (warp == kMmaWarp) {
        // ---- MMA issuer ----
        const bool leader = elect_one();
        int prev_agent = -1;
        uint32_t w1_loads = 0, g = 0;
        const uint64_t da1 = umma_desc(sA1, kRows * 16, 128);
        const uint64_t da0 = umma_desc(sA2, kRows * 16, 128);
        const uint64_t db0 = umma_desc(sRing, kFc2Pad * 16, 128);
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / tiles;
            const uint32_t ph = (uint32_t)it & 1u;
            // layer 1: [128 x 16] x [16 x 400] -> TMEM columns [0, 400). The previous item's epilogue 2 still reads
            // columns [0, 304): wait until every epilogue thread has left it.
            if (it > 0) mbar_wait(bar_done, (uint32_t)(it - 1) & 1u);
            mbar_wait(bar_a1, ph);
            if (lane == 0 && it == 0) stamp(8);
            if (agent != prev_agent) {
                mbar_wait(bar_w1, w1_loads & 1u);
                ++w1_loads;
                prev_agent = agent;
            }
            if (lane == 0 && it == 0) stamp(9);
            tc_fence_after();
            if (leader) {
                umma_bf16(tmem + 0, da1, umma_desc(sW1, kFc1 * 16, 128), umma_idesc(kRows, 256), 0u);
                umma_bf16(tmem + 256, da1, umma_desc(sW1 + 256 * 16, kFc1 * 16, 128), umma_idesc(kRows, 144), 0u);
                umma_commit(bar_mma1);
            }
            __syncwarp();
            // layer 2: [128 x 416] x [416 x 304] -> TMEM columns [208, 512). K steps 12..25 first: their A operand
            // (phase A of epilogue 1) is ready early and the accumulators only overlap layer-1 columns that phase A
            // has consumed; K steps 0..11 follow when phase B is done.
            mbar_wait(bar_a2, ph);
            if (lane == 0 && it == 0) stamp(10);
            tc_fence_after();
#pragma unroll
            for (int idx = 0; idx < kChunks; ++idx, ++g) {
                constexpr int kAStep = (2 * kRows * 16) >> 4, kBSlot = kChunkBytes >> 4,
                              kBStep = kStepBytes >> 4;   // 16-byte units of the descriptor start-address field
                const int c = (idx + kFirstChunk) % kChunks;
                if (c == 0) {                              // first chunk of phase B
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
                        const uint64_t da = da0 + (uint64_t)(s * kAStep);          // no carry out of the 14-bit field
                        const uint64_t db = db0 + (uint64_t)(st * kBSlot + j * kBStep);
                        const uint32_t acc = (idx > 0 || j > 0) ? 1u : 0u;
                        // N = 256 + 48 (measured faster than 160 + 144: 171 vs 194 cycles per K step)
                        umma_bf16(tmem + kL2Col, da, db, umma_idesc(kRows, 256), acc);
                        umma_bf16(tmem + kL2Col + 256, da, db + (256 * 16 >> 4), umma_idesc(kRows, kFc2Pad - 256), acc);
                    }
                    umma_commit(bar_empty(st));
                }
                __syncwarp();
            }
            if (leader) umma_commit(bar_mma2);
            __syncwarp();
            if (lane == 0 && it == 0) stamp(11);
        }
    } else {
        // ---- epilogue warps: thread = (env row, column group) ----
        const int q = warp & 3, cg = warp >> 2;
        const int row = q * 32 + lane;                     // TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        const float* g1 = par;
        const float* be1 = par + kFc1;
        const float* g2 = par + 2 * kFc1;
        const float* be2 = g2 + kFc2Pad;
        const float* w3a = g2 + 2 * kFc2Pad;
        const float* w3b = g2 + 3 * kFc2Pad;
        const float* b3 = g2 + 4 * kFc2Pad;
        // layer-1 columns of this column group: phase A = units [12, 25), phase B = units [0, 12)
        const int cab = (kPhaseAUnit + unit_begin(kFc1 / 16 - kPhaseAUnit, cg)) * 16,
                  cae = (kPhaseAUnit + unit_begin(kFc1 / 16 - kPhaseAUnit, cg + 1)) * 16;
        const int cbb = unit_begin(kPhaseAUnit, cg) * 16, cbe = unit_begin(kPhaseAUnit, cg + 1) * 16;
        const int c2b = unit_begin(kFc2Pad / 16, cg) * 16, c2e = unit_begin(kFc2Pad / 16, cg + 1) * 16;
        const uint32_t trow2 = trow + kL2Col;                  // layer-2 accumulators
        int prev_agent = -1;
        uint32_t w1_loads = 0;
        // observation ring: the head of the NEXT item's env is loaded one item ahead, so the rotated row loads of an
        // item issue at once instead of behind a second dependent global round trip (the dependent head load cost 9 us per policy step at 4096 x 32)
        int ring_head = 0;
        if (cg == 0 && ou.obs_head != nullptr && item0 < item1) {
            const int e0 = (item0 - (item0 / tiles) * tiles) * kRows + row;
            if (e0 < E) ring_head = ou.obs_head[e0];
        }
        if (cg == 1) {
            // the bias K step of layer 2 never changes: A2[:, 400] = A2[:, 401] = 1, A2[:, 402..415] = 0
            sts128(sA2 + (kFc1 / 8) * (kRows * 16) + row * 16, pack_bf16(1.0f, 1.0f), 0u, 0u, 0u);
            sts128(sA2 + (kFc1 / 8 + 1) * (kRows * 16) + row * 16, 0u, 0u, 0u, 0u);
        }
        for (int item = item0, it = 0; item < item1; ++item, ++it) {
            const int agent = item / tiles, tile = item - agent * tiles;
            const uint32_t ph = (uint32_t)it & 1u;
            const int env = tile * kRows + row;
            const bool valid = env < E;
            const bool first = it == 0 && threadIdx.x == 0;
            int ring_head_next = 0;
            if (cg == 0 && item + 1 < item1) {   // pull the next item's observation row towards L2 while this item computes
                const int nagent = (item + 1) / tiles, ntile = (item + 1) - nagent * tiles;
                const int nenv = ntile * kRows + row;
                if (nenv < E && ou.obs_head == nullptr)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(obs + ((size_t)nenv * N + nagent) * in_dims) : "memory");
                if (ou.obs_head != nullptr) ring_head_next = nenv < E ? ou.obs_head[nenv] : 0;
            }
            if (cg == 0) {
                // this row's observation = the layer-1 A operand; the two slots after the inputs are the
                // constant 1 that multiplies the bias rows of the packed W1. (MMA1 of the previous item is long
                // complete: its accumulators were consumed before this thread got here.)
                float xin[kInPad];
#pragma unroll
                for (int i = 0; i < kInPad; ++i) xin[i] = 0.0f;
                if (valid && ou.obs_head != nullptr) {   // the env's observation ring, read in place (rotated)
                    const int head = ring_head;
                    auto load_rows = [&](auto rk_tag) {      // ring_k as a compile-time constant: xin stays in registers
                        constexpr int RK = decltype(rk_tag)::value;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < ou.ring_h && (r + 1) * RK <= kMaxIn) {
                                const int slot = (head + r) % ou.ring_h;
                                const float* src = obs + (((size_t)env * ou.ring_h + slot) * N + agent) * RK;
#pragma unroll
                                for (int c = 0; c < RK; ++c) xin[r * RK + c] = src[c];
                            }
                        }
                    };
                    if (ou.ring_k == 3) load_rows(std::integral_constant<int, 3>{});
                    else if (ou.ring_k == 2) load_rows(std::integral_constant<int, 2>{});
                    else load_rows(std::integral_constant<int, 1>{});
                } else if (valid) {
                    const float* src = obs + ((size_t)env * N + agent) * in_dims;
                    if (in_dims == 12) {
                        const float4* s4 = reinterpret_cast<const float4*>(src);
                        const float4 a = s4[0], b = s4[1], c = s4[2];
                        xin[0] = a.x; xin[1] = a.y; xin[2] = a.z; xin[3] = a.w;
                        xin[4] = b.x; xin[5] = b.y; xin[6] = b.z; xin[7] = b.w;
                        xin[8] = c.x; xin[9] = c.y; xin[10] = c.z; xin[11] = c.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < kMaxIn; ++i)
                            if (i < in_dims) xin[i] = src[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < kInPad; ++i)
                    if (i == in_dims || i == in_dims + 1) xin[i] = 1.0f;
                sts128(sA1 + row * 16, pack_bf16(xin[0], xin[1]), pack_bf16(xin[2], xin[3]), pack_bf16(xin[4], xin[5]),
                       pack_bf16(xin[6], xin[7]));
                sts128(sA1 + kRows * 16 + row * 16, pack_bf16(xin[8], xin[9]), pack_bf16(xin[10], xin[11]),
                       pack_bf16(xin[12], xin[13]), pack_bf16(xin[14], xin[15]));
                fence_proxy_async();        // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(bar_a1);
                ring_head = ring_head_next;
                if (first) stamp(2);
                if (agent != prev_agent) mbar_wait(bar_w1, w1_loads & 1u);   // this agent's fp32 parameters have landed
                // Layer-1 LayerNorm statistics without touching the accumulators: the packed weights are centred
                // (row mean == 0) and sum_n h_n^2 = x^T G x with G the Gram matrix of the packed weights and x the
                // bf16-rounded inputs the MMA sees. Runs while the MMA is in flight.
                const float* G = par + kGramOff;
                float xr[kInPad];
#pragma unroll
                for (int i = 0; i < kInPad; ++i) xr[i] = __bfloat162float(__float2bfloat16_rn(xin[i]));
                float ss = 0.0f;
#pragma unroll
                for (int i = 0; i < kInPad; ++i) {
                    float acc = 0.0f;
#pragma unroll
                    for (int jb = (i >> 2) << 2; jb < kInPad; jb += 4) {     // upper triangle, off-diagonals pre-doubled
                        const float4 gq = *reinterpret_cast<const float4*>(G + i * kInPad + jb);
                        if (jb + 0 >= i) acc = fmaf(gq.x, xr[jb + 0], acc);
                        if (jb + 1 >= i) acc = fmaf(gq.y, xr[jb + 1], acc);
                        if (jb + 2 >= i) acc = fmaf(gq.z, xr[jb + 2], acc);
                        if (jb + 3 >= i) acc = fmaf(gq.w, xr[jb + 3], acc);
                    }
                    ss = fmaf(xr[i], acc, ss);
                }
                float rsum = 0.0f;
#pragma unroll
                for (int i = 0; i < kInPad; ++i) rsum = fmaf(xr[i], par[kSumOff + i], rsum);
                const float mean1 = rsum * (1.0f / kFc1);      // residue of the centring after bf16 rounding (tiny)
                const float rstd1 = rsqrtf(fmaxf(ss * (1.0f / kFc1) - mean1 * mean1, 0.0f) + 1.0e-5f);
                row_stat[row] = make_float2(rstd1, -mean1 * rstd1);
            }
            if (agent != prev_agent) {
                if (cg != 0) mbar_wait(bar_w1, w1_loads & 1u);       // this agent's fp32 parameters have landed
                ++w1_loads;
                prev_agent = agent;
            }

            // ---- epilogue 1: LayerNorm(400) + ReLU -> bf16 A operand of layer 2 (the bias came with the MMA) ----
            mbar_wait(bar_mma1, ph);
            tc_fence_after();
            if (first) stamp(3);
            epi_sync();                     // row_rstd of every row is written (and the previous item's reads are over)
            if (first) stamp(4);
            float rstd = row_stat[row].x, nmr = row_stat[row].y;
            auto norm_unit = [&](int c0, const uint32_t (&r)[16]) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] = fmaxf(fmaf(fmaf(__uint_as_float(r[i]), rstd, nmr), g1[c0 + i], be1[c0 + i]), 0.0f);
                const uint32_t dst = sA2 + (c0 >> 3) * (kRows * 16) + row * 16;
                sts128(dst, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                sts128(dst + kRows * 16, pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                       pack_bf16(v[14], v[15]));
            };
            for_each_unit(trow, cab, cae, norm_unit);       // phase A: columns [192, 400) -> K steps 12..24
            fence_proxy_async();
            tc_fence_before();              // our tcgen05.ld of columns [192,400) precede the layer-2 MMAs that overwrite them
            mbar_arrive(bar_a2);
            for_each_unit(trow, cbb, cbe, norm_unit);       // phase B: columns [0, 192) -> K steps 0..11, under the first MMAs
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_a2b);
            if (first) stamp(5);

            // ---- epilogue 2: LayerNorm(300) + ReLU, mu head (300 -> 2), tanh ----
            // (the four padded columns 300..303 are exactly 0: zero weights, zero bias)
            mbar_wait(bar_mma2, ph);
            tc_fence_after();
            if (first) stamp(6);
            // centred W2 and bias: the row mean is only the rounding residue, and the MMA delivers its 300-fold in the
            // padding column 300 (see flock_actor_pack_kernel) -- one pass for the sum of squares is all that is left
            float sq = 0.0f, s300 = 0.0f;
            for_each_unit(trow2, c2b, c2e, [&](int, const uint32_t (&r)[16]) {
#pragma unroll
                for (int i = 0; i < 16; ++i) sq = fmaf(__uint_as_float(r[i]), __uint_as_float(r[i]), sq);
            });
            if (cg == kColGroups - 1) {     // this group's last unit holds column 300: take it out of the squares
                uint32_t r[16];
                tmem_ld16_issue(trow2 + (kFc2Pad - 16), r);
                tmem_ld16_wait(r);
                s300 = __uint_as_float(r[kFc2 - (kFc2Pad - 16)]);
                sq = fmaf(-s300, s300, sq);
            }
            red_stat[cg * kRows + row] = make_float2(sq, s300);   // (all epilogue-1 reads happened before bar_a2 completed)
            epi_sync();
            sq = 0.0f;
#pragma unroll
            for (int g = 0; g < kColGroups; ++g) sq += red_stat[g * kRows + row].x;
            const float mean2 = red_stat[(kColGroups - 1) * kRows + row].y * (1.0f / kFc2);
            rstd = rsqrtf(fmaxf(sq * (1.0f / kFc2) - mean2 * mean2, 0.0f) + 1.0e-5f);
            nmr = -mean2 * rstd;
            float o0 = 0.0f, o1 = 0.0f;
            for_each_unit(trow2, c2b, c2e, [&](int c0, const uint32_t (&r)[16]) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    // padded columns: g2 = be2 = w3 = 0
                    const float yv = fmaxf(fmaf(fmaf(__uint_as_float(r[i]), rstd, nmr), g2[c0 + i], be2[c0 + i]), 0.0f);
                    o0 = fmaf(yv, w3a[c0 + i], o0);
                    o1 = fmaf(yv, w3b[c0 + i], o1);
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
                a.x = tanhf(o0 + b3[0]);
                a.y = tanhf(o1 + b3[1]);
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
            // this item's TMEM reads and parameter reads are over: the next layer-1 MMAs / parameter loads may go
            tc_fence_before();
            mbar_arrive(bar_done);
            if (first) stamp(7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
        if (lane == 0) stamp(12);
    }
}

// Pre-pack the per-agent parameters (fp32, the layout of policies.BatchedActors: w [A][in][out],
// vectors [A][out]) into the shared-memory images of the kernel above. One thread per 16 bytes.
struct PackArgs {
    const float *w1, *b1, *g1, *be1, *w2, *b2, *g2, *be2, *w3, *b3;
    int agents, in_dims;
    const float* means;      // per agent: column means of W1 [16] | mean(b1) | column means of W2 [400] | mean(b2), see prep
};
constexpr int kMeanFloats = kInPad + 1 + kFc1 + 1;

__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// LayerNorm subtracts the row mean of h = x W + b; that mean is x (W 1/n) + mean(b), linear in the weights, so it
// is removed ONCE at pack time by centring every weight row (and the bias) over the output features: the MMAs
// then deliver mean-free pre-activations and the epilogues only need the sum of squares.
// Pack step 1: the means. One CTA per agent, thread t owns input row t.
__global__ void __launch_bounds__(512) flock_actor_prep_kernel(PackArgs a, float* __restrict__ means) {
    const int ag = blockIdx.x, t = threadIdx.x;
    float* m = means + (size_t)ag * kMeanFloats;
    if (t < kInPad) {
        float acc = 0.0f;
        if (t < a.in_dims)
            for (int n = 0; n < kFc1; ++n) acc += a.w1[((size_t)ag * a.in_dims + t) * kFc1 + n];
        m[t] = acc * (1.0f / kFc1);
    } else if (t == kInPad) {
        float acc = 0.0f;
        for (int n = 0; n < kFc1; ++n) acc += a.b1[(size_t)ag * kFc1 + n];
        m[kInPad] = acc * (1.0f / kFc1);
    } else if (t >= 32 && t < 32 + kFc1) {
        const int k = t - 32;
        float acc = 0.0f;
        for (int n = 0; n < kFc2; ++n) acc += a.w2[((size_t)ag * kFc1 + k) * kFc2 + n];
        m[kInPad + 1 + k] = acc * (1.0f / kFc2);
    } else if (t == 32 + kFc1) {
        float acc = 0.0f;
        for (int n = 0; n < kFc2; ++n) acc += a.b2[(size_t)ag * kFc2 + n];
        m[kInPad + 1 + kFc1] = acc * (1.0f / kFc2);
    }
}

// Pack step 3 (after the images exist): Gram matrix of the packed, bf16-rounded, centred layer-1 weights over the
// 400 outputs, G[i][j] = sum_n W[i][n] W[j][n]. With it the layer-1 row variance is the quadratic form x^T G x / 400
// of the 16 (bf16-rounded) inputs -- known before the accumulators are read, so epilogue 1 needs no statistics
// pass. Stored upper-triangular with doubled off-diagonal entries. One CTA per agent, thread = (i, j).
__global__ void __launch_bounds__(kInPad * kInPad) flock_actor_gram_kernel(uint8_t* __restrict__ blobs) {
    const int ag = blockIdx.x, i = threadIdx.x / kInPad, j = threadIdx.x % kInPad;
    uint8_t* blob = blobs + (size_t)ag * kBlobBytes;
    const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(blob);       // W1 image: [k-group][n][8]
    float acc = 0.0f;
    if (j >= i)
        for (int n = 0; n < kFc1; ++n)
            acc = fmaf(__bfloat162float(w[((i >> 3) * kFc1 + n) * 8 + (i & 7)]),
                       __bfloat162float(w[((j >> 3) * kFc1 + n) * 8 + (j & 7)]), acc);
    float* par = reinterpret_cast<float*>(blob + kW1Bytes);
    par[kGramOff + i * kInPad + j] = j > i ? 2.0f * acc : (j == i ? acc : 0.0f);
    if (j == 0) {   // what is left of the row mean after rounding the centred weights to bf16: x . rowsum / 400
        float rs = 0.0f;
        for (int n = 0; n < kFc1; ++n) rs += __bfloat162float(w[((i >> 3) * kFc1 + n) * 8 + (i & 7)]);
        par[kSumOff + i] = rs;
    }
}

__global__ void flock_actor_pack_kernel(PackArgs a, uint8_t* __restrict__ blobs) {
    const int units = kBlobBytes / 16;
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (gid >= (size_t)units * a.agents) return;
    const int ag = (int)(gid / units);
    int u = (int)(gid % units);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (u < kW1Bytes / 16) {                       // W1 image: [k-group][n] x 8 bf16; K slots in_dims, in_dims+1 = bias hi, lo
        const int kg = u / kFc1, n = u % kFc1;
        float v[8];
        const float* mm = a.means + (size_t)ag * kMeanFloats;
        for (int j = 0; j < 8; ++j) {
            const int k = kg * 8 + j;
            const float bias = a.b1[(size_t)ag * kFc1 + n] - mm[kInPad];
            v[j] = k < a.in_dims ? a.w1[((size_t)ag * a.in_dims + k) * kFc1 + n] - mm[k]
                 : (k == a.in_dims ? bias : (k == a.in_dims + 1 ? bias - bf16_hi(bias) : 0.0f));
        }
        o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    } else if (u < (kW1Bytes + kParamBytes) / 16) {   // fp32 parameters
        const int f0 = (u - kW1Bytes / 16) * 4;
        float v[4];
        for (int j = 0; j < 4; ++j) {
            int f = f0 + j;
            float x = 0.0f;
            if (f >= kGramOff) {
                x = 0.0f;                                   // Gram matrix, row sums: written by flock_actor_gram_kernel
            } else if (f < 2 * kFc1) {
                const float* src = f < kFc1 ? a.g1 : a.be1;
                x = src[(size_t)ag * kFc1 + f % kFc1];
            } else {
                f -= 2 * kFc1;
                const int which = f / kFc2Pad, c = f % kFc2Pad;
                if (which < 2) {
                    const float* src = which == 0 ? a.g2 : a.be2;
                    x = c < kFc2 ? src[(size_t)ag * kFc2 + c] : 0.0f;
                } else if (which < 4) {
                    x = c < kFc2 ? a.w3[((size_t)ag * kFc2 + c) * kAct + (which - 2)] : 0.0f;
                } else {
                    x = c < kAct ? a.b3[(size_t)ag * kAct + c] : 0.0f;
                }
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
            const float* mm = a.means + (size_t)ag * kMeanFloats + kInPad + 1;
            auto entry = [&](int kk, int nn) {     // centred layer-2 weight / bias-hi / bias-lo entry (kk, nn), nn < 300
                const float bias = a.b2[(size_t)ag * kFc2 + nn] - mm[kFc1];
                if (kk < kFc1) return a.w2[((size_t)ag * kFc1 + kk) * kFc2 + nn] - mm[kk];
                if (kk == kFc1) return bias;
                if (kk == kFc1 + 1) return bias - bf16_hi(bias);
                return 0.0f;
            };
            if (n < kFc2) {
                x = entry(k, n);
            } else if (n == kFc2 && k <= kFc1 + 1) {
                // output column 300 (padding) = row sum of the ROUNDED entries: the MMA then delivers the exact sum of
                // the 300 real pre-activations of each row, i.e. what is left of the row mean after rounding
                for (int nn = 0; nn < kFc2; ++nn) x += bf16_hi(entry(k, nn));
            }
            v[j] = x;
        }
        o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    reinterpret_cast<uint4*>(blobs)[gid] = o;
}

}  // namespace actor

size_t actor_blob_bytes() { return (size_t)actor::kBlobBytes; }
int actor_max_in_dims() { return actor::kMaxIn; }
void actor_dims(int* fc1, int* fc2, int* n_actions) {
    *fc1 = actor::kFc1;
    *fc2 = actor::kFc2;
    *n_actions = actor::kAct;
}

cudaError_t launch_actor_pack(int agents, int in_dims, const float* const* ptrs, void* blobs, cudaStream_t s) {
    float* means = nullptr;      // stream-ordered scratch for the centring means
    cudaError_t err = cudaMallocAsync(&means, (size_t)agents * actor::kMeanFloats * sizeof(float), s);
    if (err != cudaSuccess) return err;
    actor::PackArgs a{ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], ptrs[6], ptrs[7], ptrs[8], ptrs[9], agents, in_dims, means};
    actor::flock_actor_prep_kernel<<<agents, 512, 0, s>>>(a, means);
    const size_t total = (size_t)(actor::kBlobBytes / 16) * agents;
    actor::flock_actor_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, static_cast<uint8_t*>(blobs));
    actor::flock_actor_gram_kernel<<<agents, actor::kInPad * actor::kInPad, 0, s>>>(static_cast<uint8_t*>(blobs));
    err = cudaGetLastError();
    cudaFreeAsync(means, s);
    return err;
}

cudaError_t launch_actor_forward(const void* blobs, const float* obs, float* actions, int E, int N, int in_dims,
                                 float* ou_state, float ou_theta, float ou_mu, float ou_sigma, float ou_dt, uint64_t seed,
                                 uint32_t step, int env_offset, NoiseCounters ctr, const int32_t* obs_head, int ring_h, int ring_k,
                                 cudaStream_t s) {
    actor::OuArgs ou;
    ou.env_step = ctr.env_step;
    ou.env_epoch = ctr.env_epoch;
    ou.obs_head = obs_head;
    ou.ring_h = ring_h;
    ou.ring_k = ring_k;
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
        [] { return cudaFuncSetAttribute(actor::flock_actor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, actor::kSmemBytes); },
        &sm_count);
    if (configured != cudaSuccess) return configured;
    const int tiles = (E + actor::kRows - 1) / actor::kRows;
    const int total = tiles * N;
    const int per_cta = (total + sm_count - 1) / sm_count;          // contiguous run of items per CTA
    const int grid = (total + per_cta - 1) / per_cta;
    const uint8_t* bl = static_cast<const uint8_t*>(blobs);
    // FLOCK_ACTOR_TIMING=1 (developer knob): the first launch records clock64 at the phase boundaries of the first
    // item of every CTA, synchronises and prints the mean phase lengths to stderr.
    static bool timing = getenv("FLOCK_ACTOR_TIMING") != nullptr;
    if (timing) {
        timing = false;
        long long* dbg = nullptr;
        if (cudaMalloc(&dbg, (size_t)grid * 16 * sizeof(long long)) == cudaSuccess) {
            cudaMemsetAsync(dbg, 0, (size_t)grid * 16 * sizeof(long long), s);
            actor::flock_actor_kernel<<<grid, actor::kThreads, actor::kSmemBytes, s>>>(bl, obs, actions, E, N, in_dims, tiles,
                                                                                   per_cta, ou, dbg);
            cudaError_t e = cudaGetLastError();
            cudaStreamSynchronize(s);
            long long* h = static_cast<long long*>(malloc((size_t)grid * 16 * sizeof(long long)));
            cudaMemcpy(h, dbg, (size_t)grid * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
            static const char* names[12] = {"setup (alloc, barrier init, sync)", "obs -> A1", "wait MMA1", "epilogue-1 pass 1 + sync",
                                            "epilogue-1 pass 2 -> A2", "wait MMA2", "epilogue 2", "items 2.. + teardown",
                                            "[mma warp] start -> A1 ready", "[mma warp] W1 ready", "[mma warp] -> A2 ready",
                                            "[mma warp] layer-2 issue loop"};
            static const int from[12] = {0, 1, 2, 3, 4, 5, 6, 7, 1, 8, 9, 10}, to[12] = {1, 2, 3, 4, 5, 6, 7, 12, 8, 9, 10, 11};
            fprintf(stderr, "flock_actor_kernel: %d CTAs x %d items, phase means of the first item (SM clocks):\n", grid, per_cta);
            for (int ph = 0; ph < 12; ++ph) {
                double acc = 0;
                for (int c = 0; c < grid; ++c) acc += (double)(h[c * 16 + to[ph]] - h[c * 16 + from[ph]]);
                fprintf(stderr, "  %-38s %9.0f\n", names[ph], acc / grid);
            }
            double tot = 0;
            for (int c = 0; c < grid; ++c) tot += (double)(h[c * 16 + 12] - h[c * 16 + 0]);
            fprintf(stderr, "  %-38s %9.0f\n", "CTA total (all items)", tot / grid);
            free(h);
            cudaFree(dbg);
            return e;
        }
    }
    return launch_pdl(actor::flock_actor_kernel, dim3(grid), dim3(actor::kThreads), actor::kSmemBytes, s, bl, obs, actions, E, N, in_dims,
                      tiles, per_cta, ou, static_cast<long long*>(nullptr));
}

}  // namespace flock
