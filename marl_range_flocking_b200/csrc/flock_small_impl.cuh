// flock_small_impl.cuh -- templates of the warp-per-env-group path (N <= 32): device helpers, the fused
// step kernel, reset_groups() and the launch dispatch. Included by flock_small.cu (reset / misc
// kernels, top-level dispatch) and by the per-variant translation units flock_small_<variant>.cu,
// which explicitly instantiate launch_step_small_vpi<V, PER, IDX> so the ~290 kernel variants
// compile in parallel.
#pragma once
#include <cstdlib>

#include "flock_device.cuh"
#include "flock_launch.h"

#ifndef FLOCK_LATE_TRIGGER
#define FLOCK_LATE_TRIGGER 0        // 1: developer A/B build with the dependents' trigger before the result stores for every variant
#endif
#ifndef FLOCK_PDL_PREFETCH
#define FLOCK_PDL_PREFETCH 1      // 0: developer A/B build without the pre-wait L2 prefetch
#endif

namespace flock {

constexpr int kSmallThreads = 64;    // 64-thread CTAs measured best (8.27 vs 8.43 us on cfg3, 7.57 vs 7.89 on cfg4)
constexpr int kSmallWarps = kSmallThreads / kWarp;
constexpr int kSlots = 64;  // floats per staged array per warp: G * roundup(N,4) <= 64

struct LaneMap {
    int g, a;
    bool lane_ok;
    unsigned gmask;
};

__device__ __forceinline__ LaneMap lane_map(int lane, int N, int G, int g_magic) {
    LaneMap m;
    m.lane_ok = lane < G * N;
    m.g = m.lane_ok ? (lane * g_magic) >> 16 : 0;      // lane / N without the integer divide
    m.a = m.lane_ok ? lane - m.g * N : 0;
    const unsigned ones = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
    m.gmask = m.lane_ok ? (ones << (m.g * N)) : (1u << lane);
    return m;
}

// vectorised row store: n 4-byte values per agent, rows contiguous in memory (dst = base + row*n)
template <typename T, int MAXN>
__device__ __forceinline__ void store_row(T* dst, const T (&v)[MAXN], int n) {
    static_assert(sizeof(T) == 4, "4-byte elements");
    struct alignas(16) Vec4 { T a, b, c, d; };
    if constexpr (MAXN == 4) {   // K = 4 is only instantiated for k == 4: always one 16-byte store
        reinterpret_cast<Vec4*>(dst)[0] = Vec4{v[0], v[1], v[2], v[3]};
        return;
    }
    if constexpr (MAXN >= 8) {
        if (n == 8) {
            reinterpret_cast<Vec4*>(dst)[0] = Vec4{v[0], v[1], v[2], v[3]};
            reinterpret_cast<Vec4*>(dst)[1] = Vec4{v[4], v[5], v[6], v[7]};
            return;
        }
    }
    if constexpr (MAXN >= 4) {
        if (n == 4) {
            reinterpret_cast<Vec4*>(dst)[0] = Vec4{v[0], v[1], v[2], v[3]};
            return;
        }
    }
#pragma unroll
    for (int s = 0; s < MAXN; ++s)
        if (s < n) dst[s] = v[s];
}

// obs row(s) of one agent after a step (`fresh` = false) or a reset (`fresh` = true: the history is zero-filled,
// gym_flock_uw.py:100-102). H == 1: obs[idx][k]. H > 1 (uw): window layout = shift by one row and put the new
// ranges first (gym_flock_uw.py:120-123); ring layout = write the new row into slot `slot` (the env's NEW head
// after a step, its unchanged head after a reset) and, when fresh, zero the other slots.
template <int K>
__device__ __forceinline__ void write_obs(const Params& p, int env, int a, size_t idx, const float (&dist)[K], bool fresh,
                                          int slot) {
    const int k = p.k;
    if (p.H == 1) {
        store_row<float, K>(p.obs + idx * k, dist, k);
        return;
    }
    if (p.obs_head != nullptr) {
        for (int sl = 0; sl < p.H; ++sl) {
            if (sl != slot && !fresh) continue;
            float* o = ring_row(p, env, sl, a);
#pragma unroll
            for (int s = 0; s < K; ++s)
                if (s < k) o[s] = (sl == slot) ? dist[s] : 0.0f;
        }
        return;
    }
    float* o = p.obs + idx * (size_t)(p.H * k);
    const int keep = (p.H - 1) * k;
    if (k == 3 && p.H == 4) {  // 12 floats = 3 x float4, the reference's configuration
        float4* o4 = reinterpret_cast<float4*>(o);
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
        if (!fresh) {
            r0 = o4[0];
            r1 = o4[1];
            r2 = o4[2];
        }
        o4[0] = make_float4(dist[0], dist[1], dist[2 % K], r0.x);
        o4[1] = make_float4(r0.y, r0.z, r0.w, r1.x);
        o4[2] = make_float4(r1.y, r1.z, r1.w, r2.x);
        return;
    }
    for (int t = keep - 1; t >= 0; --t) o[t + k] = fresh ? 0.0f : o[t];
#pragma unroll
    for (int s = 0; s < K; ++s)
        if (s < k) o[s] = dist[s];
}

// Optional sensing noise fused into the step epilogue (north-star extension; the reference has none): the new row of
// range observations becomes clamp(d + std * z, 0, sensor_range), z ~ N(0,1) from the Philox stream (global env,
// agent + 65536 * chunk, ep_len AFTER the step, tag 3 + 4 * reset_epoch). Same arithmetic, same stream as the stand-alone
// flock_range_noise_kernel (which still serves resets and the tiled path); collisions, dones and rewards keep the true ranges.
template <int K>
__device__ __forceinline__ void add_range_noise(const Params& p, int env_global, int agent, uint32_t epoch, uint32_t repoch,
                                                float (&dist)[K]) {
    const uint32_t word = stream_word(kTagRange, repoch);
#pragma unroll
    for (int c = 0; c * 4 < K; ++c) {
        if (c * 4 >= p.k) break;
        const uint4 r = philox4x32_10((uint32_t)env_global, (uint32_t)agent + 65536u * (uint32_t)c, epoch, word, p.seed_lo, p.seed_hi);
        float z[4];
        normal2(r.x, r.y, z[0], z[1]);
        normal2(r.z, r.w, z[2], z[3]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (c * 4 + s < K && c * 4 + s < p.k) {
                const float nz = p.range_noise_std * z[s];
                float d = dist[c * 4 + s] + nz;
                d = fminf(fmaxf(d, 0.0f), p.sensor_range);
                dist[c * 4 + s] = d;
            }
        }
    }
}

// Sum over one env group of reward_fx(reward), order free and exact: rewards take at most 8 values
// (one per combination of threshold flags), so the sum is sum_c popc(lanes of class c) * fx(class c)
// with compile-time class constants -- three ballots instead of a 64-bit segmented reduction.
// Must be called by ALL 32 lanes (full-mask ballots; lanes without an agent pass false flags).
template <int V>
__device__ __forceinline__ long long group_return_fx(unsigned gmask, bool coll, bool f1, bool f2, unsigned& bc_out) {
    const unsigned bc = __ballot_sync(0xffffffffu, coll) & gmask;
    bc_out = bc;
    if (V == FLOCK_V2) {
        const int nc = __popc(bc), n = __popc(gmask);
        return (long long)nc * reward_fx(reward_from_flags<V>(true, false, false)) +
               (long long)(n - nc) * reward_fx(reward_from_flags<V>(false, false, false));
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, f1) & gmask;
    const unsigned b2 = (V == FLOCK_UW) ? (__ballot_sync(0xffffffffu, f2) & gmask) : 0u;
    long long sum = 0;
#pragma unroll
    for (int c = 0; c < ((V == FLOCK_UW) ? 8 : 4); ++c) {
        const bool cc = c & 1, c1 = c & 2, c2 = c & 4;
        unsigned m = (cc ? bc : ~bc) & (c1 ? b1 : ~b1) & gmask;
        if (V == FLOCK_UW) m &= (c2 ? b2 : ~b2);
        sum += (long long)__popc(m) * reward_fx(reward_from_flags<V>(cc, c1, c2));
    }
    return sum;
}

// all-pairs range + k-NN of row `a` against the staged positions of its env group.
// NJ4 > 0: the stride is 4*NJ4 floats and the loop is fully unrolled (straight-line code, padding
// slots of the last group skipped by warp-uniform branches); NJ4 == 0: generic runtime loop.
// SUMS: also accumulate the sequential (agent order) float32 sums of x and y that the uw centre of
// mass needs (gym_flock_uw.py:193) from the values the loop loads anyway.
template <int K, bool PER, int NJ4, bool SUMS, typename List>
__device__ __forceinline__ void knn_small(const float* sxg, const float* syg, int a, int N, int sstride, float x,
                                          float y, float B, List& t, float& sumx, float& sumy) {
    t.init();
    sumx = 0.0f;
    sumy = 0.0f;
    const float4* px = reinterpret_cast<const float4*>(sxg);
    const float4* py = reinterpret_cast<const float4*>(syg);
#define FLOCK_PAIR(XC, YC, JJ)                          \
    {                                                   \
        float d = pair_d2<PER>(x, y, XC, YC, B);        \
        d = ((JJ) == a) ? kInf : d;                     \
        t.insert(d, (JJ));                              \
        if (SUMS) {                                     \
            sumx = sumx + (XC);                         \
            sumy = sumy + (YC);                         \
        }                                               \
    }
    if (NJ4 > 0) {
#pragma unroll
        for (int j4 = 0; j4 < NJ4; ++j4) {
            const float4 X = px[j4];
            const float4 Y = py[j4];
            const int j = j4 << 2;
            const bool last = j4 == NJ4 - 1;
            FLOCK_PAIR(X.x, Y.x, j)
            if (!last || j + 1 < N) FLOCK_PAIR(X.y, Y.y, j + 1)
            if (!last || j + 2 < N) FLOCK_PAIR(X.z, Y.z, j + 2)
            if (!last || j + 3 < N) FLOCK_PAIR(X.w, Y.w, j + 3)
        }
        return;
    }
    const int n4 = sstride >> 2;
    for (int j4 = 0; j4 < n4; ++j4) {
        const float4 X = px[j4];
        const float4 Y = py[j4];
        const int j = j4 << 2;
        const bool last = j4 == n4 - 1;
        FLOCK_PAIR(X.x, Y.x, j)
        if (!last || j + 1 < N) FLOCK_PAIR(X.y, Y.y, j + 1)
        if (!last || j + 2 < N) FLOCK_PAIR(X.z, Y.z, j + 2)
        if (!last || j + 3 < N) FLOCK_PAIR(X.w, Y.w, j + 3)
    }
#undef FLOCK_PAIR
}

// stage one value per lane plus +inf / 0 padding up to the group stride
__device__ __forceinline__ void stage_xy(float* sx, float* sy, const LaneMap& m, int N, int sstride, bool live,
                                         float x, float y) {
    if (m.lane_ok) {
        sx[m.g * sstride + m.a] = live ? x : kInf;
        sy[m.g * sstride + m.a] = live ? y : 0.0f;
        if (m.a < sstride - N) {
            sx[m.g * sstride + N + m.a] = kInf;
            sy[m.g * sstride + N + m.a] = 0.0f;
        }
    }
}

// sequential float32 sum over the env's agents (canonical order 0..N-1); the scalar loop measured
// faster than a 128-bit-load version with tail guards
__device__ __forceinline__ float seq_sum(const float* s, int N) {
    float acc = 0.0f;
    for (int j = 0; j < N; ++j) acc = acc + s[j];
    return acc;
}

// -------------------------------------------------------------------------------------------------
// step: MultiAgentEnv.step of the three variants (gym_flock_v2.py:71-83, gym_flock_uw.py:69-81,
// gym_flock_uw_discrete.py:110-122).
// -------------------------------------------------------------------------------------------------
template <int K>
struct ResetResult {      // what a caller that keeps the state in registers needs back from reset_groups()
    float x, y, h;
    float dist[K];
    int idx[K];
};

template <int K>
__device__ __forceinline__ void reset_groups(const Params& p, const LaneMap& m, float* sx, float* sy, int env,
                                             size_t idx, bool want, bool keep_outputs, float* row_dst = nullptr,
                                             int* nn_dst = nullptr, ResetResult<K>* res = nullptr);

template <int K, bool IDX>
struct ListOf { using type = TopK<K>; };
template <int K>
struct ListOf<K, false> { using type = TopKValues<K>; };

// Kernel modes -- separate instantiations, because even an untaken branch costs ~3 % in the plain kernel:
//   Step       one step, actions from p.actions
//   Mirror     Step + the results also written to device-visible HOST memory (flock_step_host zero-copy path)
//   AutoReset  Step + env groups whose step ended with a collision are re-drawn in the same launch (reward and done
//              flags of the finishing step stay, obs becomes the first observation of the new episode)
//   Noise      Step + Philox range-sensing noise on the stored observation row (range_noise_std > 0)
//   Multi      flock_step_n: p.num_steps steps, state in registers, canonical Philox random actions, outputs = last step
//   Rollout    flock_rollout_n: Multi + actions read from and obs | reward | dones written to time-major trajectory
//              buffers every step, optional in-kernel auto-reset (p.fused_auto_reset)
enum : int { kModeStep = kSmallModeStep, kModeMirror = kSmallModeMirror, kModeAutoReset = kSmallModeAutoReset, kModeMulti = kSmallModeMulti,
             kModeRollout = kSmallModeRollout, kModeNoise = kSmallModeNoise };

// IDX = false: neighbour indices are not tracked (values-only selection network).
template <int V, int K, bool PER, int MODE, int NJ4, bool IDX>
__global__ void __launch_bounds__(kSmallThreads) flock_step_small_kernel(const __grid_constant__ Params p) {
    constexpr bool MULTI = MODE == kModeMulti || MODE == kModeRollout;
    constexpr bool ROLLOUT = MODE == kModeRollout;
    constexpr bool MIRROR = MODE == kModeMirror;
    constexpr bool AUTORESET = MODE == kModeAutoReset;
    constexpr bool NOISE = MODE == kModeNoise;
#ifdef FLOCK_TIMELINE
    const unsigned long long tl_entry = timeline_now();
#endif
    __shared__ __align__(16) float s_stage[kSmallWarps][3][kSlots];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* sx = s_stage[wib][0];
    float* sy = s_stage[wib][1];
    float* sh = s_stage[wib][2];
    // The unrolled instantiations serve the BASELINE swarm sizes EXACTLY (NJ4 = 3: N = 10, 4: N = 16, 8: N = 32; other sizes
    // take the generic loop, NJ4 = 0): N, the envs per warp, the stride and the lane map are compile-time constants there,
    // which removes ~40 instructions of the one-task-per-warp preamble and the tail guards of the pair loop.
    constexpr int NX = NJ4 == 3 ? 10 : NJ4 == 4 ? 16 : NJ4 == 8 ? 32 : 0;
    const int N = NX > 0 ? NX : p.N, G = NX > 0 ? 32 / NX : p.G, k = p.k, sstride = NX > 0 ? 4 * NJ4 : p.sstride;
    const LaneMap m = lane_map(lane, N, G, NX > 0 ? (65536 + NX - 1) / NX : p.g_magic);
    const int num_tasks = p.num_tasks;
    const int warps_total = gridDim.x * kSmallWarps;
    const int nsteps = MULTI ? p.num_steps : 1;
    const int GN = G * N;
    const unsigned EN = (unsigned)p.E * (unsigned)N;
#if FLOCK_PDL_PREFETCH
    if (!MULTI) {
        // Launched as a programmatic dependent, the CTA sits in griddepcontrol.wait for 0.3-0.5 us while the previous kernel
        // of the stream drains (tools/cta_timeline.py). Loads must wait -- the previous kernel may be the one that writes
        // our inputs -- but an L2 PREFETCH of the lines this warp is about to read is always safe (L2 is the point of
        // coherence: a later write by the previous kernel lands in the same line) and turns the DRAM round trip after the
        // wait into L2 hits when the state is cold.
        const int task0 = blockIdx.x * kSmallWarps + wib;
        const unsigned flat0 = (unsigned)task0 * (unsigned)GN + (unsigned)lane;
        if (task0 < num_tasks && lane < GN && flat0 < EN) {
            prefetch_l2(p.x + flat0);
            prefetch_l2(p.y + flat0);
            prefetch_l2(p.h + flat0);
            if (V == FLOCK_UWD) {
                prefetch_l2(p.actions + flat0);
                if (p.noise != nullptr) prefetch_l2(p.noise + 2 * (size_t)flat0);
            } else {
                prefetch_l2(p.actions + 2 * (size_t)flat0);
            }
        }
    }
#endif
    pdl_wait_prior_grid();
#ifdef FLOCK_TIMELINE
    const unsigned long long tl_go = timeline_now();
#endif

    const bool ring = (V == FLOCK_UW) && p.obs_head != nullptr;
    for (int task = blockIdx.x * kSmallWarps + wib; task < num_tasks; task += warps_total) {
        // the warp's agents are one contiguous run: index = task*G*N + lane, so the loads below can
        // issue after two integer operations (the env / agent split is only needed later).
        // (Prefetching the warp's next task with two tasks per warp was measured slower on B200 for
        // all BASELINE configs: at these batch sizes more resident warps beat software pipelining.)
        const unsigned flat = (unsigned)task * (unsigned)GN + (unsigned)lane;
        const bool live = lane < GN && flat < EN;
        const size_t idx = live ? flat : 0u;
        const int env = live ? task * G + m.g : 0;
        float x = 0.f, y = 0.f, h = 0.f, prev_h = 0.f;
        float prev_h_in = 0.f;
        uint32_t ep_base = 0u, repoch = 0u;   // per-env Philox epoch (episode step of this launch's step 0, reset epoch)
        // Every global load of the step is issued here, up front, so the warp pays ONE DRAM round
        // trip: the read-modify-write operands of the epilogue (episode counters, uw obs window)
        // are prefetched into registers together with the state and the actions.
        long long ep_ret0 = 0;
        float act0 = 0.f, act1 = 0.f, nz0 = 0.f, nz1 = 0.f;
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0, w2 = w0;   // uw window (k = 3 fast path)
        const bool fast_win = (V == FLOCK_UW) && !MULTI && !ring && k == 3 && p.H == 4;
        int head = 0;                          // uw ring: slot of the newest row before this launch
        if (live) {
            ep_base = (uint32_t)p.ep_len[env];
            if (MULTI || NOISE || V == FLOCK_UWD) repoch = p.reset_epoch[env];
            if (m.a == 0 && p.ep_return_fx != nullptr) ep_ret0 = p.ep_return_fx[env];
            x = p.x[idx];
            y = p.y[idx];
            h = p.h[idx];
            if (V == FLOCK_UW) prev_h = prev_h_in = p.prev_h[idx];
            if (ring) head = p.obs_head[env];
            if (!MULTI) {
                if (V == FLOCK_UWD) {
                    act0 = p.actions[idx];
                    if (p.noise != nullptr) {
                        const float2 nz = reinterpret_cast<const float2*>(p.noise)[idx];
                        nz0 = nz.x;
                        nz1 = nz.y;
                    }
                } else {
                    const float2 act = reinterpret_cast<const float2*>(p.actions)[idx];
                    act0 = act.x;
                    act1 = act.y;
                }
                if (fast_win) {
                    const float4* o4 = reinterpret_cast<const float4*>(p.obs + idx * 12);
                    w0 = o4[0];
                    w1 = o4[1];
                    w2 = o4[2];
                }
            } else if (ROLLOUT && p.traj_actions != nullptr) {   // actions of step 0; step t+1 is prefetched during step t
                if (V == FLOCK_UWD) {
                    act0 = p.traj_actions[idx];
                } else {
                    const float2 act = reinterpret_cast<const float2*>(p.traj_actions)[idx];
                    act0 = act.x;
                    act1 = act.y;
                }
            }
        }
        float vx = 0.f, vy = 0.f, rew = 0.f;
        float dist[K];
        using List = typename ListOf<K, IDX>::type;
        List t;
        bool coll = false, env_coll = false;
        long long ret_fx = 0;
        float hist[(V == FLOCK_UW) ? 3 * K : 1];   // rows 1..3 of the window after the current step (MULTI)
        if (MULTI && V == FLOCK_UW && live) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float* o = ring ? ring_row(p, env, (head + r) % p.H, m.a) : p.obs + idx * (size_t)(4 * k) + r * k;
#pragma unroll
                for (int s = 0; s < K; ++s) hist[r * K + s] = (s < k) ? o[s] : 0.0f;
            }
        }

        for (int st = 0; st < nsteps; ++st) {
            float a0 = 0.f, a1 = 0.f, nzu = 0.f, nzw = 0.f;
            if (live) {
                const uint32_t step = ep_base + (uint32_t)st;
                if (ROLLOUT && p.traj_actions != nullptr) {
                    a0 = act0;
                    a1 = act1;
                    if (st + 1 < nsteps) {   // prefetch the next step's action under this step's arithmetic
                        const size_t nidx = (size_t)(st + 1) * EN + idx;
                        if (V == FLOCK_UWD) {
                            act0 = p.traj_actions[nidx];
                        } else {
                            const float2 act = reinterpret_cast<const float2*>(p.traj_actions)[nidx];
                            act0 = act.x;
                            act1 = act.y;
                        }
                    }
                } else if (MULTI) {
                    random_action<V>(p, p.env_offset + env, m.a, step, repoch, a0, a1);
                } else {
                    a0 = act0;
                    a1 = act1;
                }
                if (V == FLOCK_UWD) {
                    if (!MULTI && p.noise != nullptr) {
                        nzu = nz0;
                        nzw = nz1;
                    } else if (p.noise_std > 0.0f) {
                        act_noise(p, p.env_offset + env, m.a, step, repoch, nzu, nzw);
                    }
                }
                integrate_agent<V>(p, a0, a1, nzu, nzw, x, y, h, vx, vy);
            }
            __syncwarp();
            stage_xy(sx, sy, m, N, sstride, live, x, y);
            if (V == FLOCK_UWD && m.lane_ok) sh[m.g * sstride + m.a] = h;
            __syncwarp();
            bool f1 = false, f2 = false;
            coll = false;
            if (live) {
                const float* sxg = sx + m.g * sstride;
                const float* syg = sy + m.g * sstride;
                float comx = 0.f, comy = 0.f, hmean = 0.f, sumx, sumy;
                if (V == FLOCK_UWD) hmean = mean_of_sum(p, seq_sum(sh + m.g * sstride, N));  // uwd:256
                if (MULTI && V == FLOCK_UW && st > 0) {
#pragma unroll
                    for (int s = 3 * K - 1; s >= K; --s) hist[s] = hist[s - K];   // K-strided shift of the window
#pragma unroll
                    for (int s = 0; s < K; ++s) hist[s] = dist[s];
                }
                knn_small<K, PER, NJ4, V == FLOCK_UW, List>(sxg, syg, m.a, N, sstride, x, y, p.B, t, sumx, sumy);
                if (V != FLOCK_UW && !MULTI && !FLOCK_LATE_TRIGGER) pdl_launch_dependents();   // the next kernel may start launching (and prefetching)
                if (V == FLOCK_UW) {  // torch.mean(positions, 0), gym_flock_uw.py:193
                    comx = mean_of_sum(p, sumx);
                    comy = mean_of_sum(p, sumy);
                }
                coll = finish_row<K, List>(t, k, p.sensor_range, p.cd, dist);
                reward_flags<V>(p, x, y, h, prev_h, comx, comy, hmean, f1, f2);
                rew = reward_from_flags<V>(coll, f1, f2);
                if (V == FLOCK_UW) prev_h = h;
            }
            // env-level reductions: warp-wide ballots, each env reads its own lane group
            unsigned bc;
            ret_fx += group_return_fx<V>(m.gmask, live && coll, live && f1, live && f2, bc);
            env_coll = bc != 0u;
            if (ROLLOUT) {
                // this step's slice of the trajectory: obs row | reward | agent_done | env_done (| nn)
                const size_t tidx = (size_t)st * EN + idx;
                if (live) {
                    store_row<float, K>(p.traj_obs + tidx * k, dist, k);
                    if constexpr (IDX) {
                        if (p.traj_nn != nullptr) store_row<int, K>(p.traj_nn + tidx * k, t.idx, k);
                    }
                    p.traj_reward[tidx] = rew;
                    p.traj_agent_done[tidx] = coll ? 1 : 0;
                    if (m.a == 0) p.traj_env_done[(size_t)st * p.E + env] = env_coll ? 1 : 0;
                }
                const bool want = live && env_coll && p.fused_auto_reset != 0;
                if (__any_sync(0xffffffffu, want)) {
                    // the batched "caller resets on done[1]" (main.py:31-51) inside the rollout: close the episode
                    // counters of the finished envs in memory, re-draw them, and continue from the new state
                    if (want && m.a == 0) {
                        p.ep_len[env] = (int)(ep_base + (uint32_t)st + 1u);
                        if (p.ep_return_fx != nullptr) p.ep_return_fx[env] = ep_ret0 + ret_fx;
                    }
                    __syncwarp();
                    ResetResult<K> rr;
                    reset_groups<K>(p, m, sx, sy, env, idx, want, true, p.traj_obs + tidx * k,
                                    (IDX && p.traj_nn != nullptr) ? p.traj_nn + tidx * k : nullptr, &rr);
                    if (want) {
                        x = rr.x;
                        y = rr.y;
                        h = rr.h;
                        prev_h = 0.0f;
                        vx = 0.0f;
                        vy = 0.0f;
#pragma unroll
                        for (int s = 0; s < K; ++s) dist[s] = rr.dist[s];
                        if constexpr (IDX) {
#pragma unroll
                            for (int s = 0; s < K; ++s) t.idx[s] = rr.idx[s];
                        }
                        if (V == FLOCK_UW) {
#pragma unroll
                            for (int s = 0; s < 3 * K; ++s) hist[s] = 0.0f;
                        }
                        ep_base = 0u - (uint32_t)(st + 1);      // episode step 0 at launch step st + 1
                        repoch = p.reset_epoch[env];
                        ep_ret0 = 0;
                        ret_fx = 0;
                    }
                }
            }
        }

        // Programmatic dependent launch trigger: v2 / uwd fire it right after the pair loop (see there), uw here, before
        // the result stores (measured with the pre-wait prefetch in place: the early trigger gives cfg2 2.86 -> 2.78 us and
        // cfg4 4.97 -> 4.88 us, but cfg3 5.22 -> 5.40 us: the uw epilogue is long and the waiting dependents get in its way)
        if (V == FLOCK_UW || MULTI || FLOCK_LATE_TRIGGER) pdl_launch_dependents();
        if (live) {
            p.xo[idx] = x;
            p.yo[idx] = y;
            if (V != FLOCK_UW) p.ho[idx] = h;
            if (V == FLOCK_UW && !(prev_h == prev_h_in)) p.prev_h[idx] = prev_h;   // constant after the first step
            if (p.vx != nullptr) {
                p.vx[idx] = vx;
                p.vy[idx] = vy;
            }
            if (NOISE) add_range_noise<K>(p, p.env_offset + env, m.a, ep_base + (uint32_t)nsteps, repoch, dist);
            if (MULTI && V == FLOCK_UW) {
                // window after the launch: newest first = dist, then the K-strided register history
                const int new_head = ring ? (head + p.H - (nsteps % p.H)) % p.H : 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float* o = ring ? ring_row(p, env, (new_head + r) % p.H, m.a) : p.obs + idx * (size_t)(4 * k) + r * k;
#pragma unroll
                    for (int s = 0; s < K; ++s)
                        if (s < k) o[s] = (r == 0) ? dist[s] : hist[(r == 0 ? 0 : r - 1) * K + s];
                }
                if (ring && m.a == 0) p.obs_head[env] = new_head;
            } else if (ring) {
                const int new_head = ring_prev_slot(head, p.H);
                float* o = ring_row(p, env, new_head, m.a);
#pragma unroll
                for (int s = 0; s < K; ++s)
                    if (s < k) o[s] = dist[s];
                if (m.a == 0) p.obs_head[env] = new_head;
            } else if (fast_win) {
                float4* o4 = reinterpret_cast<float4*>(p.obs + idx * 12);   // shift by one row of 3, newest first
                o4[0] = make_float4(dist[0], dist[1], dist[2 % K], w0.x);
                o4[1] = make_float4(w0.y, w0.z, w0.w, w1.x);
                o4[2] = make_float4(w1.y, w1.z, w1.w, w2.x);
            } else if (V != FLOCK_UW) {
                store_row<float, K>(p.obs + idx * k, dist, k);      // H == 1
            } else {
                write_obs<K>(p, env, m.a, idx, dist, false, 0);
            }
            if constexpr (IDX) {
                if (p.nn != nullptr) store_row<int, K>(p.nn + idx * k, t.idx, k);
            }
            p.reward[idx] = rew;
            p.agent_done[idx] = coll ? 1 : 0;
            if (m.a == 0) {
                p.env_done[env] = env_coll ? 1 : 0;
                if (p.ep_return_fx != nullptr) p.ep_return_fx[env] = ep_ret0 + ret_fx;
                p.ep_len[env] = (int)(ep_base + (uint32_t)nsteps);
            }
            if (MIRROR) {
                // host-call path: the results also go straight to mapped host memory (posted PCIe
                // writes overlap the rest of the kernel; no separate device->host copy)
                if (fast_win) {
                    float4* o4 = reinterpret_cast<float4*>(p.m_obs + idx * 12);
                    o4[0] = make_float4(dist[0], dist[1], dist[2 % K], w0.x);
                    o4[1] = make_float4(w0.y, w0.z, w0.w, w1.x);
                    o4[2] = make_float4(w1.y, w1.z, w1.w, w2.x);
                } else if (V != FLOCK_UW) {
                    store_row<float, K>(p.m_obs + idx * k, dist, k);
                } else {   // generic window (the ring layout never takes the zero-copy path)
                    const size_t hk = (size_t)p.H * k;
                    for (size_t u = 0; u < hk; ++u) p.m_obs[idx * hk + u] = p.obs[idx * hk + u];
                }
                p.m_reward[idx] = rew;
                p.m_agent_done[idx] = coll ? 1 : 0;
                if (m.a == 0) p.m_env_done[env] = env_coll ? 1 : 0;
            }
        }
        if (AUTORESET) {
            const bool want = live && env_coll;
            if (__any_sync(0xffffffffu, want)) reset_groups<K>(p, m, sx, sy, env, idx, want, true);
        }
    }
#ifdef FLOCK_TIMELINE
    if (threadIdx.x == 0 && blockIdx.x < kTimelineCtas) {
        unsigned long long* d = p.timeline + (size_t)blockIdx.x * kTimelineSlots;
        d[0] = tl_entry;
        d[1] = tl_go;
        d[2] = timeline_now();
        d[3] = timeline_smid();
    }
#endif
}

// -------------------------------------------------------------------------------------------------
// reset: MultiAgentEnv.reset (gym_flock_v2.py:85-108, gym_flock_uw.py:83-111,
// gym_flock_uw_discrete.py:124-156) with a BOUNDED rejection loop, masked and batched.
// reset_groups() (re)draws the env groups of the calling warp whose lanes pass `want` (uniform per
// group); ALL 32 lanes must call it. It is the body of the reset kernel and, with keep_outputs, the
// fused auto-reset tail of the step / rollout kernels (row_dst / nn_dst: the first observation row and
// neighbour list go to a trajectory slice instead of the env's buffers; res: the new state for a caller
// that keeps it in registers).
// -------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void reset_groups(const Params& p, const LaneMap& m, float* sx, float* sy, int env,
                                             size_t idx, bool want, bool keep_outputs, float* row_dst, int* nn_dst,
                                             ResetResult<K>* res) {
    const int N = p.N, k = p.k, sstride = p.sstride;
    const size_t EN = (size_t)p.E * N;
    float x = 0.f, y = 0.f, h = 0.f;
    const uint32_t epoch = want ? p.reset_epoch[env] : 0u;
    uint32_t attempts = 0;
    bool need = want;       // group still needs a (re)draw
    bool coll = false, env_coll = false;
    float dist[K];
    TopK<K> t;
    const int max_att = p.init_state != nullptr ? 1 : p.max_attempts;
    while (__any_sync(0xffffffffu, need)) {
        if (need) {
            if (p.init_state != nullptr) {
                x = p.init_state[idx];
                y = p.init_state[EN + idx];
                h = p.init_state[2 * EN + idx];
            } else {
                const uint4 r = philox4x32_10((uint32_t)(p.env_offset + env), (uint32_t)m.a, epoch + attempts, kTagReset,
                                              p.seed_lo, p.seed_hi);
                const float span = p.range_lo - p.reset_hi;      // (r0 - r1) * U + r1, gym_flock_v2.py:87-89
                const float tx = span * u24(r.x);
                x = tx + p.reset_hi;
                const float ty = span * u24(r.y);
                y = ty + p.reset_hi;
                const float th = (0.0f - p.heading_hi) * u24(r.z);  // gym_flock_v2.py:96
                h = th + p.heading_hi;
            }
            x = wrap_coord(x, p.B, p.fill_hi, p.fill_lo);                      // check_boundary, v2:99
            y = wrap_coord(y, p.B, p.fill_hi, p.fill_lo);
            attempts += 1;
        }
        __syncwarp();
        stage_xy(sx, sy, m, N, sstride, want, x, y);
        __syncwarp();
        if (need) {   // whole group shares `need`
            float unused_sx, unused_sy;
            knn_small<K, false, 0, false, TopK<K>>(sx + m.g * sstride, sy + m.g * sstride, m.a, N, sstride, x, y, p.B, t,
                                          unused_sx, unused_sy);  // Euclidean, v2:100
            coll = finish_row<K, TopK<K>>(t, k, p.sensor_range, p.reset_cd, dist);
        }
        {
            const unsigned bc = __ballot_sync(0xffffffffu, need && coll) & m.gmask;   // full-mask ballot
            if (need) {
                env_coll = bc != 0u;
                need = env_coll && (int)attempts < max_att;
            }
        }
    }
    if (want) {
        p.xo[idx] = x;
        p.yo[idx] = y;
        p.ho[idx] = h;
        p.prev_h[idx] = 0.0f;                                     // v2:95
        if (p.vx != nullptr) {
            p.vx[idx] = 0.0f;                                     // v2:94
            p.vy[idx] = 0.0f;
        }
        if (row_dst != nullptr) {
            store_row<float, K>(row_dst, dist, k);
        } else {
            write_obs<K>(p, env, m.a, idx, dist, true, p.obs_head != nullptr ? p.obs_head[env] : 0);
        }
        if (nn_dst != nullptr) store_row<int, K>(nn_dst, t.idx, k);
        else if (row_dst == nullptr && p.nn != nullptr) store_row<int, K>(p.nn + idx * k, t.idx, k);
        if (res != nullptr) {
            res->x = x;
            res->y = y;
            res->h = h;
#pragma unroll
            for (int s = 0; s < K; ++s) {
                res->dist[s] = dist[s];
                res->idx[s] = t.idx[s];
            }
        }
        if (!keep_outputs) {
            p.reward[idx] = 0.0f;
            p.agent_done[idx] = coll ? 1 : 0;
        }
        if (m.a == 0) {
            if (!keep_outputs) p.env_done[env] = env_coll ? 1 : 0;
            if (p.init_state == nullptr) p.reset_epoch[env] = epoch + attempts;
            const int len = p.ep_len[env];
            if (p.stats != nullptr) {
                if (len > 0) {
                    atomicAdd(&p.stats[FLOCK_STAT_EPISODES], 1ULL);
                    atomicAdd(&p.stats[FLOCK_STAT_EP_STEPS], (unsigned long long)len);
                    if (p.ep_return_fx != nullptr)
                        atomicAdd(&p.stats[FLOCK_STAT_EP_RETURN_FX], (unsigned long long)p.ep_return_fx[env]);
                }
                if (p.init_state == nullptr) {
                    atomicAdd(&p.stats[FLOCK_STAT_RESET_ATTEMPTS], (unsigned long long)attempts);
                    if (env_coll) atomicAdd(&p.stats[FLOCK_STAT_RESET_GAVE_UP], 1ULL);
                }
            }
            p.ep_len[env] = 0;
            if (p.ep_return_fx != nullptr) p.ep_return_fx[env] = 0;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// host-side dispatch
// -------------------------------------------------------------------------------------------------
inline int small_grid(const Params& p, int sm_count) {
    const int tasks = p.num_tasks;
    const int blocks = (tasks + kSmallWarps - 1) / kSmallWarps;
    const int cap = sm_count * 32;     // 32 resident 64-thread CTAs per SM
    return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

template <int V, int K, bool PER, int NJ4, bool IDX>
cudaError_t launch_step_small_vkpni(const Params& p, int mode, int sm_count, cudaStream_t s) {
    const int grid = small_grid(p, sm_count);
    switch (mode) {
        case kModeMulti:
            flock_step_small_kernel<V, K, PER, kModeMulti, NJ4, IDX><<<grid, kSmallThreads, 0, s>>>(p);
            return cudaGetLastError();
        case kModeRollout:
            flock_step_small_kernel<V, K, PER, kModeRollout, NJ4, IDX><<<grid, kSmallThreads, 0, s>>>(p);
            return cudaGetLastError();
        case kModeAutoReset:   // step + restart of the finished envs in one launch
            flock_step_small_kernel<V, K, PER, kModeAutoReset, NJ4, IDX><<<grid, kSmallThreads, 0, s>>>(p);
            return cudaGetLastError();
        default: break;
    }
    // Programmatic dependent launch (default on, FLOCK_PDL=0 disables): the kernel triggers its
    // dependents right before its epilogue, so the next step's launch overlaps our result stores.
    static const bool use_pdl = [] {
        const char* v = getenv("FLOCK_PDL");
        return v == nullptr || v[0] != '0';
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kSmallThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    if (mode == kModeMirror) return cudaLaunchKernelEx(&cfg, flock_step_small_kernel<V, K, PER, kModeMirror, NJ4, IDX>, p);
    if (mode == kModeNoise) return cudaLaunchKernelEx(&cfg, flock_step_small_kernel<V, K, PER, kModeNoise, NJ4, IDX>, p);
    return cudaLaunchKernelEx(&cfg, flock_step_small_kernel<V, K, PER, kModeStep, NJ4, IDX>, p);
}

// unrolled pair loops for the swarm sizes of the BASELINE configs (N = 10, 16, 32), the generic loop for all others
template <int V, int K, bool PER, bool IDX>
cudaError_t launch_step_small_vkpi(const Params& p, int mode, int sm_count, cudaStream_t s) {
    switch (p.sstride) {
        case 12: if (p.N == 10 && p.G == 3) return launch_step_small_vkpni<V, K, PER, 3, IDX>(p, mode, sm_count, s); break;
        case 16: if (p.N == 16 && p.G == 2) return launch_step_small_vkpni<V, K, PER, 4, IDX>(p, mode, sm_count, s); break;
        case 32: if (p.N == 32 && p.G == 1) return launch_step_small_vkpni<V, K, PER, 8, IDX>(p, mode, sm_count, s); break;
        default: break;
    }
    return launch_step_small_vkpni<V, K, PER, 0, IDX>(p, mode, sm_count, s);
}

// one explicit instantiation of this per translation unit (flock_small_<variant>.cu)
template <int V, bool PER, bool IDX>
cudaError_t launch_step_small_vpi(const Params& p, int mode, int sm_count, cudaStream_t s) {
    if (p.k <= 3) return launch_step_small_vkpi<V, 3, PER, IDX>(p, mode, sm_count, s);
    if (p.k == 4) return launch_step_small_vkpi<V, 4, PER, IDX>(p, mode, sm_count, s);
    return launch_step_small_vkpi<V, 8, PER, IDX>(p, mode, sm_count, s);
}

}  // namespace flock
