// flock_tiled.cu -- tiled all-pairs kernels for large swarms (32 < N <= FLOCK_MAX_AGENTS).
//
// One CTA = one i-tile (R = blockDim.x rows, chosen per launch for SM balance) of one env. The
// env's old positions are pulled into shared memory with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier); every CTA of the env re-integrates all N agents in place in shared memory (headings
// and actions stream straight from global; ~5 % of the pair work at N = 2048, and it removes any
// grid-wide dependency), writes back only its own rows to the OTHER state copy (ping-pong, so
// concurrent CTAs of the same env still read the old state), then runs its rows against all j as
// broadcast 128-bit shared-memory reads with a register-resident k-smallest list per row.
// One launch = one env step, as on the small path.
#include <cstdlib>

#include "flock_device.cuh"
#include "flock_launch.h"

namespace flock {

static bool use_rowwarp(const Params& p, int sm_count, int tiled_mode);
static int choose_rows(int N, int E, int sm_count);

constexpr int kMaxTileThreads = 256;   // upper bound of rows per CTA
constexpr int kCandCap = 8;         // deferred k-NN candidates per row between two merges

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA: global -> shared bulk copy, completion signalled on the mbarrier (16-byte granularity)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    // volatile + memory clobber: must stay after the __syncthreads() that publishes the staged env
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}

__host__ __device__ __forceinline__ int padded_agents(int N) { return ((N + 3) & ~3) + 4; }

// warp-shuffle bitonic top-k merge of 32 per-lane sorted lists (defined with the warp-per-row kernel below)
template <int K>
__device__ __forceinline__ void warp_bitonic_topk(unsigned long long (&key)[K], int width = 32);
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src);

// shared-memory carve-up: sx | sy | [sh] | candidate buffer. sh (headings) exists only where the
// kernel needs every agent's heading afterwards (uwd mean heading, reset).
size_t tiled_smem_bytes(int num_agents, int rows, bool need_sh) {
    return (size_t)padded_agents(num_agents) * (need_sh ? 3 : 2) * sizeof(float) +
           (size_t)kCandCap * rows * sizeof(uint2);
}

struct TileSmem {
    float *sx, *sy, *sh;
    uint2* cand;   // [kCandCap][rows] (d2 bits, j), column = thread
};
__device__ __forceinline__ TileSmem carve(float* base, int N, bool need_sh) {
    const int P = padded_agents(N);
    float* after = base + (need_sh ? 3 : 2) * P;
    return TileSmem{base, base + P, need_sh ? base + 2 * P : nullptr, reinterpret_cast<uint2*>(after)};
}

// stage the env's positions: TMA bulk copies when rows are 16-byte multiples, plain loads otherwise
__device__ __forceinline__ void stage_xy(const Params& p, const TileSmem& sm, uint64_t* bar, int env) {
    const int N = p.N;
    const size_t base = (size_t)env * N;
    if ((N & 3) == 0) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            const uint32_t row = (uint32_t)N * 4u;
            mbar_expect_tx(bar, 2u * row);
            tma_load_1d(sm.sx, p.x + base, row, bar);
            tma_load_1d(sm.sy, p.y + base, row, bar);
        }
        __syncthreads();          // barrier initialised and armed before anyone polls it
        mbar_wait(bar, 0);
    } else {
        for (int a = threadIdx.x; a < N; a += blockDim.x) {
            sm.sx[a] = p.x[base + a];
            sm.sy[a] = p.y[base + a];
        }
        __syncthreads();
    }
}

// All-pairs scan of one row against the staged env, ALL 32 lanes of the warp must call it.
//
// The hot loop is 9 FP32 instructions per pair plus a warp-uniform threshold vote per 4 pairs;
// candidates with d2 <= thr are only APPENDED (predicated) to a per-thread shared-memory buffer and
// merged into the register-resident sorted k-list in batches, so the 40-instruction sorted insert
// is off the hot path. `thr` is any upper bound of the row's k-th smallest d2 (the caller derives
// it from the previous step's neighbour list; FLT_MAX = no bound; negative = dummy row, never
// hits). Candidates reach the list in ascending j and the insert is strict, so ties keep the lower
// index; the row's own index is dropped at merge time, which keeps the j == i test out of the loop.
template <int K, bool PER>
__device__ __forceinline__ void knn_tiled(const float* sx, const float* sy, uint2* cand, int i, int N, float x,
                                          float y, float B, float thr, TopK<K>& t) {
    constexpr unsigned kFull = 0xffffffffu;
    const int cstride = blockDim.x;
    t.init();
    int cnt = 0;
    auto merge = [&]() {
#pragma unroll 1
        for (int c = 0; c < kCandCap; ++c) {
            if (!__any_sync(kFull, c < cnt)) break;
            const uint2 e = cand[c * cstride];
            const bool ok = c < cnt && (int)e.y != i;
            t.insert(ok ? __uint_as_float(e.x) : kInf, (int)e.y);
        }
        cnt = 0;
        thr = fminf(thr, t.worst());
    };
    // explicit shared-window addresses: one register per array, bumped by 16 bytes per iteration
    uint32_t ax = smem_u32(sx), ay = smem_u32(sy);
    const int n4 = (N + 3) >> 2;
#pragma unroll 2
    for (int j4 = 0; j4 < n4; ++j4, ax += 16u, ay += 16u) {
        const float4 X = lds128(ax);
        const float4 Y = lds128(ay);
        const float d0 = pair_d2<PER>(x, y, X.x, Y.x, B);
        const float d1 = pair_d2<PER>(x, y, X.y, Y.y, B);
        const float d2 = pair_d2<PER>(x, y, X.z, Y.z, B);
        const float d3 = pair_d2<PER>(x, y, X.w, Y.w, B);
        const float m = fminf(fminf(d0, d1), fminf(d2, d3));
        if (__any_sync(kFull, m <= thr)) {
            // (one vote per candidate instead of four predicated appends was measured slower)
            const int j = j4 << 2;
            if (d0 <= thr) { cand[cnt * cstride] = make_uint2(__float_as_uint(d0), (unsigned)j); ++cnt; }
            if (d1 <= thr) { cand[cnt * cstride] = make_uint2(__float_as_uint(d1), (unsigned)(j + 1)); ++cnt; }
            if (d2 <= thr) { cand[cnt * cstride] = make_uint2(__float_as_uint(d2), (unsigned)(j + 2)); ++cnt; }
            if (d3 <= thr) { cand[cnt * cstride] = make_uint2(__float_as_uint(d3), (unsigned)(j + 3)); ++cnt; }
            if (__any_sync(kFull, cnt > kCandCap - 4)) merge();
        }
    }
    merge();
}

// Upper bound of the k-th smallest d2 of row i from a previous neighbour list (any k distinct valid
// indices != i bound it): max over their CURRENT distances. Returns FLT_MAX when the list is not a
// set of k distinct in-range indices (first use, foreign data), so a bad hint can never change the
// result, only the speed.
template <int K, bool PER>
__device__ __forceinline__ float hint_bound(const int (&hint)[K], int k, int i, int N, const float* sx,
                                            const float* sy, float x, float y, float B) {
    bool ok = true;
    float bound = 0.0f;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        if (s < k) {
            const int j = hint[s];
            const bool in = j >= 0 && j < N && j != i;
            ok = ok && in;
#pragma unroll
            for (int u = 0; u < s; ++u) ok = ok && (hint[u] != j);
            const int jj = in ? j : 0;
            bound = fmaxf(bound, pair_d2<PER>(x, y, sx[jj], sy[jj], B));
        }
    }
    return ok ? bound : kFltMax;
}

template <typename T, int MAXN>
__device__ __forceinline__ void store_row_t(T* dst, const T (&v)[MAXN], int n) {
    struct alignas(16) Vec4 { T a, b, c, d; };
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

template <int K>
__device__ __forceinline__ void write_obs_t(const Params& p, int env, int a, size_t idx, const float (&dist)[K], bool fresh) {
    const int k = p.k;
    if (p.H == 1) {
        store_row_t<float, K>(p.obs + idx * k, dist, k);
        return;
    }
    if (p.obs_head != nullptr) {   // uw ring layout (flock_device.cuh): only the new row is written; a reset zero-fills the rest
        const int head = p.obs_head[env];                       // published by the env's last CTA AFTER all rows are written
        const int slot = fresh ? head : ring_prev_slot(head, p.H);
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
    if (p.H == 4) {
        // the reference's history depth: all old rows are loaded into registers BEFORE the first store (the compiler must
        // keep a load behind every earlier store to this overlapping window, which made the element-wise shift nine
        // serialised memory round trips per agent: 147 -> 118 us for the sensing kernel at 64 x 2048)
        float old[3 * K];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < K; ++s) old[r * K + s] = (s < k && !fresh) ? o[r * k + s] : 0.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < K; ++s)
                if (s < k) o[(r + 1) * k + s] = old[r * K + s];
    } else {
        for (int t = (p.H - 1) * k - 1; t >= 0; --t) o[t + k] = fresh ? 0.0f : o[t];
    }
#pragma unroll
    for (int s = 0; s < K; ++s)
        if (s < k) o[s] = dist[s];
}

// step, grid = (tiles per env, E), blockDim.x = rows per tile.
template <int V, int K, bool PER>
__global__ void __launch_bounds__(kMaxTileThreads) flock_step_tiled_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int N = p.N, k = p.k;
    const int env = blockIdx.y, tile = blockIdx.x;
    const int rows = blockDim.x;
    const TileSmem sm = carve(smem, N, false);
    const size_t base = (size_t)env * N;
    // Row assignment. With a row order (p.perm, refreshed every few steps by
    // flock_perm_refresh_kernel) the 32 rows of a warp are spatial neighbours, so their k-NN
    // candidates coincide and the candidate-append path below is entered far less often. Any
    // permutation gives the same results: every row is computed independently.
    const int slot = tile * rows + threadIdx.x;
    const bool has_row = slot < N;
    const int i = has_row ? (p.perm != nullptr ? p.perm[base + slot] : slot) : N;
    // last step's neighbour list of this thread's row: the hint that bounds this step's k-th distance
    int hint[K];
#pragma unroll
    for (int s = 0; s < K; ++s) hint[s] = (p.nn != nullptr && has_row && s < k) ? p.nn[(base + i) * k + s] : -1;
    // the state was integrated in place by the pre-pass (flock_integrate_kernel): stage the NEW positions
    stage_xy(p, sm, &bar, env);
    if (threadIdx.x < padded_agents(N) - N) {
        sm.sx[N + threadIdx.x] = kInf;
        sm.sy[N + threadIdx.x] = 0.0f;
    }
    __syncthreads();
    float x = 0.f, y = 0.f, h = 0.f;
    if (has_row) {
        x = sm.sx[i];
        y = sm.sy[i];
        h = p.h[base + i];
    }

    // per-env sums of the new state, formed once by the integrate pre-pass (see flock_integrate_kernel)
    float comx = 0.f, comy = 0.f, hmean = 0.f;
    if (V == FLOCK_UW) {
        comx = mean_of_sum(p, p.env_sums[2 * env]);
        comy = mean_of_sum(p, p.env_sums[2 * env + 1]);
    }
    if (V == FLOCK_UWD) hmean = mean_of_sum(p, p.env_sums[2 * env]);

    long long fx = 0;
    bool coll = false;
    float thr = -1.0f;                      // rows past N never hit
    if (has_row) thr = hint_bound<K, PER>(hint, k, i, N, sm.sx, sm.sy, x, y, p.B);
    TopK<K> t;
    knn_tiled<K, PER>(sm.sx, sm.sy, sm.cand + threadIdx.x, i, N, x, y, p.B, thr, t);
    if (has_row) {
        float dist[K];
        coll = finish_row<K>(t, k, p.sensor_range, p.cd, dist);
        const size_t idx = base + i;
        float prev_h = 0.f;
        if (V == FLOCK_UW) prev_h = p.prev_h[idx];
        const float rew = agent_reward<V>(p, coll, x, y, h, prev_h, comx, comy, hmean);
        fx = reward_fx(rew);
        if (V == FLOCK_UW && !(prev_h == h)) p.prev_h[idx] = h;
        write_obs_t<K>(p, env, i, idx, dist, false);
        if (p.nn != nullptr) store_row_t<int, K>(p.nn + idx * k, t.idx, k);
        p.reward[idx] = rew;
        p.agent_done[idx] = coll ? 1 : 0;
    }
    // env-level reductions. The episode return is an order-free integer sum. env_done and the
    // episode-step counter are published by the LAST CTA of the env to finish (arrival counter), so
    // no CTA of this launch can observe a half-updated ep_len (it is the Philox epoch of uwd noise)
    // and no separate memset of env_done is needed.
    const bool warp_coll = __any_sync(0xffffffffu, coll);
    const unsigned lo = (unsigned)fx & 0xffffu, mid = (unsigned)(fx >> 16) & 0xffffu;
    const int hi = (int)(fx >> 32);
    const unsigned slo = __reduce_add_sync(0xffffffffu, lo);
    const unsigned smid = __reduce_add_sync(0xffffffffu, mid);
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    unsigned int* arrive = p.tile_scratch + env;
    unsigned int* collide = p.tile_scratch + p.E + env;
    if ((threadIdx.x & 31) == 0) {
        if (warp_coll) atomicAdd(collide, 1u);
        if (p.ep_return_fx != nullptr) {
            const long long s = ((long long)shi << 32) + ((long long)smid << 16) + (long long)slo;
            atomicAdd(reinterpret_cast<unsigned long long*>(p.ep_return_fx + env), (unsigned long long)s);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(arrive, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence();
            const unsigned c = atomicExch(collide, 0u);
            p.env_done[env] = c != 0u ? 1 : 0;
            p.ep_len[env] += 1;
            if (p.H > 1 && p.obs_head != nullptr) p.obs_head[env] = ring_prev_slot(p.obs_head[env], p.H);
            *arrive = 0u;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Many envs x large swarm (all three variants): thread-per-row kernel with EXACT spatial pruning.
//
// The env is staged in the spatially sorted row order (p.perm: Morton cells), so both the 32 rows of a
// warp and every block of 32 consecutive neighbour slots are compact in space. For each block the
// warp compares a conservative lower bound of the (min-image) distance between its rows' bounding
// box and the block's bounding box with the largest per-row threshold in the warp; blocks that
// cannot contain a k-nearest neighbour of any of the 32 rows are skipped, the others run the same
// 9-FP32-instruction pair loop as the all-pairs kernel. The thresholds come from the previous
// step's neighbour lists (an upper bound of the k-th distance, see hint_bound), so typically ~10 %
// of the blocks are evaluated. Results are identical to the all-pairs kernels bit for bit:
// pruning only removes candidates that provably lose, and candidates are merged by the full
// (d2, agent index) key because they no longer arrive in index order.
// -------------------------------------------------------------------------------------------------
size_t pruned_smem_bytes(int N, int rows) {
    const size_t PS = (size_t)((N + 31) / 32) * 32;
    return 3 * PS * sizeof(float) + (size_t)16 * rows * sizeof(unsigned short);   // record + kPrunedCandCap slots per row
}
size_t pruned_hint_bytes(int N, int E) {   // [E][PS] rows of 8 two-byte slots
    return (size_t)E * ((size_t)((N + 31) / 32) * 32) * 16;
}
size_t pruned_scratch_floats(int N, int E) {   // per env: x by slot | y by slot | boxes | agent ids (3 * PS floats)
    const size_t PS = (size_t)((N + 31) / 32) * 32;
    return (size_t)E * 3 * PS;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// conservative lower bound of the per-axis (min-image) distance between intervals [a0,a1], [b0,b1]
template <bool PER>
__device__ __forceinline__ float axis_gap(float a0, float a1, float b0, float b1, float B, float slack) {
    float g = fmaxf(b0 - a1, a0 - b1);                  // direct gap (negative when overlapping)
    if (PER) {
        const float span = fmaxf(a1 - b0, b1 - a0);     // largest |dx| over the two intervals
        g = fminf(g, B - span);                         // the wrapped image can be closer
    }
    return fmaxf(g - slack, 0.0f);
}

// -------------------------------------------------------------------------------------------------
// Pruned path for large swarms (v2: min-image metric; uw / uw_discrete: Euclidean). Two launches per step:
//  1. flock_integrate_kernel<V2, SORTED>: one thread per sorted SLOT integrates its agent ONCE (same code
//     as everywhere else), writes the new state at the agent's index and, per env, a staging record
//     [x by slot | y by slot | bounding boxes of every 8 consecutive slots] that is contiguous in memory.
//  2. flock_step_pruned_kernel: one thread per row (in slot order); the CTA stages the env's record
//     with ONE TMA bulk copy, each warp tests its 32-row bounding box against the 8-slot boxes (one
//     box per lane, ballot) and evaluates only boxes whose conservative distance lower bound does not
//     exceed the warp's largest current k-th-neighbour threshold. Exact: a skipped box cannot hold
//     a candidate that any row of the warp would accept; ties are resolved by (d2, agent index).
// -------------------------------------------------------------------------------------------------
constexpr int kBoxSlots = 8;

constexpr int kPrunedCandCap = 16;   // deferred candidates (slots) per row between two merges
// per env: x by slot | y by slot | one (x0, x1, y0, y1) box per 8 slots | agent id (u16) by slot
__host__ __device__ __forceinline__ size_t sorted_record_floats(int PS) { return (size_t)PS * 3; }

constexpr int kSumChunk = 2048;   // agents per shared-memory chunk of the per-env sequential sums

template <int V, bool SORTED>
__global__ void __launch_bounds__(256) flock_integrate_kernel(const __grid_constant__ Params p) {
    const int N = p.N, env = blockIdx.y;
    const int PS = ((N + 31) / 32) * 32;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t base = (size_t)env * N;
    const bool real = slot < N;
    const bool in_record = SORTED && slot < PS;   // PS is a multiple of 32: whole warps are in or out
    float x = kInf, y = 0.0f;                     // padding slots: never selected, never collide
    int agent = 0;
    if (real) {
        const int a = SORTED ? p.perm[base + slot] : slot;
        agent = a;
        const size_t i = base + (size_t)a;
        float h = p.h[i], vx, vy;
        x = p.x[i];
        y = p.y[i];
        const float x_old = x, y_old = y;
        float a0, a1 = 0.0f, nzu = 0.f, nzw = 0.f;
        if (V == FLOCK_UWD) {
            a0 = p.actions[i];
            if (p.noise != nullptr) {
                const float2 nz = reinterpret_cast<const float2*>(p.noise)[i];
                nzu = nz.x;
                nzw = nz.y;
            } else if (p.noise_std > 0.0f) {
                act_noise(p, p.env_offset + env, a, (uint32_t)p.ep_len[env], p.reset_epoch[env], nzu, nzw);
            }
        } else {
            const float2 act = reinterpret_cast<const float2*>(p.actions)[i];
            a0 = act.x;
            a1 = act.y;
        }
        integrate_agent<V>(p, a0, a1, nzu, nzw, x, y, h, vx, vy);
        p.xo[i] = x;                              // in place: xo/yo/ho alias x/y/h on the tiled path
        p.yo[i] = y;
        p.ho[i] = h;
        if (p.vx != nullptr) {
            p.vx[i] = vx;
            p.vy[i] = vy;
        }
        // wrap-around (check_boundary): the agent is now a world away from the rows it shares a warp with until the next
        // row-order refresh; the pruned kernel scans such rows on their own (see there)
        if (SORTED && p.far_rows != nullptr && (fabsf(x - x_old) > 0.5f * p.B || fabsf(y - y_old) > 0.5f * p.B))
            p.far_rows[(size_t)env * PS + slot] = 1;
    }
    if (in_record) {
        float* rec = p.sorted_xy + (size_t)env * sorted_record_floats(PS);
        rec[slot] = x;
        rec[PS + slot] = y;
        float x0 = real ? x : kFltMax, x1 = real ? x : -kFltMax, y0 = real ? y : kFltMax, y1 = real ? y : -kFltMax;
#pragma unroll
        for (int m = 1; m < kBoxSlots; m <<= 1) {
            x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, m));
            x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, m));
            y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, m));
            y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, m));
        }
        if ((slot & (kBoxSlots - 1)) == 0)
            reinterpret_cast<float4*>(rec + 2 * (size_t)PS)[slot / kBoxSlots] = make_float4(x0, x1, y0, y1);
        reinterpret_cast<unsigned short*>(rec + 2 * (size_t)PS + PS / 2)[slot] = (unsigned short)(real ? agent : 0xffff);
    }
    if (V != FLOCK_V2 && p.env_sums != nullptr) {
        // uw: torch.mean(positions, 0) (gym_flock_uw.py:193); uwd: sum(headings) / N (gym_flock_uw_discrete.py:256). The
        // canonical float32 sum is sequential in agent order, so it is formed ONCE per env -- by the env's last CTA to
        // finish integrating (arrival counter), two threads running the two serial chains over shared-memory chunks --
        // instead of by every row of the sensing kernel.
        __shared__ __align__(16) float cs[2][kSumChunk];
        __shared__ bool last;
        unsigned int* arrive = p.tile_scratch + env;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            last = atomicAdd(arrive, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (last) {
            __threadfence();
            float acc = 0.0f;
            for (int c0 = 0; c0 < N; c0 += kSumChunk) {
                const int n = min(kSumChunk, N - c0);
                for (int j = threadIdx.x; j < n; j += blockDim.x) {     // L2 loads: other CTAs wrote these values
                    if (V == FLOCK_UW) {
                        cs[0][j] = __ldcg(p.xo + base + c0 + j);
                        cs[1][j] = __ldcg(p.yo + base + c0 + j);
                    } else {
                        cs[0][j] = __ldcg(p.ho + base + c0 + j);
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0 || (V == FLOCK_UW && threadIdx.x == 32)) {
                    const float* v = cs[threadIdx.x >> 5];
                    const float4* v4 = reinterpret_cast<const float4*>(v);
                    const int n4 = n >> 2;
#pragma unroll 8
                    for (int j4 = 0; j4 < n4; ++j4) {      // 128-bit loads run ahead of the serial add chain
                        const float4 q = v4[j4];
                        acc = acc + q.x;
                        acc = acc + q.y;
                        acc = acc + q.z;
                        acc = acc + q.w;
                    }
                    for (int j = n4 << 2; j < n; ++j) acc = acc + v[j];
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                p.env_sums[2 * env] = acc;
                *arrive = 0u;                     // the sensing kernel of this step counts its CTAs from zero again
            }
            if (V == FLOCK_UW && threadIdx.x == 32) p.env_sums[2 * env + 1] = acc;
        }
    }
}

template <int V, bool SORTED>
static cudaError_t launch_integrate(const Params& p, cudaStream_t s) {
    const int PS = ((p.N + 31) / 32) * 32;
    flock_integrate_kernel<V, SORTED><<<dim3((PS + 255) / 256, p.E), 256, 0, s>>>(p);
    return cudaGetLastError();
}

// (min 4 CTAs of 256 threads per SM = at most 64 registers: the rare multi-row pass below must not cost the common path its occupancy)
template <int V, int K, bool PER>
__global__ void __launch_bounds__(kMaxTileThreads, 4) flock_step_pruned_kernel(const __grid_constant__ Params p) {
#ifdef FLOCK_TIMELINE
    const unsigned long long tl_entry = timeline_now();
    const long long tl_c0 = clock64();
#endif
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    constexpr unsigned kFull = 0xffffffffu;
    const int N = p.N, k = p.k;
    const int env = blockIdx.y, tile = blockIdx.x, rows = blockDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int PS = ((N + 31) / 32) * 32, nbox = PS / kBoxSlots;
    float* ss_x = smem;                 // new positions by sorted SLOT (padded to whole warps)
    float* ss_y = ss_x + PS;
    const float4* bb = reinterpret_cast<const float4*>(ss_y + PS);      // [nbox] (x0, x1, y0, y1)
    const unsigned short* sid = reinterpret_cast<const unsigned short*>(ss_y + PS + 4 * nbox);   // agent id by slot
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + sorted_record_floats(PS)) + threadIdx.x;
    const size_t base = (size_t)env * N;
    const int slot = tile * rows + threadIdx.x;
    const bool has_row = slot < N;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        const uint32_t bytes = (uint32_t)(sorted_record_floats(PS) * sizeof(float));
        mbar_expect_tx(&bar, bytes);
        tma_load_1d(smem, p.sorted_xy + (size_t)env * sorted_record_floats(PS), bytes, &bar);
    }
    const int i = has_row ? p.perm[base + slot] : N;    // this thread's agent
    // Threshold hints: the slots of last step's k neighbours, kept by this kernel in slot order (one
    // coalesced 16-byte row per agent). Any k distinct other slots bound the k-th distance, so stale
    // rows are harmless; rows invalidated by a row-order refresh or a reset (0xffff) fall back to the
    // neighbour list in agent ids (p.nn) translated through p.inv.
    int hslot[K];
    bool hint_ok = has_row;
    {
        uint4 hv = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (has_row) hv = reinterpret_cast<const uint4*>(p.hint_slots)[(size_t)env * PS + slot];
        const unsigned w[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int s = 0; s < K; ++s) {
            hslot[s] = (int)((w[(s >> 1) & 3] >> ((s & 1) * 16)) & 0xffffu);
            if (s < k) hint_ok = hint_ok && hslot[s] < N && hslot[s] != slot;
        }
    }
    if (has_row && !hint_ok && p.nn != nullptr) {
        hint_ok = true;
        const int* row = p.nn + (base + i) * k;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            if (s < k) {
                const int j = row[s];
                const bool in = j >= 0 && j < N && j != i;
                hint_ok = hint_ok && in;
                hslot[s] = in ? p.inv[base + j] : 0;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < K; ++s)
        if (s >= k || !hint_ok) hslot[s] = 0;
#pragma unroll
    for (int s = 1; s < K; ++s)
#pragma unroll
        for (int u = 0; u < s; ++u)
            if (s < k) hint_ok = hint_ok && (hslot[u] != hslot[s]);
    __syncthreads();
    mbar_wait(&bar, 0);

    const float x = has_row ? ss_x[slot] : 0.f, y = has_row ? ss_y[slot] : 0.f;
    // uw / uwd: the epilogue's read-modify-write operands (heading, previous heading, the three older rows of the uw
    // observation window) are fetched now, so their DRAM round trip runs under the scan instead of after it
    float h_row = 0.f, prev_h_row = 0.f;
    float win[(V == FLOCK_UW) ? 3 * K : 1];
    const bool pre_win = V == FLOCK_UW && p.H == 4 && p.obs_head == nullptr;
    if (V != FLOCK_V2 && has_row) {
        h_row = p.h[base + i];
        if (V == FLOCK_UW) {
            prev_h_row = p.prev_h[base + i];
            if (pre_win) {
                const float* o = p.obs + (base + i) * (size_t)(4 * k);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int s = 0; s < K; ++s) win[r * K + s] = s < k ? o[r * k + s] : 0.0f;
            }
        }
    }
    // per-row threshold: the largest CURRENT distance to last step's neighbours bounds the k-th best
    float thr = -1.0f;                                     // lanes without a row accept nothing
    if (has_row) {
        float bound = 0.0f;
#pragma unroll
        for (int s = 0; s < K; ++s)
            if (s < k) bound = fmaxf(bound, pair_d2<PER>(x, y, ss_x[hslot[s]], ss_y[hslot[s]], p.B));
        thr = hint_ok ? bound : kFltMax;
    }
    // Far rows: an agent that wrapped around since the last row-order refresh sits a world away (in coordinates) from
    // the other rows of its warp. Left in, it stretches the warp's bounding box over the whole world (every box passes the
    // gap test along that axis) and, in a Euclidean world, brings a threshold of ~B^2 on the step of the wrap -- one such
    // row made its warp up to 13x slower, and the kernel as slow as its slowest warp (uwd 64 x 2048: 82 vs 42 us).
    // They are taken out of the shared pass and scanned one at a time by the whole warp below.
    bool far = false;
    if (p.far_rows != nullptr && has_row) far = p.far_rows[(size_t)env * PS + slot] != 0;
    {
        // ... and rows whose THRESHOLD is an outlier in their warp (more than 8x the warp's median d2): the old
        // neighbours of an agent that has just wrapped (their hint bound jumps to ~B^2 in a Euclidean world), rows without a
        // valid hint, and genuinely isolated agents (uw draws its swarm in a quarter of the world: whoever wraps sits alone on
        // the far side, 1000 away from everyone; after a refresh such agents collect in the first / last slots). In the
        // shared pass one such row makes all 32 lanes evaluate every box (measured: uw 64 x 2048, 1-5 isolated rows in the
        // last tile of each env: that CTA 190 k cycles against a mean of 33 k, the kernel 95 instead of 50 us).
        // The yardstick is the warp's MEDIAN threshold (radix select on sign | exponent | two mantissa bits with ballots,
        // ~70 instructions per warp): the smallest threshold of a warp is no yardstick -- k-th-neighbour distances of
        // uniformly scattered agents spread over more than 10x in d2 within 32 rows for k = 3 or 4 -- while 8x the median
        // is exceeded by a normal row with probability ~1e-7.
        const bool cand_row = has_row && !far;
        const unsigned tkey = __float_as_uint(fmaxf(thr, 0.0f)) >> 21;       // 11 bits, monotone in thr
        const unsigned act = __ballot_sync(kFull, cand_row);
        int need = (__popc(act) + 1) >> 1;
        unsigned prefix = 0u;
#pragma unroll
        for (int b = 10; b >= 0; --b) {
            const unsigned zero = __ballot_sync(kFull, cand_row && (tkey >> (b + 1)) == (prefix >> (b + 1)) && ((tkey >> b) & 1u) == 0u);
            const int c0 = __popc(zero);
            if (need > c0) {
                need -= c0;
                prefix |= 1u << b;
            }
        }
        far = far || (cand_row && act != 0u && tkey > prefix + 12u);          // 12 quarter-octaves = 8x in d2
        // ... and rows that lie far from the warp's patch without having wrapped: agents that diffused out of the swarm's
        // bulk are sorted into distant cells of the row order and end up among unrelated rows (measured, uw: two such rows in
        // the last warp of an env stretched its bounding box over the whole swarm -> every box opened, 119 k cycles against
        // a mean of 39 k). Yardstick: distance from the warp's mean position against 20 median neighbour radii (a patch of
        // 32 rows is 4-6 such radii across).
        float sx = cand_row ? x : 0.0f, sy = cand_row ? y : 0.0f;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            sx += __shfl_xor_sync(kFull, sx, m);
            sy += __shfl_xor_sync(kFull, sy, m);
        }
        const float inv_cnt = 1.0f / (float)max(__popc(act), 1);
        const float ddx = x - sx * inv_cnt, ddy = y - sy * inv_cnt;
        const float med_thr = fmaxf(__uint_as_float(prefix << 21), 1.0e-6f * p.B * p.B);
        far = far || (cand_row && (ddx * ddx + ddy * ddy) > 400.0f * med_thr);
    }
    unsigned far_mask = __ballot_sync(kFull, far);
    if (__popc(far_mask) > 8) {        // a warp of strays (dense wrap-around, stale order after a masked reset): shared pass
        far_mask = 0u;
        far = false;
    }
    const float thr_hint = thr;
    const bool main_row = has_row && !far;
    if (far) thr = -1.0f;
    const float wx0 = warp_min(main_row ? x : kFltMax), wx1 = warp_max(main_row ? x : -kFltMax);
    const float wy0 = warp_min(main_row ? y : kFltMax), wy1 = warp_max(main_row ? y : -kFltMax);
    // thresholds are non-negative floats: their bit patterns order like unsigned integers
    float thr_max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fmaxf(thr, 0.0f))));
    const float slack = p.B * 1.0e-6f + 1.0e-30f;

    TopK<K> t;
    t.init();
    int cnt = 0;
    const int cstride = rows;
    // Candidates are appended as SLOTS (2 bytes each); the merge recomputes their distance (same
    // inputs, same code: same bits) and inserts them by the full (d2, agent id) key, because they
    // do not arrive in index order. The row's own slot is dropped here, not in the pair loop.
    auto merge = [&]() {
        const int maxc = __reduce_max_sync(kFull, cnt);
#pragma unroll 1
        for (int c = 0; c < maxc; ++c) {
            if (c < cnt) {
                const int sj = cand[c * cstride];
                if (sj != slot) t.insert_lex(pair_d2<PER>(x, y, ss_x[sj], ss_y[sj], p.B), ((int)sid[sj] << 16) | sj);
            }
        }
        cnt = 0;
        thr = fminf(thr, t.worst());
        thr_max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(fmaxf(thr, 0.0f))));
    };
    const int ngroups = (nbox + 31) / 32;
    const int own_g = slot / (kBoxSlots * 32);            // group of 32 boxes holding this warp's first row
    unsigned long long evaluated = 0;
    if (__any_sync(kFull, main_row)) {
        for (int o = 0; o < ngroups; ++o) {
            int g = __shfl_sync(kFull, own_g, 0) + o;
            if (g >= ngroups) g -= ngroups;
            const int bl = g * 32 + lane;                 // this lane's box
            bool take = false;
            if (bl < nbox) {
                const float4 q = bb[bl];
                const float gx = axis_gap<PER>(wx0, wx1, q.x, q.y, p.B, slack);
                const float gy = axis_gap<PER>(wy0, wy1, q.z, q.w, p.B, slack);
                take = (gx * gx + gy * gy) * 0.9999f <= thr_max;   // some row of the warp may gain from it
            }
            unsigned mask = __ballot_sync(kFull, take);
            while (mask != 0u) {
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1u;
                const int s0 = (g * 32 + bit) * kBoxSlots;
                evaluated += 1;
                uint32_t ax = smem_u32(ss_x + s0), ay = smem_u32(ss_y + s0);
#pragma unroll
                for (int j4 = 0; j4 < kBoxSlots / 4; ++j4, ax += 16u, ay += 16u) {
                    const float4 X = lds128(ax);
                    const float4 Y = lds128(ay);
                    const float d0 = pair_d2<PER>(x, y, X.x, Y.x, p.B);
                    const float d1 = pair_d2<PER>(x, y, X.y, Y.y, p.B);
                    const float d2 = pair_d2<PER>(x, y, X.z, Y.z, p.B);
                    const float d3 = pair_d2<PER>(x, y, X.w, Y.w, p.B);
                    const float m = fminf(fminf(d0, d1), fminf(d2, d3));
                    if (__any_sync(kFull, m <= thr)) {
                        const unsigned sj = (unsigned)(s0 + (j4 << 2));
                        if (d0 <= thr) { cand[cnt * cstride] = (unsigned short)sj; ++cnt; }
                        if (d1 <= thr) { cand[cnt * cstride] = (unsigned short)(sj + 1u); ++cnt; }
                        if (d2 <= thr) { cand[cnt * cstride] = (unsigned short)(sj + 2u); ++cnt; }
                        if (d3 <= thr) { cand[cnt * cstride] = (unsigned short)(sj + 3u); ++cnt; }
                        if (__any_sync(kFull, cnt > kPrunedCandCap - 4)) merge();
                    }
                }
            }
        }
        merge();
    }
    // far rows, one at a time, all 32 lanes on the same row: lanes split the 8-slot boxes, a box is opened when its gap to
    // the row's POINT does not exceed the row's threshold, and the 32 private lists are merged by the warp-shuffle bitonic
    // top-k on (d2, agent id) keys. Threshold: the smaller of the hint bound and a bound from the boxes themselves -- a
    // complete box other than the row's own holds 8 >= k other agents, all within the largest distance from the point to
    // the box's corners -- so the step of the wrap (hint bound ~B^2) is pruned as well. Exact like the shared pass.
#ifdef FLOCK_TIMELINE
    const long long tl_c1 = clock64();
    const int tl_nfar = __popc(far_mask);
#endif
    // Up to four such rows are scanned at a time, one per group of 32 / 16 / 8 lanes (a row costs ~1000 warp-instructions
    // of mostly fixed work -- box tests, the butterfly merge -- so a warp with six of them spent 40 k cycles here).
    while (far_mask != 0u) {
        const int left = __popc(far_mask);
        // lanes per row (k > 4: one row at a time -- the 8-entry lists make the grouped form cost registers and time, and
        // measured on cfg5, which has few such rows, it lost 2 us)
        const int gs = K > 4 ? 32 : (left >= 4 ? 8 : (left >= 2 ? 16 : 32));
        const int ngroups = 32 / gs, grp = lane / gs, li = lane - grp * gs;
        int r0 = 0, r1 = 0, r2 = 0, r3 = 0;                              // scalars, not an array: no local memory
        r0 = __ffs(far_mask) - 1;
        far_mask &= far_mask - 1u;
        if (ngroups >= 2) {
            r1 = __ffs(far_mask) - 1;
            far_mask &= far_mask - 1u;
        }
        if (ngroups >= 4) {
            r2 = __ffs(far_mask) - 1;
            far_mask &= far_mask - 1u;
            r3 = __ffs(far_mask) - 1;
            far_mask &= far_mask - 1u;
        }
        const int r = grp == 0 ? r0 : (grp == 1 ? r1 : (grp == 2 ? r2 : r3));   // the row this lane works for
        const float xq = __shfl_sync(kFull, x, r), yq = __shfl_sync(kFull, y, r);
        const int sq = __shfl_sync(kFull, slot, r);
        float tb = __shfl_sync(kFull, thr_hint, r);
        const int nfull = N / kBoxSlots;
        for (int bl = li; bl < nfull; bl += gs) {
            if (bl == sq / kBoxSlots) continue;
            const float4 q = bb[bl];
            float ddx = fmaxf(fabsf(xq - q.x), fabsf(xq - q.y)), ddy = fmaxf(fabsf(yq - q.z), fabsf(yq - q.w));
            if (PER) {      // a min-image distance never exceeds the direct one, nor B/2
                ddx = fminf(ddx, 0.5f * p.B);
                ddy = fminf(ddy, 0.5f * p.B);
            }
            tb = fminf(tb, fmaf(ddy, ddy, ddx * ddx) * 1.0001f + 1.0e-30f);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {                              // minimum over the row's lane group
            const float o = __shfl_xor_sync(kFull, tb, m);
            if (m < gs) tb = fminf(tb, o);
        }
        TopK<K> tq;
        tq.init();
        for (int bl = li; bl < nbox; bl += gs) {
            const float4 q = bb[bl];
            const float gx = axis_gap<PER>(xq, xq, q.x, q.y, p.B, slack);
            const float gy = axis_gap<PER>(yq, yq, q.z, q.w, p.B, slack);
            if ((gx * gx + gy * gy) * 0.9999f <= tb) {
                evaluated += 1;
#pragma unroll
                for (int u = 0; u < kBoxSlots; ++u) {
                    const int sj = bl * kBoxSlots + u;
                    const float d = pair_d2<PER>(xq, yq, ss_x[sj], ss_y[sj], p.B);
                    if (sj != sq && d <= tb) tq.insert_lex(d, ((int)sid[sj] << 16) | sj);
                }
            }
        }
        unsigned long long key[K];
#pragma unroll
        for (int s = 0; s < K; ++s) key[s] = ((unsigned long long)__float_as_uint(tq.d[s]) << 32) | (unsigned)tq.idx[s];
        warp_bitonic_topk<K>(key, gs);   // every lane of a group ends with the group's merged list
        if (K > 4) {
            if (lane == r0) {
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    t.d[s] = __uint_as_float((unsigned)(key[s] >> 32));
                    t.idx[s] = (int)(unsigned)key[s];
                }
            }
            continue;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {    // hand each list to the lane that owns the row (it may sit in another group)
            if (g < ngroups) {
                const int owner = g == 0 ? r0 : (g == 1 ? r1 : (g == 2 ? r2 : r3));
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const unsigned long long kk = shfl_u64(key[s], g * gs);
                    if (lane == owner) {
                        t.d[s] = __uint_as_float((unsigned)(kk >> 32));
                        t.idx[s] = (int)(unsigned)kk;
                    }
                }
            }
        }
    }

#ifdef FLOCK_TIMELINE
    if (lane == 0) {      // slowest warp of the CTA: shared pass, one-at-a-time pass (SM clocks); rows taken out, boxes opened
        const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < (size_t)kTimelineCtas) {
            unsigned long long* d = p.timeline + cta * kTimelineSlots;
            atomicMax(&d[4], (unsigned long long)(tl_c1 - tl_c0));
            atomicMax(&d[5], (unsigned long long)(clock64() - tl_c1));
            atomicAdd(&d[6], (unsigned long long)tl_nfar);
            atomicAdd(&d[7], evaluated);
        }
    }
#endif
    long long fx = 0;
    bool coll = false;
    if (has_row) {
        float dist[K];
        coll = finish_row<K, TopK<K>>(t, k, p.sensor_range, p.cd, dist);
        const size_t idx = base + i;
        float rew;
        if (V == FLOCK_V2) {
            rew = reward_from_flags<FLOCK_V2>(coll, false, false);
        } else {
            // uw: centre of mass + heading change (gym_flock_uw.py:186-221); uwd: alignment with the mean heading
            // (gym_flock_uw_discrete.py:234-276); the per-env sums come from the integrate pre-pass
            float comx = 0.f, comy = 0.f, hmean = 0.f;
            const float h = h_row, prev_h = prev_h_row;
            if (V == FLOCK_UW) {
                comx = mean_of_sum(p, p.env_sums[2 * env]);
                comy = mean_of_sum(p, p.env_sums[2 * env + 1]);
            } else {
                hmean = mean_of_sum(p, p.env_sums[2 * env]);
            }
            rew = agent_reward<V>(p, coll, x, y, h, prev_h, comx, comy, hmean);
            if (V == FLOCK_UW && !(prev_h == h)) p.prev_h[idx] = h;
        }
        fx = reward_fx(rew);
        if (V == FLOCK_UW && pre_win) {     // window shifted by one row (gym_flock_uw.py:120-123), old rows from registers
            float* o = p.obs + idx * (size_t)(4 * k);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < K; ++s)
                    if (s < k) o[(r + 1) * k + s] = win[r * K + s];
#pragma unroll
            for (int s = 0; s < K; ++s)
                if (s < k) o[s] = dist[s];
        } else {
            write_obs_t<K>(p, env, i, idx, dist, false);
        }
        // list entries are (agent id << 16 | slot): ids go to the neighbour list, slots to next step's hints
        int ids[K];
        unsigned hw[4] = {~0u, ~0u, ~0u, ~0u};
#pragma unroll
        for (int s = 0; s < K; ++s) {
            ids[s] = t.idx[s] >> 16;
            const unsigned sl = (unsigned)t.idx[s] & 0xffffu;
            hw[(s >> 1) & 3] = (s & 1) ? ((hw[(s >> 1) & 3] & 0xffffu) | (sl << 16)) : ((hw[(s >> 1) & 3] & 0xffff0000u) | sl);
        }
        if (p.nn != nullptr) store_row_t<int, K>(p.nn + idx * k, ids, k);
        reinterpret_cast<uint4*>(p.hint_slots)[(size_t)env * PS + slot] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        p.reward[idx] = rew;
        p.agent_done[idx] = coll ? 1 : 0;
    }
    const bool warp_coll = __any_sync(kFull, coll);
    const unsigned lo = (unsigned)fx & 0xffffu, mid = (unsigned)(fx >> 16) & 0xffffu;
    const int hi = (int)(fx >> 32);
    const unsigned slo = __reduce_add_sync(kFull, lo);
    const unsigned smid = __reduce_add_sync(kFull, mid);
    const int shi = __reduce_add_sync(kFull, hi);
    unsigned int* arrive = p.tile_scratch + env;
    unsigned int* collide = p.tile_scratch + p.E + env;
    if (lane == 0) {
        if (warp_coll) atomicAdd(collide, 1u);
        if (p.ep_return_fx != nullptr) {
            const long long sum = ((long long)shi << 32) + ((long long)smid << 16) + (long long)slo;
            atomicAdd(reinterpret_cast<unsigned long long*>(p.ep_return_fx + env), (unsigned long long)sum);
        }
        if (p.pair_counter != nullptr) atomicAdd(p.pair_counter, evaluated * (unsigned long long)(kBoxSlots * 32));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(arrive, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence();
            const unsigned c = atomicExch(collide, 0u);
            p.env_done[env] = c != 0u ? 1 : 0;
            p.ep_len[env] += 1;
            if (p.H > 1 && p.obs_head != nullptr) p.obs_head[env] = ring_prev_slot(p.obs_head[env], p.H);
            *arrive = 0u;
        }
    }
    (void)wid;
#ifdef FLOCK_TIMELINE
    if (threadIdx.x == 0) {
        const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < (size_t)kTimelineCtas) {
            unsigned long long* d = p.timeline + cta * kTimelineSlots;
            d[0] = tl_entry;
            d[1] = tl_entry;
            d[2] = timeline_now();
            d[3] = timeline_smid();
        }
    }
#endif
}

static int pruned_rows_override() {   // FLOCK_PRUNED_ROWS=64|128|256: rows per CTA of the pruned kernel (tuning knob)
    static const int v = [] {
        const char* e = getenv("FLOCK_PRUNED_ROWS");
        const int r = e != nullptr ? atoi(e) : 0;
        return (r >= 32 && r <= kMaxTileThreads && r % 32 == 0) ? r : 0;
    }();
    return v;
}

template <int V, int K, bool PER>
static cudaError_t launch_pruned(const Params& p, int sm_count, cudaStream_t s) {
    {   // launch 1 of 2: integrate every agent once, in slot order
        cudaError_t e = launch_integrate<V, true>(p, s);
        if (e != cudaSuccess) return e;
    }
    const int rows = pruned_rows_override() ? pruned_rows_override() : choose_rows(p.N, p.E, sm_count);
    const dim3 grid((p.N + rows - 1) / rows, p.E);
    // (launching this kernel as a programmatic dependent of the pre-pass was measured: +2 us at 64 envs when launched
    // eagerly, nothing under graph replay, and a LOSS of 5-14 us per step at 16-48 envs -- plain stream order it is)
    flock_step_pruned_kernel<V, K, PER><<<grid, rows, pruned_smem_bytes(p.N, rows), s>>>(p);
    return cudaGetLastError();
}

// Row order for the thread-per-row kernels: agents sorted by the Hilbert index of their cell on a
// G x G grid with about two agents per cell (counting sort, one CTA per env), so any run of
// consecutive slots is a compact, connected patch of the world. Positions drift slowly (<= 0.25 per
// step), so the order is refreshed only every few steps; the order within a cell is whatever the
// atomics give -- only speed depends on the order, never results.
__device__ __forceinline__ unsigned hilbert_index(unsigned n, unsigned x, unsigned y) {   // n = grid side, power of two
    unsigned d = 0;
    for (unsigned s = n >> 1; s > 0; s >>= 1) {
        const unsigned rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0u) {
            if (rx == 1u) {
                x = n - 1u - x;
                y = n - 1u - y;
            }
            const unsigned tmp = x;
            x = y;
            y = tmp;
        }
    }
    return d;
}

static int order_grid_side(int N) {   // power of two, cells of ~2 agents, 4 <= G <= 64
    int g = 4;
    while (g < 64 && g * g * 2 < N) g <<= 1;
    return g;
}

__global__ void __launch_bounds__(256) flock_perm_refresh_kernel(const __grid_constant__ Params p, int* perm, int* inv, int G) {
    __shared__ int part[256];
    extern __shared__ __align__(16) float smem[];
    const int env = blockIdx.x, N = p.N, cells = G * G;
    int* count = reinterpret_cast<int*>(smem);                            // [cells] counts, then start offsets
    unsigned short* key = reinterpret_cast<unsigned short*>(count + cells);   // [N]
    unsigned short* pos = key + N;                                         // [N]
    const size_t base = (size_t)env * N;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) count[c] = 0;
    __syncthreads();
    const float scale = (float)G / p.B;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        const float fx = p.x[base + a] * scale, fy = p.y[base + a] * scale;
        const unsigned cx = (unsigned)min(G - 1, max(0, (int)fx)), cy = (unsigned)min(G - 1, max(0, (int)fy));
        const unsigned c = hilbert_index((unsigned)G, cx, cy);
        key[a] = (unsigned short)c;
        pos[a] = (unsigned short)atomicAdd(&count[c], 1);
    }
    __syncthreads();
    // exclusive prefix sum over the cells: per-thread chunk sums, serial scan of the 256 partials
    const int chunk = (cells + 255) / 256;
    const int c0 = threadIdx.x * chunk, c1 = min(cells, c0 + chunk);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += count[c];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int t = 0; t < 256; ++t) {
            const int v = part[t];
            part[t] = acc;
            acc += v;
        }
    }
    __syncthreads();
    int acc = part[threadIdx.x];
    for (int c = c0; c < c1; ++c) {
        const int v = count[c];
        count[c] = acc;
        acc += v;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        const int s = count[key[a]] + pos[a];
        perm[base + s] = a;
        inv[base + a] = s;
    }
    if (p.hint_slots != nullptr) {   // slots changed meaning: next step takes its hints from the neighbour list
        const int PS = ((N + 31) / 32) * 32;
        uint4* hs = reinterpret_cast<uint4*>(p.hint_slots) + (size_t)env * PS;
        for (int t = threadIdx.x; t < PS; t += blockDim.x) hs[t] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    if (p.far_rows != nullptr) {     // every row is back among its spatial neighbours
        const int PS = ((N + 31) / 32) * 32;
        for (int t = threadIdx.x; t < PS; t += blockDim.x) p.far_rows[(size_t)env * PS + t] = 0;
    }
}

cudaError_t launch_perm_refresh(const Params& p, int* perm, int* inv, cudaStream_t s) {
    const int G = order_grid_side(p.N);
    const size_t bytes = (size_t)G * G * sizeof(int) + (size_t)p.N * 2 * sizeof(unsigned short);
    flock_perm_refresh_kernel<<<p.E, 256, bytes, s>>>(p, perm, inv, G);
    return cudaGetLastError();
}

__global__ void flock_perm_identity_kernel(int* perm, int* inv, int N, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        perm[i] = (int)(i % (size_t)N);
        inv[i] = (int)(i % (size_t)N);
    }
}

cudaError_t launch_perm_identity(int* perm, int* inv, int N, int E, cudaStream_t s) {
    const size_t total = (size_t)N * E;
    int grid = (int)((total + 255) / 256);
    if (grid > 1024) grid = 1024;
    flock_perm_identity_kernel<<<grid, 256, 0, s>>>(perm, inv, N, total);
    return cudaGetLastError();
}

bool tiled_uses_row_order(const Params& p, int sm_count, int tiled_mode) { return !use_rowwarp(p, sm_count, tiled_mode); }

// -------------------------------------------------------------------------------------------------
// Few envs x large swarm (fewer row tiles than SMs): one WARP per row instead of one thread per row.
// Lane l scans neighbours j = l, l+32, ... into a private sorted k-list, then the 32 partial lists
// are merged by a warp-shuffle bitonic top-k: five xor-butterfly rounds, each taking
// min(mine[i], partner[K-1-i]) (a bitonic sequence holding the K smallest of both lists) and
// re-sorting it with a log2(K)-stage bitonic network, on 64-bit keys (d2 bits << 32 | j) whose
// unsigned order IS the canonical (d2, j) order, so ties resolve to the lower index across lanes.
// grid = (row tiles, E), blockDim = 256 (8 warps), each warp handles `rows_per_warp` rows.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}

// `width` (32, 16 or 8, warp-uniform): merge within aligned groups of that many lanes only
template <int K>
__device__ __forceinline__ void warp_bitonic_topk(unsigned long long (&key)[K], int width) {
    static_assert((K & (K - 1)) == 0, "K must be a power of two");
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        if (m >= width) break;
        unsigned long long other[K];
#pragma unroll
        for (int s = 0; s < K; ++s) other[s] = shfl_xor_u64(key[s], m);
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const unsigned long long o = other[K - 1 - s];
            key[s] = o < key[s] ? o : key[s];
        }
#pragma unroll
        for (int stride = K / 2; stride >= 1; stride >>= 1) {
#pragma unroll
            for (int s = 0; s < K; ++s) {
                if ((s & stride) == 0) {
                    const unsigned long long a = key[s], b = key[s + stride];
                    key[s] = a < b ? a : b;
                    key[s + stride] = a < b ? b : a;
                }
            }
        }
    }
}

template <int V, int K, bool PER>
__global__ void __launch_bounds__(kMaxTileThreads) flock_step_rowwarp_kernel(const __grid_constant__ Params p,
                                                                           int rows_per_warp) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int N = p.N, k = p.k;
    const int env = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int rows_per_cta = warps * rows_per_warp;
    const int row0 = tile * rows_per_cta;
    const TileSmem sm = carve(smem, N, true);
    const size_t base = (size_t)env * N;
    // the state was integrated in place by the pre-pass: stage the NEW positions and headings
    stage_xy(p, sm, &bar, env);
    for (int a = threadIdx.x; a < N; a += blockDim.x) sm.sh[a] = p.h[base + a];
    __syncthreads();

    // per-env sums of the new state, formed once by the integrate pre-pass (see flock_integrate_kernel)
    float comx = 0.f, comy = 0.f, hmean = 0.f;
    if (V == FLOCK_UW) {
        comx = mean_of_sum(p, p.env_sums[2 * env]);
        comy = mean_of_sum(p, p.env_sums[2 * env + 1]);
    }
    if (V == FLOCK_UWD) hmean = mean_of_sum(p, p.env_sums[2 * env]);

    long long fx_acc = 0;
    bool coll_any = false;
    for (int r = 0; r < rows_per_warp; ++r) {
        const int i = row0 + wid * rows_per_warp + r;     // warp-uniform
        if (i >= N) break;
        const size_t idx = base + i;
        const float xi = sm.sx[i], yi = sm.sy[i];
        // threshold from last step's neighbour list: lane s < k checks hint s
        float thr = kFltMax;
        if (p.nn != nullptr) {
            const int hj = lane < k ? p.nn[idx * k + lane] : -1;
            bool ok = lane >= k || (hj >= 0 && hj < N && hj != i);
#pragma unroll
            for (int u = 1; u < FLOCK_MAX_K; ++u) {     // distinctness among the first k lanes
                const int other = __shfl_sync(0xffffffffu, hj, (lane + u) & (FLOCK_MAX_K - 1));
                if (lane < k && ((lane + u) & (FLOCK_MAX_K - 1)) < k && other == hj) ok = false;
            }
            const int jj = (lane < k && ok) ? hj : 0;
            float hd = lane < k ? pair_d2<PER>(xi, yi, sm.sx[jj], sm.sy[jj], p.B) : 0.0f;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) hd = fmaxf(hd, __shfl_xor_sync(0xffffffffu, hd, m));
            thr = __all_sync(0xffffffffu, ok) ? hd : kFltMax;
        }
        TopK<K> t;
        t.init();
        for (int j = lane; j < N; j += 32) {
            const float d = pair_d2<PER>(xi, yi, sm.sx[j], sm.sy[j], p.B);
            if (j != i && d <= thr && d < t.worst()) t.insert(d, j);
        }
        unsigned long long key[K];
#pragma unroll
        for (int s = 0; s < K; ++s)
            key[s] = ((unsigned long long)__float_as_uint(t.d[s]) << 32) | (unsigned)t.idx[s];
        warp_bitonic_topk<K>(key);
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < K; ++s) {
                t.d[s] = __uint_as_float((unsigned)(key[s] >> 32));
                t.idx[s] = (int)(unsigned)key[s];
            }
            float dist[K];
            const bool coll = finish_row<K>(t, k, p.sensor_range, p.cd, dist);
            const float h = sm.sh[i];
            float prev_h = 0.f;
            if (V == FLOCK_UW) prev_h = p.prev_h[idx];
            const float rew = agent_reward<V>(p, coll, xi, yi, h, prev_h, comx, comy, hmean);
            fx_acc += reward_fx(rew);
            coll_any = coll_any || coll;
            if (V == FLOCK_UW && !(prev_h == h)) p.prev_h[idx] = h;
            write_obs_t<K>(p, env, i, idx, dist, false);
            if (p.nn != nullptr) store_row_t<int, K>(p.nn + idx * k, t.idx, k);
            p.reward[idx] = rew;
            p.agent_done[idx] = coll ? 1 : 0;
        }
    }
    unsigned int* arrive = p.tile_scratch + env;
    unsigned int* collide = p.tile_scratch + p.E + env;
    if (lane == 0) {
        if (coll_any) atomicAdd(collide, 1u);
        if (p.ep_return_fx != nullptr && fx_acc != 0)
            atomicAdd(reinterpret_cast<unsigned long long*>(p.ep_return_fx + env), (unsigned long long)fx_acc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(arrive, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence();
            const unsigned c = atomicExch(collide, 0u);
            p.env_done[env] = c != 0u ? 1 : 0;
            p.ep_len[env] += 1;
            if (p.H > 1 && p.obs_head != nullptr) p.obs_head[env] = ring_prev_slot(p.obs_head[env], p.H);
            *arrive = 0u;
        }
    }
}

// reset, grid = E, one CTA per env, bounded rejection loop (gym_flock_v2.py:85-108). The CTA is as wide as the
// hardware allows (1024 threads, one row each per pass): a restart is a latency problem -- typically one env of a
// batch restarts while the rest of the GPU waits for the stream -- so the env's rows are spread over 32 warps.
constexpr int kResetThreads = 1024;
template <int K>
__global__ void __launch_bounds__(kResetThreads) flock_reset_tiled_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) float smem[];
    const int N = p.N, k = p.k;
    const int env = blockIdx.x;
    if (p.env_mask != nullptr && p.env_mask[env] == 0) return;
    const TileSmem sm = carve(smem, N, true);
    const size_t base = (size_t)env * N;
    const size_t EN = (size_t)p.E * N;
    const uint32_t epoch = p.reset_epoch[env];
    const int max_att = p.init_state != nullptr ? 1 : p.max_attempts;
    int attempts = (int)p.step_offset;   // > 0: continues where the multi-CTA attempt launches stopped (env_mask = their need flags)
    int env_coll = 1;
    const bool keep = (p.reset_flags & FLOCK_RESET_KEEP_OUTPUTS) != 0;
    if (threadIdx.x < padded_agents(N) - N) {
        sm.sx[N + threadIdx.x] = kInf;
        sm.sy[N + threadIdx.x] = 0.0f;
    }
    while (env_coll && attempts < max_att) {
        __syncthreads();
        for (int a = threadIdx.x; a < N; a += blockDim.x) {
            float x, y, h;
            if (p.init_state != nullptr) {
                x = p.init_state[base + a];
                y = p.init_state[EN + base + a];
                h = p.init_state[2 * EN + base + a];
            } else {
                const uint4 r = philox4x32_10((uint32_t)(p.env_offset + env), (uint32_t)a, epoch + (uint32_t)attempts,
                                              kTagReset, p.seed_lo, p.seed_hi);
                const float span = p.range_lo - p.reset_hi;
                const float tx = span * u24(r.x);
                x = tx + p.reset_hi;
                const float ty = span * u24(r.y);
                y = ty + p.reset_hi;
                const float th = (0.0f - p.heading_hi) * u24(r.z);
                h = th + p.heading_hi;
            }
            sm.sx[a] = wrap_coord(x, p.B, p.fill_hi, p.fill_lo);
            sm.sy[a] = wrap_coord(y, p.B, p.fill_hi, p.fill_lo);
            sm.sh[a] = h;
        }
        attempts += 1;
        __syncthreads();
        int coll_any = 0;
        for (int i0 = 0; i0 < N; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            const bool has_row = i < N;
            TopK<K> t;
            const float x = sm.sx[has_row ? i : 0], y = sm.sy[has_row ? i : 0];
            knn_tiled<K, false>(sm.sx, sm.sy, sm.cand + threadIdx.x, i, N, x, y, p.B, has_row ? kFltMax : -1.0f, t);
            if (!has_row) continue;
            float dist[K];
            const bool coll = finish_row<K>(t, k, p.sensor_range, p.reset_cd, dist);
            coll_any |= coll ? 1 : 0;
            const size_t idx = base + i;
            p.xo[idx] = x;
            p.yo[idx] = y;
            p.ho[idx] = sm.sh[i];
            p.prev_h[idx] = 0.0f;
            if (p.vx != nullptr) {
                p.vx[idx] = 0.0f;
                p.vy[idx] = 0.0f;
            }
            write_obs_t<K>(p, env, i, idx, dist, true);
            if (p.nn != nullptr) store_row_t<int, K>(p.nn + idx * k, t.idx, k);
            if (!keep) {
                p.reward[idx] = 0.0f;
                p.agent_done[idx] = coll ? 1 : 0;
            }
        }
        env_coll = __syncthreads_or(coll_any);
    }
    if (p.hint_slots != nullptr) {   // the pruned kernel's slot hints describe the old episode: drop them
        const int PS = ((N + 31) / 32) * 32;
        uint4* hs = reinterpret_cast<uint4*>(p.hint_slots) + (size_t)env * PS;
        for (int t = threadIdx.x; t < PS; t += blockDim.x) hs[t] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    if (threadIdx.x == 0) {
        if (!keep) p.env_done[env] = env_coll ? 1 : 0;
        if (p.init_state == nullptr) p.reset_epoch[env] = epoch + (uint32_t)attempts;
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

// -------------------------------------------------------------------------------------------------
// Large-swarm reset, one rejection attempt per launch, grid = (tiles, E): the env's rows are spread over several
// CTAs (one CTA per env does 2048^2 pairs on ONE SM: 315 us for a single restarting env of 2048 agents).
// Attempt `p.step_offset` of this flock_reset call: every CTA of an active env re-draws the WHOLE env into shared
// memory (N / rows Philox draws per thread -- cheap next to N pairs per row, and no inter-CTA dependency), senses
// its own rows, and the last CTA of the env to finish (arrival counter, as in the step kernels) decides: collision
// free (or out of attempts) -> close the episode counters exactly as flock_reset_tiled_kernel does; otherwise
// p.reset_need[env] stays set and the next attempt launch (or the single-CTA loop kernel, which continues at
// attempt `step_offset` for the rare env that needs more than kFastResetAttempts) re-draws it.
// Same draws, same arithmetic, same outputs as the single-CTA kernel (gym_flock_v2.py:85-108).
template <int K>
__global__ void __launch_bounds__(kMaxTileThreads) flock_reset_attempt_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_last;
    const int N = p.N, k = p.k;
    const int env = blockIdx.y;
    const int attempt = (int)p.step_offset;
    const bool active = attempt == 0 ? (p.env_mask == nullptr || p.env_mask[env] != 0) : (p.reset_need[env] != 0);
    if (!active) {
        if (attempt == 0 && blockIdx.x == 0 && threadIdx.x == 0) p.reset_need[env] = 0;
        return;
    }
    const TileSmem sm = carve(smem, N, false);
    const size_t base = (size_t)env * N;
    const size_t EN = (size_t)p.E * N;
    const uint32_t epoch = p.reset_epoch[env];
    const int max_att = p.init_state != nullptr ? 1 : p.max_attempts;
    const bool keep = (p.reset_flags & FLOCK_RESET_KEEP_OUTPUTS) != 0;
    const int row0 = blockIdx.x * blockDim.x;
    if (threadIdx.x < padded_agents(N) - N) {
        sm.sx[N + threadIdx.x] = kInf;
        sm.sy[N + threadIdx.x] = 0.0f;
    }
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        float x, y, h;
        if (p.init_state != nullptr) {
            x = p.init_state[base + a];
            y = p.init_state[EN + base + a];
            h = p.init_state[2 * EN + base + a];
        } else {
            const uint4 r = philox4x32_10((uint32_t)(p.env_offset + env), (uint32_t)a, epoch + (uint32_t)attempt, kTagReset,
                                          p.seed_lo, p.seed_hi);
            const float span = p.range_lo - p.reset_hi;
            const float tx = span * u24(r.x);
            x = tx + p.reset_hi;
            const float ty = span * u24(r.y);
            y = ty + p.reset_hi;
            const float th = (0.0f - p.heading_hi) * u24(r.z);
            h = th + p.heading_hi;
        }
        x = wrap_coord(x, p.B, p.fill_hi, p.fill_lo);
        y = wrap_coord(y, p.B, p.fill_hi, p.fill_lo);
        sm.sx[a] = x;
        sm.sy[a] = y;
        if (a >= row0 && a < row0 + (int)blockDim.x) {   // this CTA's rows: install the state
            p.xo[base + a] = x;
            p.yo[base + a] = y;
            p.ho[base + a] = h;
            p.prev_h[base + a] = 0.0f;
            if (p.vx != nullptr) {
                p.vx[base + a] = 0.0f;
                p.vy[base + a] = 0.0f;
            }
        }
    }
    __syncthreads();
    const int i = row0 + threadIdx.x;
    const bool has_row = i < N;
    TopK<K> t;
    const float x = sm.sx[has_row ? i : 0], y = sm.sy[has_row ? i : 0];
    knn_tiled<K, false>(sm.sx, sm.sy, sm.cand + threadIdx.x, i, N, x, y, p.B, has_row ? kFltMax : -1.0f, t);
    bool coll = false;
    if (has_row) {
        float dist[K];
        coll = finish_row<K>(t, k, p.sensor_range, p.reset_cd, dist);
        const size_t idx = base + i;
        write_obs_t<K>(p, env, i, idx, dist, true);
        if (p.nn != nullptr) store_row_t<int, K>(p.nn + idx * k, t.idx, k);
        if (!keep) {
            p.reward[idx] = 0.0f;
            p.agent_done[idx] = coll ? 1 : 0;
        }
    }
    const int cta_coll = __syncthreads_or(coll ? 1 : 0);
    unsigned int* arrive = p.tile_scratch + env;
    unsigned int* collide = p.tile_scratch + p.E + env;
    if (threadIdx.x == 0) {
        if (cta_coll) atomicAdd(collide, 1u);
        __threadfence();
        const unsigned prev = atomicAdd(arrive, 1u);
        s_last = prev == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // last CTA of the env for this attempt
    __threadfence();
    int env_coll = 0, again = 0;
    if (threadIdx.x == 0) {
        env_coll = atomicExch(collide, 0u) != 0u ? 1 : 0;
        *arrive = 0u;
        again = env_coll && attempt + 1 < max_att;
        s_last = again ? 2 : (env_coll ? 3 : 1);
    }
    __syncthreads();
    again = s_last == 2;
    env_coll = s_last >= 2;
    if (again) {
        if (threadIdx.x == 0) p.reset_need[env] = 1;
        return;
    }
    if (p.hint_slots != nullptr) {   // the pruned kernel's slot hints describe the old episode: drop them
        const int PS = ((N + 31) / 32) * 32;
        uint4* hs = reinterpret_cast<uint4*>(p.hint_slots) + (size_t)env * PS;
        for (int tt = threadIdx.x; tt < PS; tt += blockDim.x) hs[tt] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    if (threadIdx.x == 0) {
        const int attempts = attempt + 1;
        p.reset_need[env] = 0;
        if (!keep) p.env_done[env] = env_coll ? 1 : 0;
        if (p.init_state == nullptr) p.reset_epoch[env] = epoch + (uint32_t)attempts;
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

// -------------------------------------------------------------------------------------------------
// rows per CTA: balance the most loaded SM (its CTAs x rows), prefer larger tiles on ties
static int choose_rows(int N, int E, int sm_count) {
    int best = kMaxTileThreads;
    long long best_cost = -1;
    for (int r = kMaxTileThreads; r >= 128; r -= 32) {
        const long long ctas = (long long)((N + r - 1) / r) * E;
        const long long cost = ((ctas + sm_count - 1) / sm_count) * r;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = r;
        }
    }
    if (N < best) best = (N + 31) & ~31;
    return best;
}

// few envs x large swarm: fewer 128-row tiles than SMs -> warp-per-row kernel (bitonic top-k merge)
template <int V, int K, bool PER>
static cudaError_t launch_rowwarp(const Params& p, int sm_count, cudaStream_t s) {
    const int warps = kMaxTileThreads / 32;
    long long rpw = ((long long)p.N * p.E + (long long)warps * sm_count * 2 - 1) / ((long long)warps * sm_count * 2);
    if (rpw < 1) rpw = 1;
    if (rpw > 64) rpw = 64;
    const int rows_per_cta = warps * (int)rpw;
    const dim3 grid((p.N + rows_per_cta - 1) / rows_per_cta, p.E);
    flock_step_rowwarp_kernel<V, K, PER><<<grid, kMaxTileThreads, tiled_smem_bytes(p.N, kMaxTileThreads, true), s>>>(
        p, (int)rpw);
    return cudaGetLastError();
}

// tiled_mode: 0 auto, 1 thread-per-row, 2 warp-per-row (flock_set_tiled_mode). Measured on B200
// (tools/single_env_latency.py, N = 2048, k = 8): E = 1: 51 -> 16 us with warp-per-row, E = 4: 52 -> 43,
// E = 16: 58 -> 180 -- the warp-per-row kernel wins while the whole job has <= ~55 rows per SM.
static bool use_rowwarp(const Params& p, int sm_count, int tiled_mode) {
    if (tiled_mode == 1) return false;
    if (tiled_mode == 2) return true;
    return (long long)p.N * p.E <= 55LL * sm_count;
}

static bool prune_enabled() {   // FLOCK_PRUNE=0: plain all-pairs scan (same results, for comparison)
    static const bool prune = [] {
        const char* v = getenv("FLOCK_PRUNE");
        return v == nullptr || v[0] != '0';
    }();
    return prune;
}
static bool use_pruned(int /*variant*/, const Params& p, int sm_count, int tiled_mode) {
    return prune_enabled() && p.perm != nullptr && p.sorted_xy != nullptr && !use_rowwarp(p, sm_count, tiled_mode);
}
int tiled_step_launches(int, const Params&, int, int) { return 2; }   // integrate pre-pass + sensing kernel

template <int V, int K, bool PER>
static cudaError_t launch_tiled_vkp(const Params& p, int sm_count, int tiled_mode, cudaStream_t s) {
    if (use_pruned(V, p, sm_count, tiled_mode)) return launch_pruned<V, (K < 4 ? 4 : K), PER>(p, sm_count, s);
    {   // launch 1 of 2: integrate every agent once, in place
        cudaError_t e = launch_integrate<V, false>(p, s);
        if (e != cudaSuccess) return e;
    }
    if (use_rowwarp(p, sm_count, tiled_mode)) return launch_rowwarp<V, (K < 4 ? 4 : K), PER>(p, sm_count, s);
    const int rows = choose_rows(p.N, p.E, sm_count);
    const dim3 grid((p.N + rows - 1) / rows, p.E);
    flock_step_tiled_kernel<V, K, PER><<<grid, rows, tiled_smem_bytes(p.N, rows, false), s>>>(p);
    return cudaGetLastError();
}
template <int V, bool PER>
static cudaError_t launch_tiled_vp(const Params& p, int sm_count, int tiled_mode, cudaStream_t s) {
    if (p.k <= 3) return launch_tiled_vkp<V, 3, PER>(p, sm_count, tiled_mode, s);
    if (p.k == 4) return launch_tiled_vkp<V, 4, PER>(p, sm_count, tiled_mode, s);
    return launch_tiled_vkp<V, 8, PER>(p, sm_count, tiled_mode, s);
}

cudaError_t launch_step_tiled(int variant, bool periodic, const Params& p, int sm_count, int tiled_mode,
                              cudaStream_t s) {
    switch (variant) {
        case FLOCK_V2:
            return periodic ? launch_tiled_vp<FLOCK_V2, true>(p, sm_count, tiled_mode, s)
                            : launch_tiled_vp<FLOCK_V2, false>(p, sm_count, tiled_mode, s);
        case FLOCK_UW:
            return launch_tiled_vp<FLOCK_UW, false>(p, sm_count, tiled_mode, s);
        default:
            return launch_tiled_vp<FLOCK_UWD, false>(p, sm_count, tiled_mode, s);
    }
}

// Large-swarm reset = up to kFastResetAttempts multi-CTA attempt launches (the common case: a sparse world accepts the
// first draw) + the single-CTA loop kernel for the envs that still collide and have attempts left.
constexpr int kFastResetAttempts = 2;
cudaError_t launch_reset_tiled(const Params& p, uint8_t* need, int* launches, cudaStream_t s) {
    const int max_att = p.init_state != nullptr ? 1 : p.max_attempts;
    const int fast = need == nullptr ? 0 : (max_att < kFastResetAttempts ? max_att : kFastResetAttempts);
    int rows = (p.N + 31) & ~31;
    if (rows > kMaxTileThreads) rows = kMaxTileThreads;
    const dim3 grid((p.N + rows - 1) / rows, p.E);
    const size_t smem_a = tiled_smem_bytes(p.N, rows, false);
    *launches = 0;
    for (int a = 0; a < fast; ++a) {
        Params q = p;
        q.step_offset = (uint32_t)a;
        q.reset_need = need;
        if (p.k <= 3) flock_reset_attempt_kernel<3><<<grid, rows, smem_a, s>>>(q);
        else if (p.k == 4) flock_reset_attempt_kernel<4><<<grid, rows, smem_a, s>>>(q);
        else flock_reset_attempt_kernel<8><<<grid, rows, smem_a, s>>>(q);
        *launches += 1;
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (max_att <= fast) return cudaSuccess;
    Params q = p;
    q.step_offset = (uint32_t)fast;
    if (fast > 0) q.env_mask = need;      // only the envs the attempt launches left colliding
    int rows1 = (p.N + 31) & ~31;
    if (rows1 > kResetThreads) rows1 = kResetThreads;
    const size_t smem = tiled_smem_bytes(p.N, rows1, true);
    if (p.k <= 3) flock_reset_tiled_kernel<3><<<p.E, rows1, smem, s>>>(q);
    else if (p.k == 4) flock_reset_tiled_kernel<4><<<p.E, rows1, smem, s>>>(q);
    else flock_reset_tiled_kernel<8><<<p.E, rows1, smem, s>>>(q);
    *launches += 1;
    return cudaGetLastError();
}

template <typename Kern>
static cudaError_t opt_in(Kern kern, size_t bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// The opt-in is an attribute of the FUNCTION on the current device, not of an env handle, so it is always
// raised to the worst case (FLOCK_MAX_AGENTS): a later, smaller env can then never lower the limit under an
// earlier, larger one. Called by every flock_create after cudaSetDevice (idempotent per device).
cudaError_t tiled_configure(int /*num_agents*/) {
    const int num_agents = FLOCK_MAX_AGENTS;
    const size_t b = tiled_smem_bytes(num_agents, kMaxTileThreads, true);
    cudaError_t e = cudaSuccess;
    {
        const size_t bp = pruned_smem_bytes(num_agents, kMaxTileThreads);
        if (bp > 47 * 1024) {
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_V2, 4, true>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_V2, 8, true>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_V2, 4, false>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_V2, 8, false>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_UW, 4, false>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_UW, 8, false>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_UWD, 4, false>, bp);
            if (e == cudaSuccess) e = opt_in(flock_step_pruned_kernel<FLOCK_UWD, 8, false>, bp);
        }
        const int G = order_grid_side(num_agents);
        const size_t br = (size_t)G * G * sizeof(int) + (size_t)num_agents * 2 * sizeof(unsigned short);
        if (br > 46 * 1024 && e == cudaSuccess) e = opt_in(flock_perm_refresh_kernel, br);
    }
    if (b <= 48 * 1024) return e;
#define FLOCK_OPT(V, K, PER) \
    if (e == cudaSuccess) e = opt_in(flock_step_tiled_kernel<V, K, PER>, b);
    FLOCK_OPT(FLOCK_V2, 3, true) FLOCK_OPT(FLOCK_V2, 4, true) FLOCK_OPT(FLOCK_V2, 8, true)
    FLOCK_OPT(FLOCK_V2, 3, false) FLOCK_OPT(FLOCK_V2, 4, false) FLOCK_OPT(FLOCK_V2, 8, false)
    FLOCK_OPT(FLOCK_UW, 3, false) FLOCK_OPT(FLOCK_UW, 4, false) FLOCK_OPT(FLOCK_UW, 8, false)
    FLOCK_OPT(FLOCK_UWD, 3, false) FLOCK_OPT(FLOCK_UWD, 4, false) FLOCK_OPT(FLOCK_UWD, 8, false)
#undef FLOCK_OPT
#define FLOCK_OPT_RW(V, K, PER) \
    if (e == cudaSuccess) e = opt_in(flock_step_rowwarp_kernel<V, K, PER>, b);
    FLOCK_OPT_RW(FLOCK_V2, 4, true) FLOCK_OPT_RW(FLOCK_V2, 8, true) FLOCK_OPT_RW(FLOCK_V2, 4, false)
    FLOCK_OPT_RW(FLOCK_V2, 8, false) FLOCK_OPT_RW(FLOCK_UW, 4, false) FLOCK_OPT_RW(FLOCK_UW, 8, false)
    FLOCK_OPT_RW(FLOCK_UWD, 4, false) FLOCK_OPT_RW(FLOCK_UWD, 8, false)
#undef FLOCK_OPT_RW
    if (e == cudaSuccess) e = opt_in(flock_reset_attempt_kernel<3>, b);
    if (e == cudaSuccess) e = opt_in(flock_reset_attempt_kernel<4>, b);
    if (e == cudaSuccess) e = opt_in(flock_reset_attempt_kernel<8>, b);
    const size_t br2 = tiled_smem_bytes(num_agents, kResetThreads, true);
    if (e == cudaSuccess) e = opt_in(flock_reset_tiled_kernel<3>, br2);
    if (e == cudaSuccess) e = opt_in(flock_reset_tiled_kernel<4>, br2);
    if (e == cudaSuccess) e = opt_in(flock_reset_tiled_kernel<8>, br2);
    return e;
}

}  // namespace flock
