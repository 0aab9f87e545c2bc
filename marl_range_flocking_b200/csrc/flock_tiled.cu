// flock_tiled.cu -- tiled all-pairs kernels for large swarms (32 < N <= FLOCK_MAX_AGENTS).
//
// One CTA = one 128-row i-tile of one env. The env's whole old state (x, y, heading) and its
// actions are pulled into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier); every
// CTA of the env re-integrates all N agents in shared memory (about 3 % of the pair work at
// N = 2048, and it removes any grid-wide dependency), writes back only its own rows to the OTHER
// state copy (ping-pong, so concurrent CTAs of the same env still read the old state), then runs
// its rows against all j as broadcast float4 shared-memory reads with a register-resident
// k-smallest list per row. One launch = one env step, as on the small path.
#include "flock_device.cuh"
#include "flock_launch.h"

namespace flock {

constexpr int kTileThreads = 128;

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

__host__ __device__ __forceinline__ int padded_agents(int N) { return ((N + 3) & ~3) + 4; }

size_t tiled_smem_bytes(int num_agents) { return (size_t)padded_agents(num_agents) * 5 * sizeof(float); }

// shared-memory carve-up: sx | sy | sh | sa (actions, 2 floats per agent)
struct TileSmem {
    float *sx, *sy, *sh, *sa;
};
__device__ __forceinline__ TileSmem carve(float* base, int N) {
    const int P = padded_agents(N);
    return TileSmem{base, base + P, base + 2 * P, base + 3 * P};
}

__device__ __forceinline__ void stage_env(const Params& p, const TileSmem& sm, uint64_t* bar, int env, int aw,
                                          bool with_actions) {
    const int N = p.N;
    const size_t base = (size_t)env * N;
    if ((N & 3) == 0) {  // rows are 16-byte multiples and 16-byte aligned: TMA bulk path
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            const uint32_t row = (uint32_t)N * 4u;
            const uint32_t act = with_actions ? row * (uint32_t)aw : 0u;
            mbar_expect_tx(bar, 3u * row + act);
            tma_load_1d(sm.sx, p.x + base, row, bar);
            tma_load_1d(sm.sy, p.y + base, row, bar);
            tma_load_1d(sm.sh, p.h + base, row, bar);
            if (with_actions) tma_load_1d(sm.sa, p.actions + base * aw, act, bar);
        }
        __syncthreads();          // barrier initialised and armed before anyone polls it
        mbar_wait(bar, 0);
    } else {
        for (int a = threadIdx.x; a < N; a += kTileThreads) {
            sm.sx[a] = p.x[base + a];
            sm.sy[a] = p.y[base + a];
            sm.sh[a] = p.h[base + a];
            if (with_actions)
                for (int c = 0; c < aw; ++c) sm.sa[a * aw + c] = p.actions[(base + a) * aw + c];
        }
        __syncthreads();
    }
}

template <int K, bool PER>
__device__ __forceinline__ void knn_tiled(const float* sx, const float* sy, int i, int N, float x, float y, float B,
                                          TopK<K>& t) {
    t.init();
    const float4* px = reinterpret_cast<const float4*>(sx);
    const float4* py = reinterpret_cast<const float4*>(sy);
    const int n4 = (N + 3) >> 2;
#pragma unroll 2
    for (int j4 = 0; j4 < n4; ++j4) {
        const float4 X = px[j4];
        const float4 Y = py[j4];
        const int j = j4 << 2;
        float d0 = pair_d2<PER>(x, y, X.x, Y.x, B);
        float d1 = pair_d2<PER>(x, y, X.y, Y.y, B);
        float d2 = pair_d2<PER>(x, y, X.z, Y.z, B);
        float d3 = pair_d2<PER>(x, y, X.w, Y.w, B);
        d0 = (j == i) ? kInf : d0;
        d1 = (j + 1 == i) ? kInf : d1;
        d2 = (j + 2 == i) ? kInf : d2;
        d3 = (j + 3 == i) ? kInf : d3;
        if (d0 < t.worst()) t.insert(d0, j);
        if (d1 < t.worst()) t.insert(d1, j + 1);
        if (d2 < t.worst()) t.insert(d2, j + 2);
        if (d3 < t.worst()) t.insert(d3, j + 3);
    }
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
__device__ __forceinline__ void write_obs_t(const Params& p, size_t idx, const float (&dist)[K], bool fresh) {
    const int k = p.k;
    if (p.H == 1) {
        store_row_t<float, K>(p.obs + idx * k, dist, k);
        return;
    }
    float* o = p.obs + idx * (size_t)(p.H * k);
    for (int t = (p.H - 1) * k - 1; t >= 0; --t) o[t + k] = fresh ? 0.0f : o[t];
#pragma unroll
    for (int s = 0; s < K; ++s)
        if (s < k) o[s] = dist[s];
}

// step, grid = (tiles per env, E).
template <int V, int K, bool PER>
__global__ void __launch_bounds__(kTileThreads) flock_step_tiled_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int N = p.N, k = p.k;
    const int env = blockIdx.y, tile = blockIdx.x;
    const int aw = (V == FLOCK_UWD) ? 1 : 2;
    const TileSmem sm = carve(smem, N);
    stage_env(p, sm, &bar, env, aw, true);

    // integrate every agent of the env in shared memory; keep this thread's own row in registers
    const int i = tile * kTileThreads + threadIdx.x;
    const bool has_row = i < N;
    const size_t base = (size_t)env * N;
    float x = 0.f, y = 0.f, h = 0.f, vx = 0.f, vy = 0.f;
    for (int a = threadIdx.x; a < N; a += kTileThreads) {
        float ax = sm.sx[a], ay = sm.sy[a], ah = sm.sh[a];
        const float a0 = sm.sa[a * aw], a1 = (aw == 2) ? sm.sa[a * aw + 1] : 0.0f;
        float nzu = 0.f, nzw = 0.f;
        if (V == FLOCK_UWD) {
            if (p.noise != nullptr) {
                const float2 nz = reinterpret_cast<const float2*>(p.noise)[base + a];
                nzu = nz.x;
                nzw = nz.y;
            } else if (p.noise_std > 0.0f) {
                act_noise(p, p.env_offset + env, a, (uint32_t)p.ep_len[env], p.reset_epoch[env], nzu, nzw);
            }
        }
        float avx, avy;
        integrate_agent<V>(p, a0, a1, nzu, nzw, ax, ay, ah, avx, avy);
        sm.sx[a] = ax;
        sm.sy[a] = ay;
        sm.sh[a] = ah;
        if (a == i) {
            x = ax; y = ay; h = ah; vx = avx; vy = avy;
        }
    }
    if (threadIdx.x < padded_agents(N) - N) {
        sm.sx[N + threadIdx.x] = kInf;
        sm.sy[N + threadIdx.x] = 0.0f;
    }
    __syncthreads();

    float comx = 0.f, comy = 0.f, hmean = 0.f;
    if (V == FLOCK_UW) {
        float sx_ = 0.f, sy_ = 0.f;
        for (int j = 0; j < N; ++j) {
            sx_ = sx_ + sm.sx[j];
            sy_ = sy_ + sm.sy[j];
        }
        comx = __fdiv_rn(sx_, (float)N);
        comy = __fdiv_rn(sy_, (float)N);
    }
    if (V == FLOCK_UWD) {
        float sh_ = 0.f;
        for (int j = 0; j < N; ++j) sh_ = sh_ + sm.sh[j];
        hmean = __fdiv_rn(sh_, (float)N);
    }

    long long fx = 0;
    bool coll = false;
    if (has_row) {
        TopK<K> t;
        knn_tiled<K, PER>(sm.sx, sm.sy, i, N, x, y, p.B, t);
        float dist[K];
        coll = finish_row<K>(t, k, p.sensor_range, p.cd, dist);
        const size_t idx = base + i;
        float prev_h = 0.f;
        if (V == FLOCK_UW) prev_h = p.prev_h[idx];
        const float rew = agent_reward<V>(p, coll, x, y, h, prev_h, comx, comy, hmean);
        fx = reward_fx(rew);
        p.xo[idx] = x;
        p.yo[idx] = y;
        p.ho[idx] = h;
        if (V == FLOCK_UW && !(prev_h == h)) p.prev_h[idx] = h;
        if (p.vx != nullptr) {
            p.vx[idx] = vx;
            p.vy[idx] = vy;
        }
        write_obs_t<K>(p, idx, dist, false);
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
            *arrive = 0u;
        }
    }
}

// reset, grid = E, one CTA per env, bounded rejection loop (gym_flock_v2.py:85-108)
template <int K>
__global__ void __launch_bounds__(kTileThreads) flock_reset_tiled_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) float smem[];
    const int N = p.N, k = p.k;
    const int env = blockIdx.x;
    if (p.env_mask != nullptr && p.env_mask[env] == 0) return;
    const TileSmem sm = carve(smem, N);
    const size_t base = (size_t)env * N;
    const size_t EN = (size_t)p.E * N;
    const uint32_t epoch = p.reset_epoch[env];
    const int max_att = p.init_state != nullptr ? 1 : p.max_attempts;
    int attempts = 0;
    int env_coll = 1;
    const bool keep = (p.reset_flags & FLOCK_RESET_KEEP_OUTPUTS) != 0;
    if (threadIdx.x < padded_agents(N) - N) {
        sm.sx[N + threadIdx.x] = kInf;
        sm.sy[N + threadIdx.x] = 0.0f;
    }
    while (env_coll && attempts < max_att) {
        __syncthreads();
        for (int a = threadIdx.x; a < N; a += kTileThreads) {
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
            sm.sx[a] = wrap_coord(x, p.B, p.rigid);
            sm.sy[a] = wrap_coord(y, p.B, p.rigid);
            sm.sh[a] = h;
        }
        attempts += 1;
        __syncthreads();
        int coll_any = 0;
        for (int i = threadIdx.x; i < N; i += kTileThreads) {
            TopK<K> t;
            const float x = sm.sx[i], y = sm.sy[i];
            knn_tiled<K, false>(sm.sx, sm.sy, i, N, x, y, p.B, t);
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
            write_obs_t<K>(p, idx, dist, true);
            if (p.nn != nullptr) store_row_t<int, K>(p.nn + idx * k, t.idx, k);
            if (!keep) {
                p.reward[idx] = 0.0f;
                p.agent_done[idx] = coll ? 1 : 0;
            }
        }
        env_coll = __syncthreads_or(coll_any);
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
template <int V, int K, bool PER>
static cudaError_t launch_tiled_vkp(const Params& p, cudaStream_t s) {
    const dim3 grid((p.N + kTileThreads - 1) / kTileThreads, p.E);
    flock_step_tiled_kernel<V, K, PER><<<grid, kTileThreads, tiled_smem_bytes(p.N), s>>>(p);
    return cudaGetLastError();
}
template <int V, bool PER>
static cudaError_t launch_tiled_vp(const Params& p, cudaStream_t s) {
    if (p.k <= 3) return launch_tiled_vkp<V, 3, PER>(p, s);
    if (p.k == 4) return launch_tiled_vkp<V, 4, PER>(p, s);
    return launch_tiled_vkp<V, 8, PER>(p, s);
}

cudaError_t launch_step_tiled(int variant, bool periodic, const Params& p, cudaStream_t s) {
    switch (variant) {
        case FLOCK_V2:
            return periodic ? launch_tiled_vp<FLOCK_V2, true>(p, s) : launch_tiled_vp<FLOCK_V2, false>(p, s);
        case FLOCK_UW:
            return launch_tiled_vp<FLOCK_UW, false>(p, s);
        default:
            return launch_tiled_vp<FLOCK_UWD, false>(p, s);
    }
}

cudaError_t launch_reset_tiled(const Params& p, cudaStream_t s) {
    const size_t smem = tiled_smem_bytes(p.N);
    if (p.k <= 3) flock_reset_tiled_kernel<3><<<p.E, kTileThreads, smem, s>>>(p);
    else if (p.k == 4) flock_reset_tiled_kernel<4><<<p.E, kTileThreads, smem, s>>>(p);
    else flock_reset_tiled_kernel<8><<<p.E, kTileThreads, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename Kern>
static cudaError_t opt_in(Kern kern, size_t bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t tiled_configure(int num_agents) {
    const size_t b = tiled_smem_bytes(num_agents);
    if (b <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaSuccess;
#define FLOCK_OPT(V, K, PER) \
    if (e == cudaSuccess) e = opt_in(flock_step_tiled_kernel<V, K, PER>, b);
    FLOCK_OPT(FLOCK_V2, 3, true) FLOCK_OPT(FLOCK_V2, 4, true) FLOCK_OPT(FLOCK_V2, 8, true)
    FLOCK_OPT(FLOCK_V2, 3, false) FLOCK_OPT(FLOCK_V2, 4, false) FLOCK_OPT(FLOCK_V2, 8, false)
    FLOCK_OPT(FLOCK_UW, 3, false) FLOCK_OPT(FLOCK_UW, 4, false) FLOCK_OPT(FLOCK_UW, 8, false)
    FLOCK_OPT(FLOCK_UWD, 3, false) FLOCK_OPT(FLOCK_UWD, 4, false) FLOCK_OPT(FLOCK_UWD, 8, false)
#undef FLOCK_OPT
    if (e == cudaSuccess) e = opt_in(flock_reset_tiled_kernel<3>, b);
    if (e == cudaSuccess) e = opt_in(flock_reset_tiled_kernel<4>, b);
    if (e == cudaSuccess) e = opt_in(flock_reset_tiled_kernel<8>, b);
    return e;
}

}  // namespace flock
