// flock_small.cu -- warp-per-env-group kernels for small swarms (N <= 32).
//
// Mapping: lane = agent; a warp owns G = floor(32 / N) consecutive envs (N = 10 -> 3 envs, 30
// live lanes; N = 16 -> 2; N = 32 -> 1), so the warp's slice of every SoA array is one contiguous
// run of G*N floats -> fully coalesced loads/stores. New positions are staged in shared memory
// and the all-pairs loop reads them back as broadcast float4 (4 neighbours per LDS.128 pair);
// the per-row k-smallest list lives in registers. One launch = one env step
// (integrate -> wrap -> all-pairs range -> k-NN -> collisions/dones -> reward -> obs).
#include "flock_small_impl.cuh"

namespace flock {

template <int K>
__global__ void __launch_bounds__(kSmallThreads) flock_reset_small_kernel(const __grid_constant__ Params p) {
    __shared__ __align__(16) float s_stage[kSmallWarps][2][kSlots];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* sx = s_stage[wib][0];
    float* sy = s_stage[wib][1];
    const int N = p.N, G = p.G;
    const LaneMap m = lane_map(lane, N, G, p.g_magic);
    const int num_tasks = p.num_tasks;
    const int warps_total = gridDim.x * kSmallWarps;
    for (int task = blockIdx.x * kSmallWarps + wib; task < num_tasks; task += warps_total) {
        const int env = task * G + m.g;
        bool live = m.lane_ok && env < p.E;
        if (live && p.env_mask != nullptr) live = p.env_mask[env] != 0;
        const size_t idx = (size_t)(live ? env : 0) * N + m.a;
        reset_groups<K>(p, m, sx, sy, env, idx, live, (p.reset_flags & FLOCK_RESET_KEEP_OUTPUTS) != 0);
    }
}

// canonical random actions into a buffer (flock_random_actions)
template <int V>
__global__ void flock_random_actions_kernel(const __grid_constant__ Params p, float* out) {
    const size_t n = (size_t)p.E * p.N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int env = (int)(i / p.N), a = (int)(i - (size_t)env * p.N);
        float a0, a1;
        random_action<V>(p, p.env_offset + env, a, (uint32_t)p.ep_len[env] + p.step_offset, p.reset_epoch[env], a0, a1);
        if (V == FLOCK_UWD) out[i] = a0;
        else reinterpret_cast<float2*>(out)[i] = make_float2(a0, a1);
    }
}

// Optional sensing noise (north-star extension; the reference has none): the newest row of range
// observations becomes clamp(d + std * z, 0, sensor_range), z ~ N(0,1) from the Philox stream
// (global env, agent + 65536 * call, ep_len[env], tag 3 + 4 * reset_epoch). Collisions, dones and
// rewards keep using the true ranges. Launched right after step / reset only when std > 0, so the
// default (parity) path carries no trace of it.
__global__ void flock_range_noise_kernel(const __grid_constant__ Params p) {
    const size_t n = (size_t)p.E * p.N;
    const int k = p.k;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int env = (int)(i / p.N), a = (int)(i - (size_t)env * p.N);
        if (p.env_mask != nullptr && p.env_mask[env] == 0) continue;
        const uint32_t epoch = (uint32_t)p.ep_len[env], word = stream_word(kTagRange, p.reset_epoch[env]);
        // newest row: first row of the window, or the head slot of the ring
        float* o = (p.H > 1 && p.obs_head != nullptr) ? ring_row(p, env, p.obs_head[env], a) : p.obs + i * (size_t)(p.H * k);
        for (int c = 0; c * 4 < k; ++c) {
            const uint4 r = philox4x32_10((uint32_t)(p.env_offset + env), (uint32_t)a + 65536u * (uint32_t)c, epoch, word,
                                          p.seed_lo, p.seed_hi);
            float z[4];
            normal2(r.x, r.y, z[0], z[1]);
            normal2(r.z, r.w, z[2], z[3]);
            for (int s = 0; s < 4 && c * 4 + s < k; ++s) {
                const float nz = p.range_noise_std * z[s];
                float d = o[c * 4 + s] + nz;
                d = fminf(fmaxf(d, 0.0f), p.sensor_range);
                o[c * 4 + s] = d;
            }
        }
    }
}

// Materialise the newest-first (E, N, H, k) window of a ring-layout env (uw): out[e][a][r][:] = ring[e][(head[e] + r) % H][a][:].
// On request only (single-env facade, host copy-out, PyTorch policies): the consumers we own read the ring in place.
__global__ void flock_obs_window_kernel(const __grid_constant__ Params p, float* out) {
    const int H = p.H, k = p.k, N = p.N;
    const size_t n = (size_t)p.E * N * H * k;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % k);
        const size_t q = i / k;
        const int r = (int)(q % H);
        const size_t ea = q / H;
        const int a = (int)(ea % N), env = (int)(ea / N);
        const int slot = (p.obs_head[env] + r) % H;
        out[i] = ring_row(p, env, slot, a)[c];
    }
}

// newest range row of every agent -> out[E][N][k] (rollout fallback for large swarms / sensing noise)
__global__ void flock_newest_row_kernel(const __grid_constant__ Params p, float* out) {
    const int k = p.k, N = p.N;
    const size_t n = (size_t)p.E * N * k;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % k);
        const size_t ea = i / k;
        const int a = (int)(ea % N), env = (int)(ea / N);
        const float* o = (p.H > 1 && p.obs_head != nullptr) ? ring_row(p, env, p.obs_head[env], a) : p.obs + ea * (size_t)(p.H * k);
        out[i] = o[c];
    }
}

__global__ void flock_debug_sincos_kernel(const float* h, int n, float* sn, float* cs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sincos_canon(h[i], sn[i], cs[i]);
}
__global__ void flock_debug_normal2_kernel(const uint32_t* w, int n, float* z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) normal2(w[2 * i], w[2 * i + 1], z[2 * i], z[2 * i + 1]);
}
__global__ void flock_debug_philox_kernel(const uint32_t* ck, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint4 r = philox4x32_10(ck[6 * i], ck[6 * i + 1], ck[6 * i + 2], ck[6 * i + 3], ck[6 * i + 4], ck[6 * i + 5]);
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
}

// -------------------------------------------------------------------------------------------------
// host-side dispatch. The step kernels live in flock_small_<variant>.cu (explicit instantiations).
// Neighbour indices are tracked only when the caller bound an nn_idx buffer: the reference keeps
// `nearest_neighbors` in v2 (gym_flock_v2.py:150, VecEnv default track_neighbors=True) and discards
// the indices in uw / uwd; without the buffer the values-only selection network is used.
// -------------------------------------------------------------------------------------------------
extern template cudaError_t launch_step_small_vpi<FLOCK_V2, true, true>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_V2, false, true>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_V2, true, false>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_V2, false, false>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_UW, false, true>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_UW, false, false>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_UWD, false, true>(const Params&, int, int, cudaStream_t);
extern template cudaError_t launch_step_small_vpi<FLOCK_UWD, false, false>(const Params&, int, int, cudaStream_t);

cudaError_t launch_step_small(int variant, bool periodic, const Params& p, int mode, int sm_count, cudaStream_t s) {
    const bool idx = p.nn != nullptr;
    switch (variant) {
        case FLOCK_V2:
            if (idx)
                return periodic ? launch_step_small_vpi<FLOCK_V2, true, true>(p, mode, sm_count, s)
                                : launch_step_small_vpi<FLOCK_V2, false, true>(p, mode, sm_count, s);
            return periodic ? launch_step_small_vpi<FLOCK_V2, true, false>(p, mode, sm_count, s)
                            : launch_step_small_vpi<FLOCK_V2, false, false>(p, mode, sm_count, s);
        case FLOCK_UW:
            return idx ? launch_step_small_vpi<FLOCK_UW, false, true>(p, mode, sm_count, s)
                       : launch_step_small_vpi<FLOCK_UW, false, false>(p, mode, sm_count, s);
        default:
            return idx ? launch_step_small_vpi<FLOCK_UWD, false, true>(p, mode, sm_count, s)
                       : launch_step_small_vpi<FLOCK_UWD, false, false>(p, mode, sm_count, s);
    }
}

cudaError_t launch_reset_small(const Params& p, int sm_count, cudaStream_t s) {
    const int grid = small_grid(p, sm_count);
    if (p.k <= 3) flock_reset_small_kernel<3><<<grid, kSmallThreads, 0, s>>>(p);
    else if (p.k == 4) flock_reset_small_kernel<4><<<grid, kSmallThreads, 0, s>>>(p);
    else flock_reset_small_kernel<8><<<grid, kSmallThreads, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_random_actions(int variant, const Params& p, float* out, int sm_count, cudaStream_t s) {
    const size_t n = (size_t)p.E * p.N;
    int grid = (int)((n + 255) / 256);
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    if (variant == FLOCK_V2) flock_random_actions_kernel<FLOCK_V2><<<grid, 256, 0, s>>>(p, out);
    else if (variant == FLOCK_UW) flock_random_actions_kernel<FLOCK_UW><<<grid, 256, 0, s>>>(p, out);
    else flock_random_actions_kernel<FLOCK_UWD><<<grid, 256, 0, s>>>(p, out);
    return cudaGetLastError();
}

cudaError_t launch_range_noise(const Params& p, int sm_count, cudaStream_t s) {
    const size_t n = (size_t)p.E * p.N;
    int grid = (int)((n + 255) / 256);
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    flock_range_noise_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_newest_row(const Params& p, float* out, int sm_count, cudaStream_t s) {
    const size_t n = (size_t)p.E * p.N * p.k;
    int grid = (int)((n + 255) / 256);
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    flock_newest_row_kernel<<<grid, 256, 0, s>>>(p, out);
    return cudaGetLastError();
}

cudaError_t launch_obs_window(const Params& p, float* out, int sm_count, cudaStream_t s) {
    const size_t n = (size_t)p.E * p.N * p.H * p.k;
    int grid = (int)((n + 255) / 256);
    if (grid > sm_count * 8) grid = sm_count * 8;
    if (grid < 1) grid = 1;
    flock_obs_window_kernel<<<grid, 256, 0, s>>>(p, out);
    return cudaGetLastError();
}

cudaError_t launch_debug_sincos(const float* h, int n, float* sn, float* cs, cudaStream_t s) {
    flock_debug_sincos_kernel<<<(n + 255) / 256, 256, 0, s>>>(h, n, sn, cs);
    return cudaGetLastError();
}
cudaError_t launch_debug_normal2(const uint32_t* w, int n, float* z, cudaStream_t s) {
    flock_debug_normal2_kernel<<<(n + 255) / 256, 256, 0, s>>>(w, n, z);
    return cudaGetLastError();
}
cudaError_t launch_debug_philox(const uint32_t* ck, int n, uint32_t* out, cudaStream_t s) {
    flock_debug_philox_kernel<<<(n + 255) / 256, 256, 0, s>>>(ck, n, out);
    return cudaGetLastError();
}

}  // namespace flock
