// flock_api.cu -- the C ABI of libflock_b200.so (include/flock_b200.h): handle management,
// parameter marshalling and kernel dispatch. No torch types, no exceptions across the boundary.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "flock_device.cuh"
#include "flock_launch.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(FLOCK_E_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

}  // namespace

struct flock_env {
    flock_cfg_t cfg;
    flock_buffers_t b;
    bool bound;
    int device;
    int sm_count;
    int path;            // 0 small, 1 tiled
    int tiled_mode;      // 0 auto, 1 thread-per-row, 2 warp-per-row
    int auto_reset;      // flock_set_auto_reset: restart finished envs as part of flock_step
    int auto_reset_attempts;
    uint32_t step_index;
    uint64_t launches;
    float* stage_actions;  // device staging for host-call / step_n(tiled) actions
    float* stage_noise;
    float* stage_window;   // uw ring layout: materialised window for the host-call path (allocated on first use)
    unsigned int* tile_scratch;   // tiled path: per-env arrival / collision counters
    uint8_t* reset_need;          // tiled path: per-env "still colliding" flags of the multi-CTA reset
    int* tile_perm;               // tiled thread-per-row path: spatially sorted row order [E][N]
    int* tile_inv;                // its inverse (agent -> slot)
    float* sorted_xy;             // pruned path: per-env staging record (positions by slot, boxes, ids)
    unsigned short* hint_slots;   // pruned path: last step's neighbour slots per row
    float* env_sums;              // tiled path, uw / uwd: per-env sums behind the centre of mass / mean heading
    uint8_t* far_rows;            // pruned path: rows that left their spatial neighbourhood (wrap-around) since the last refresh
    unsigned long long* timeline; // -DFLOCK_TIMELINE developer builds only
    uint32_t perm_age;            // steps since the row order was refreshed
    unsigned long long* pair_counter;   // device counter of row x neighbour pairs evaluated by the pruned kernel
    const void* zc_host[5];       // last host buffers seen by flock_step_host and their device aliases
    void* zc_dev[5];
    bool zc_ok;
    size_t action_floats;
    cudaEvent_t host_event;       // flock_step_host_async / flock_wait_host
    bool host_event_live;
};

namespace {

using flock::Params;

int validate(const flock_cfg_t& c) {
    if (c.variant < FLOCK_V2 || c.variant > FLOCK_UWD) return fail(FLOCK_E_INVALID, "variant %d not in {0,1,2}", c.variant);
    if (c.num_envs < 1) return fail(FLOCK_E_INVALID, "num_envs must be >= 1 (got %d)", c.num_envs);
    if (c.k < 1 || c.k > FLOCK_MAX_K) return fail(FLOCK_E_INVALID, "k must be in [1, %d] (got %d)", FLOCK_MAX_K, c.k);
    if (c.num_agents < c.k + 1)   // the reference needs topk(k+1) <= N (gym_flock_v2.py:147)
        return fail(FLOCK_E_INVALID, "num_agents (%d) must be >= k+1 (%d)", c.num_agents, c.k + 1);
    if (c.num_agents > FLOCK_MAX_AGENTS)
        return fail(FLOCK_E_INVALID, "num_agents (%d) exceeds FLOCK_MAX_AGENTS (%d)", c.num_agents, FLOCK_MAX_AGENTS);
    const int want_h = c.variant == FLOCK_UW ? 4 : 1;
    if (c.obs_hist != want_h) return fail(FLOCK_E_INVALID, "obs_hist must be %d for variant %d", want_h, c.variant);
    if (c.periodic && c.variant != FLOCK_V2) return fail(FLOCK_E_INVALID, "periodic metric is only defined for v2");
    if (!(c.boundary > 0.0f)) return fail(FLOCK_E_INVALID, "boundary must be > 0");
    if (!(c.range_noise_std >= 0.0f)) return fail(FLOCK_E_INVALID, "range_noise_std must be >= 0");
    if ((long long)c.num_envs * c.num_agents > (1LL << 31) - 1) return fail(FLOCK_E_INVALID, "E*N too large");
    return FLOCK_OK;
}

Params make_params(const flock_env* e, float dt) {
    const flock_cfg_t& c = e->cfg;
    const flock_buffers_t& b = e->b;
    Params p;
    memset(&p, 0, sizeof(p));
    p.E = c.num_envs; p.N = c.num_agents; p.k = c.k; p.H = c.obs_hist;
    p.env_offset = c.env_offset;
    p.G = c.num_agents <= 32 ? 32 / c.num_agents : 0;
    {   // developer knob (profiles/README.md, "envs per warp" experiment): fewer env groups per warp = more, shorter warps
        static const int force_g = [] {
            const char* v = getenv("FLOCK_FORCE_G");
            return v != nullptr ? atoi(v) : 0;
        }();
        if (force_g > 0 && force_g < p.G) p.G = force_g;
    }
    p.sstride = (c.num_agents + 3) & ~3;
    p.g_magic = (65536 + c.num_agents - 1) / c.num_agents;
    p.num_tasks = p.G > 0 ? (c.num_envs + p.G - 1) / p.G : 0;
    p.B = c.boundary;
    p.sensor_range = c.sensor_range;
    p.cd = c.collision_distance;
    p.cd4 = (float)((double)c.collision_distance * 4.0);   // gym_flock_uw.py:197
    p.vmax = c.max_linear_velocity;
    p.noise_std = c.act_noise_std;
    p.dt = dt;
    p.range_lo = c.range_lo; p.reset_hi = c.reset_hi; p.heading_hi = c.heading_hi;
    p.reset_cd = c.reset_collision_distance;
    p.fill_hi = c.rigid_boundary ? c.boundary : 0.001f;
    p.fill_lo = c.rigid_boundary ? 0.0f : c.boundary;
    p.range_noise_std = c.range_noise_std;
    p.inv_n = (c.num_agents & (c.num_agents - 1)) == 0 ? 1.0f / (float)c.num_agents : 0.0f;
    p.seed_lo = (uint32_t)c.seed; p.seed_hi = (uint32_t)(c.seed >> 32);
    p.step_offset = 0;
    p.num_steps = 1;
    p.max_attempts = 1;
    p.x = b.x; p.y = b.y; p.h = b.h;       // the state is updated in place on both paths
    p.xo = b.x; p.yo = b.y; p.ho = b.h;
    p.prev_h = b.prev_h; p.vx = b.vx; p.vy = b.vy; p.obs = b.obs; p.nn = b.nn_idx;
    p.obs_head = c.obs_hist > 1 ? b.obs_head : nullptr;
    p.reward = b.reward; p.agent_done = b.agent_done; p.env_done = b.env_done;
    p.reset_epoch = b.reset_epoch;
    p.ep_return_fx = reinterpret_cast<long long*>(b.ep_return_fx);
    p.ep_len = b.ep_len;
    p.stats = reinterpret_cast<unsigned long long*>(b.stats);
    p.tile_scratch = e->tile_scratch;
    p.perm = nullptr;
    p.pair_counter = e->pair_counter;
    p.inv = e->tile_inv;
    p.sorted_xy = e->sorted_xy;
    p.hint_slots = e->hint_slots;
    p.env_sums = e->env_sums;
    p.far_rows = e->far_rows;
#ifdef FLOCK_TIMELINE
    p.timeline = e->timeline;
#endif
    return p;
}

int check_bound(const flock_env* e) {
    if (e == nullptr) return fail(FLOCK_E_INVALID, "null handle");
    if (!e->bound) return fail(FLOCK_E_UNBOUND, "flock_bind() has not been called");
    return FLOCK_OK;
}

int reset_device(flock_env* e, const uint8_t* env_mask, const float* init_state, int max_attempts, int flags,
                 cudaStream_t s) {
    Params p = make_params(e, 0.0f);
    p.env_mask = env_mask;
    p.init_state = init_state;
    p.max_attempts = max_attempts > 0 ? max_attempts : 64;
    p.reset_flags = flags;
    // reset installs the state into the CURRENT copy
    p.xo = const_cast<float*>(p.x); p.yo = const_cast<float*>(p.y); p.ho = const_cast<float*>(p.h);
    int n_launch = 1;
    cudaError_t err = e->path == 0 ? flock::launch_reset_small(p, e->sm_count, s)
                                   : flock::launch_reset_tiled(p, e->reset_need, &n_launch, s);
    e->launches += (uint64_t)n_launch;
    if (err != cudaSuccess) return cuda_fail(err, "reset kernel launch");
    if (env_mask == nullptr) e->perm_age = 0; // every env was redrawn: refresh the row order at the next step
    if (e->cfg.range_noise_std > 0.0f) {
        Params q = make_params(e, 0.0f);
        q.env_mask = env_mask;                // only the envs that were reset get a fresh noisy first observation
        err = flock::launch_range_noise(q, e->sm_count, s);
        e->launches += 1;
        if (err != cudaSuccess) return cuda_fail(err, "range noise kernel launch");
    }
    return FLOCK_OK;
}

struct HostMirrors {
    float* obs = nullptr;
    float* reward = nullptr;
    uint8_t* agent_done = nullptr;
    uint8_t* env_done = nullptr;
};

int step_device(flock_env* e, const float* actions, float dt, const float* noise, cudaStream_t s,
                const HostMirrors* mirrors = nullptr) {
    Params p = make_params(e, dt);
    p.actions = actions;
    p.noise = noise;
    if (mirrors != nullptr) {
        p.m_obs = mirrors->obs;
        p.m_reward = mirrors->reward;
        p.m_agent_done = mirrors->agent_done;
        p.m_env_done = mirrors->env_done;
    }
    cudaError_t err;
    const bool noisy = e->cfg.range_noise_std > 0.0f;
    const bool fused_reset = e->auto_reset && e->path == 0 && mirrors == nullptr && !noisy;
    // sensing noise rides in the step kernel's epilogue (small path, no host mirrors, no auto-reset);
    // otherwise it is the follow-up kernel of flock_small.cu
    const bool fused_noise = noisy && e->path == 0 && mirrors == nullptr && !e->auto_reset;
    if (e->path == 0) {
        int mode = flock::kSmallModeStep;
        if (mirrors != nullptr) mode = flock::kSmallModeMirror;
        else if (fused_reset) mode = flock::kSmallModeAutoReset;
        else if (fused_noise) mode = flock::kSmallModeNoise;
        if (fused_reset) {
            p.fused_auto_reset = 1;
            p.max_attempts = e->auto_reset_attempts;
        }
        err = flock::launch_step_small(e->cfg.variant, e->cfg.periodic != 0, p, mode, e->sm_count, s);
        e->launches += 1;
    } else {
        // keep the warps of the thread-per-row kernel spatially coherent: refresh the row order
        // every kPermRefreshSteps steps (and right after a reset); a stale order only costs speed
        constexpr uint32_t kPermRefreshSteps = 16;
        static const bool use_perm = [] {
            const char* v = getenv("FLOCK_ROW_ORDER");
            return v == nullptr || v[0] != '0';
        }();
        if (use_perm && e->tile_perm != nullptr && flock::tiled_uses_row_order(p, e->sm_count, e->tiled_mode)) {
            if (e->perm_age % kPermRefreshSteps == 0) {
                err = flock::launch_perm_refresh(p, e->tile_perm, e->tile_inv, s);
                e->launches += 1;
                if (err != cudaSuccess) return cuda_fail(err, "row order kernel launch");
            }
            e->perm_age += 1;
            p.perm = e->tile_perm;
        }
        err = flock::launch_step_tiled(e->cfg.variant, e->cfg.periodic != 0, p, e->sm_count, e->tiled_mode, s);
        e->launches += flock::tiled_step_launches(e->cfg.variant, p, e->sm_count, e->tiled_mode);
    }
    if (err != cudaSuccess) return cuda_fail(err, "step kernel launch");
    e->step_index += 1;
    if (noisy && !fused_noise) {             // optional sensing noise as a separate launch (see flock_small.cu)
        Params q = make_params(e, dt);        // reads the counters the step has just advanced
        err = flock::launch_range_noise(q, e->sm_count, s);
        e->launches += 1;
        if (err != cudaSuccess) return cuda_fail(err, "range noise kernel launch");
    }
    if (e->auto_reset && !fused_reset) {      // tiled path / host mirrors / sensing noise: second launch
        int rc = reset_device(e, e->b.env_done, nullptr, e->auto_reset_attempts, FLOCK_RESET_KEEP_OUTPUTS, s);
        if (rc != FLOCK_OK) return rc;
    }
    return FLOCK_OK;
}

int copy_results_to_host(flock_env* e, float* h_obs, float* h_reward, uint8_t* h_agent_done, uint8_t* h_env_done,
                         cudaStream_t s, bool sync = true) {
    const size_t EN = (size_t)e->cfg.num_envs * e->cfg.num_agents;
    cudaError_t err = cudaSuccess;
    // When the caller laid obs | reward | agent_done | env_done out back to back on BOTH sides (the
    // Python host does), the four results travel as one D2H copy instead of four.
    const size_t obs_bytes = EN * e->cfg.obs_hist * e->cfg.k * sizeof(float);
    const size_t rew_bytes = EN * sizeof(float);
    const bool ring = e->cfg.obs_hist > 1 && e->b.obs_head != nullptr;
    const float* d_obs = e->b.obs;
    if (ring && h_obs != nullptr) {   // the host wants the reference's newest-first window: materialise it first
        if (e->stage_window == nullptr && cudaMalloc(&e->stage_window, obs_bytes) != cudaSuccess)
            return cuda_fail(cudaGetLastError(), "window staging buffer");
        Params q = make_params(e, 0.0f);
        err = flock::launch_obs_window(q, e->stage_window, e->sm_count, s);
        e->launches += 1;
        if (err != cudaSuccess) return cuda_fail(err, "obs window kernel launch");
        d_obs = e->stage_window;
    }
    const char* d0 = reinterpret_cast<const char*>(e->b.obs);
    char* h0 = reinterpret_cast<char*>(h_obs);
    const bool packed = !ring && h_obs && h_reward && h_agent_done && h_env_done &&
                        reinterpret_cast<const char*>(e->b.reward) == d0 + obs_bytes &&
                        reinterpret_cast<const char*>(e->b.agent_done) == d0 + obs_bytes + rew_bytes &&
                        reinterpret_cast<const char*>(e->b.env_done) == d0 + obs_bytes + rew_bytes + EN &&
                        reinterpret_cast<char*>(h_reward) == h0 + obs_bytes &&
                        reinterpret_cast<char*>(h_agent_done) == h0 + obs_bytes + rew_bytes &&
                        reinterpret_cast<char*>(h_env_done) == h0 + obs_bytes + rew_bytes + EN;
    if (packed) {
        err = cudaMemcpyAsync(h_obs, e->b.obs, obs_bytes + rew_bytes + EN + (size_t)e->cfg.num_envs,
                              cudaMemcpyDeviceToHost, s);
    } else {
        if (h_obs != nullptr && err == cudaSuccess)
            err = cudaMemcpyAsync(h_obs, d_obs, obs_bytes, cudaMemcpyDeviceToHost, s);
        if (h_reward != nullptr && err == cudaSuccess)
            err = cudaMemcpyAsync(h_reward, e->b.reward, rew_bytes, cudaMemcpyDeviceToHost, s);
        if (h_agent_done != nullptr && err == cudaSuccess)
            err = cudaMemcpyAsync(h_agent_done, e->b.agent_done, EN, cudaMemcpyDeviceToHost, s);
        if (h_env_done != nullptr && err == cudaSuccess)
            err = cudaMemcpyAsync(h_env_done, e->b.env_done, (size_t)e->cfg.num_envs, cudaMemcpyDeviceToHost, s);
    }
    if (err != cudaSuccess) return cuda_fail(err, "D2H outputs");
    if (!sync) return FLOCK_OK;
    err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) return cuda_fail(err, "stream synchronize");
    return FLOCK_OK;
}

}  // namespace

extern "C" {

const char* flock_last_error(void) { return g_err; }
int flock_abi_version(void) { return FLOCK_ABI_VERSION; }

int flock_create(const flock_cfg_t* cfg, int device, flock_env_t** out) {
    if (cfg == nullptr || out == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    *out = nullptr;
    int rc = validate(*cfg);
    if (rc != FLOCK_OK) return rc;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return fail(FLOCK_E_NO_DEVICE, "no CUDA device (%s); libflock_b200 has no CPU fallback",
                    err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
    if (device < 0 || device >= count) return fail(FLOCK_E_INVALID, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    err = cudaGetDeviceProperties(&prop, device);
    if (err != cudaSuccess) return cuda_fail(err, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(FLOCK_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    err = cudaSetDevice(device);
    if (err != cudaSuccess) return cuda_fail(err, "cudaSetDevice");
    flock_env* e = new (std::nothrow) flock_env();
    if (e == nullptr) return fail(FLOCK_E_INVALID, "out of host memory");
    memset(e, 0, sizeof(*e));
    e->cfg = *cfg;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->path = cfg->num_agents <= 32 ? 0 : 1;
    const int aw = cfg->variant == FLOCK_UWD ? 1 : 2;
    e->action_floats = (size_t)cfg->num_envs * cfg->num_agents * aw;
    err = cudaMalloc(&e->stage_actions, e->action_floats * sizeof(float));
    if (err == cudaSuccess && cfg->variant == FLOCK_UWD)
        err = cudaMalloc(&e->stage_noise, (size_t)cfg->num_envs * cfg->num_agents * 2 * sizeof(float));
#ifdef FLOCK_TIMELINE
    if (err == cudaSuccess) {
        const size_t bytes = (size_t)flock::kTimelineSlots * flock::kTimelineCtas * sizeof(unsigned long long);
        err = cudaMalloc(&e->timeline, bytes);
        if (err == cudaSuccess) err = cudaMemset(e->timeline, 0, bytes);
    }
#endif
    if (err == cudaSuccess && e->path == 1) err = flock::tiled_configure(cfg->num_agents);
    if (err == cudaSuccess && e->path == 1) {
        err = cudaMalloc(&e->tile_scratch, (size_t)cfg->num_envs * 2 * sizeof(unsigned int));
        if (err == cudaSuccess) err = cudaMemset(e->tile_scratch, 0, (size_t)cfg->num_envs * 2 * sizeof(unsigned int));
        if (err == cudaSuccess) err = cudaMalloc(&e->reset_need, (size_t)cfg->num_envs);
        if (err == cudaSuccess) err = cudaMemset(e->reset_need, 0, (size_t)cfg->num_envs);
        if (err == cudaSuccess) err = cudaMalloc(&e->tile_perm, (size_t)cfg->num_envs * cfg->num_agents * sizeof(int));
        if (err == cudaSuccess) err = cudaMalloc(&e->tile_inv, (size_t)cfg->num_envs * cfg->num_agents * sizeof(int));
        if (err == cudaSuccess)
            err = cudaMalloc(&e->sorted_xy, flock::pruned_scratch_floats(cfg->num_agents, cfg->num_envs) * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&e->hint_slots, flock::pruned_hint_bytes(cfg->num_agents, cfg->num_envs));
        if (err == cudaSuccess)
            err = cudaMemset(e->hint_slots, 0xff, flock::pruned_hint_bytes(cfg->num_agents, cfg->num_envs));
        if (err == cudaSuccess && cfg->variant != FLOCK_V2)
            err = cudaMalloc(&e->env_sums, (size_t)cfg->num_envs * 2 * sizeof(float));
        if (err == cudaSuccess) err = cudaMalloc(&e->far_rows, flock::pruned_hint_bytes(cfg->num_agents, cfg->num_envs) / 16);
        if (err == cudaSuccess) err = cudaMemset(e->far_rows, 0, flock::pruned_hint_bytes(cfg->num_agents, cfg->num_envs) / 16);
        if (err == cudaSuccess)
            err = flock::launch_perm_identity(e->tile_perm, e->tile_inv, cfg->num_agents, cfg->num_envs, nullptr);
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
    }
    if (err != cudaSuccess) {
        cudaFree(e->tile_scratch);
        cudaFree(e->reset_need);
        cudaFree(e->tile_perm);
        cudaFree(e->tile_inv);
        cudaFree(e->sorted_xy);
        cudaFree(e->hint_slots);
        cudaFree(e->env_sums);
        cudaFree(e->far_rows);
        cudaFree(e->timeline);
        cudaFree(e->pair_counter);
        cudaFree(e->stage_actions);
        cudaFree(e->stage_noise);
        delete e;
        return cuda_fail(err, "flock_create");
    }
    *out = e;
    return FLOCK_OK;
}

void flock_destroy(flock_env_t* e) {
    if (e == nullptr) return;
    cudaFree(e->stage_actions);
    cudaFree(e->stage_noise);
    cudaFree(e->stage_window);
    cudaFree(e->tile_scratch);
    cudaFree(e->reset_need);
    cudaFree(e->tile_perm);
    cudaFree(e->tile_inv);
    cudaFree(e->sorted_xy);
    cudaFree(e->hint_slots);
    cudaFree(e->env_sums);
    cudaFree(e->far_rows);
    cudaFree(e->timeline);
    cudaFree(e->pair_counter);
    if (e->host_event_live) cudaEventDestroy(e->host_event);
    delete e;
}

int flock_bind(flock_env_t* e, const flock_buffers_t* b) {
    if (e == nullptr || b == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    if (!b->x || !b->y || !b->h || !b->prev_h || !b->obs || !b->reward || !b->agent_done || !b->env_done ||
        !b->reset_epoch || !b->ep_len)
        return fail(FLOCK_E_UNBOUND, "a required buffer pointer is NULL");
    e->b = *b;
    e->bound = true;
    return FLOCK_OK;
}

int flock_reset(flock_env_t* e, const uint8_t* env_mask, const float* init_state, int max_attempts, int flags,
                void* stream) {
    int rc = check_bound(e);
    if (rc != FLOCK_OK) return rc;
    return reset_device(e, env_mask, init_state, max_attempts, flags, static_cast<cudaStream_t>(stream));
}

int flock_set_auto_reset(flock_env_t* e, int enabled, int max_attempts) {
    if (e == nullptr) return fail(FLOCK_E_INVALID, "null handle");
    e->auto_reset = enabled != 0;
    e->auto_reset_attempts = max_attempts > 0 ? max_attempts : 64;
    return FLOCK_OK;
}

int flock_step(flock_env_t* e, const float* actions, float dt, const float* noise, void* stream) {
    int rc = check_bound(e);
    if (rc != FLOCK_OK) return rc;
    if (actions == nullptr) return fail(FLOCK_E_INVALID, "actions is NULL");
    // the kernels read (a0, a1) / (n_u, n_w) pairs as 8-byte vectors
    if ((reinterpret_cast<uintptr_t>(actions) & (e->cfg.variant == FLOCK_UWD ? 3u : 7u)) || (reinterpret_cast<uintptr_t>(noise) & 7u))
        return fail(FLOCK_E_INVALID, "actions / noise must be 8-byte aligned (4-byte for uw_discrete action ids)");
    return step_device(e, actions, dt, noise, static_cast<cudaStream_t>(stream));
}

int flock_random_actions(flock_env_t* e, uint32_t step_offset, float* actions, void* stream) {
    int rc0 = check_bound(e);
    if (rc0 != FLOCK_OK) return rc0;
    if (actions == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    Params p = make_params(e, 0.0f);
    p.step_offset = step_offset;
    cudaError_t err = flock::launch_random_actions(e->cfg.variant, p, actions, e->sm_count,
                                                   static_cast<cudaStream_t>(stream));
    e->launches += 1;
    if (err != cudaSuccess) return cuda_fail(err, "random_actions kernel launch");
    return FLOCK_OK;
}

int flock_step_n(flock_env_t* e, int num_steps, float dt, void* stream) {
    int rc = check_bound(e);
    if (rc != FLOCK_OK) return rc;
    if (num_steps < 1) return fail(FLOCK_E_INVALID, "num_steps must be >= 1");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (e->path == 0 && e->cfg.range_noise_std == 0.0f) {   // sensing noise needs the per-step follow-up kernel
        Params p = make_params(e, dt);
        p.num_steps = num_steps;
        cudaError_t err = flock::launch_step_small(e->cfg.variant, e->cfg.periodic != 0, p, flock::kSmallModeMulti, e->sm_count, s);
        e->launches += 1;
        if (err != cudaSuccess) return cuda_fail(err, "step_n kernel launch");
        e->step_index += (uint32_t)num_steps;
        return FLOCK_OK;
    }
    const int saved_auto_reset = e->auto_reset;   // step_n never restarts envs (same as the persistent kernel)
    e->auto_reset = 0;
    for (int t = 0; t < num_steps && rc == FLOCK_OK; ++t) {
        rc = flock_random_actions(e, 0u, e->stage_actions, stream);
        if (rc == FLOCK_OK) rc = step_device(e, e->stage_actions, dt, nullptr, s);
    }
    e->auto_reset = saved_auto_reset;
    return rc;
}

int flock_rollout_n(flock_env_t* e, int num_steps, const float* actions_T, float dt, float* obs_T, float* reward_T,
                    uint8_t* agent_done_T, uint8_t* env_done_T, int32_t* nn_T, void* stream) {
    int rc = check_bound(e);
    if (rc != FLOCK_OK) return rc;
    if (num_steps < 1) return fail(FLOCK_E_INVALID, "num_steps must be >= 1");
    if (obs_T == nullptr || reward_T == nullptr || agent_done_T == nullptr || env_done_T == nullptr)
        return fail(FLOCK_E_INVALID, "a trajectory buffer is NULL");
    if (nn_T != nullptr && e->b.nn_idx == nullptr) return fail(FLOCK_E_INVALID, "nn_T needs a bound nn_idx buffer");
    const bool uwd = e->cfg.variant == FLOCK_UWD;
    if ((reinterpret_cast<uintptr_t>(actions_T) & (uwd ? 3u : 7u)) || (reinterpret_cast<uintptr_t>(obs_T) & 15u) ||
        (reinterpret_cast<uintptr_t>(nn_T) & 15u) || (reinterpret_cast<uintptr_t>(reward_T) & 3u))
        return fail(FLOCK_E_INVALID, "trajectory buffers must be aligned (obs / nn 16 bytes, actions 8, reward 4)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t EN = (size_t)e->cfg.num_envs * e->cfg.num_agents, k = (size_t)e->cfg.k;
    if (e->path == 0 && e->cfg.range_noise_std == 0.0f) {
        Params p = make_params(e, dt);
        p.num_steps = num_steps;
        p.traj_actions = actions_T;
        p.traj_obs = obs_T;
        p.traj_reward = reward_T;
        p.traj_agent_done = agent_done_T;
        p.traj_env_done = env_done_T;
        p.traj_nn = nn_T;
        if (e->auto_reset) {
            p.fused_auto_reset = 1;
            p.max_attempts = e->auto_reset_attempts;
        }
        cudaError_t err = flock::launch_step_small(e->cfg.variant, e->cfg.periodic != 0, p, flock::kSmallModeRollout, e->sm_count, s);
        e->launches += 1;
        if (err != cudaSuccess) return cuda_fail(err, "rollout kernel launch");
        e->step_index += (uint32_t)num_steps;
        return FLOCK_OK;
    }
    // large swarms / sensing noise: one step at a time, results copied into the slices (device to device)
    const size_t aw = uwd ? 1 : 2, E = (size_t)e->cfg.num_envs;
    for (int t = 0; t < num_steps; ++t) {
        const float* act = actions_T != nullptr ? actions_T + (size_t)t * EN * aw : e->stage_actions;
        if (actions_T == nullptr) {
            rc = flock_random_actions(e, 0u, e->stage_actions, stream);
            if (rc != FLOCK_OK) return rc;
        }
        rc = step_device(e, act, dt, nullptr, s);
        if (rc != FLOCK_OK) return rc;
        Params q = make_params(e, dt);
        cudaError_t err = flock::launch_newest_row(q, obs_T + (size_t)t * EN * k, e->sm_count, s);
        e->launches += 1;
        if (err == cudaSuccess) err = cudaMemcpyAsync(reward_T + (size_t)t * EN, e->b.reward, EN * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (err == cudaSuccess) err = cudaMemcpyAsync(agent_done_T + (size_t)t * EN, e->b.agent_done, EN, cudaMemcpyDeviceToDevice, s);
        if (err == cudaSuccess) err = cudaMemcpyAsync(env_done_T + (size_t)t * E, e->b.env_done, E, cudaMemcpyDeviceToDevice, s);
        if (err == cudaSuccess && nn_T != nullptr)
            err = cudaMemcpyAsync(nn_T + (size_t)t * EN * k, e->b.nn_idx, EN * k * sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
        if (err != cudaSuccess) return cuda_fail(err, "rollout slice copy");
    }
    return FLOCK_OK;
}

int flock_obs_window(flock_env_t* e, float* out, void* stream) {
    int rc = check_bound(e);
    if (rc != FLOCK_OK) return rc;
    if (out == nullptr) return fail(FLOCK_E_INVALID, "out is NULL");
    if (e->cfg.obs_hist < 2 || e->b.obs_head == nullptr)
        return fail(FLOCK_E_INVALID, "flock_obs_window needs the uw ring layout (obs_head bound); the window layout IS the window");
    Params p = make_params(e, 0.0f);
    cudaError_t err = flock::launch_obs_window(p, out, e->sm_count, static_cast<cudaStream_t>(stream));
    e->launches += 1;
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "obs window kernel launch");
}

static int step_host_impl(flock_env_t* e, const float* h_actions, float dt, const float* h_noise, float* h_obs,
                          float* h_reward, uint8_t* h_agent_done, uint8_t* h_env_done, void* stream, bool sync) {
    int rc = check_bound(e);
    if (rc != FLOCK_OK) return rc;
    if (h_actions == nullptr) return fail(FLOCK_E_INVALID, "h_actions is NULL");
    if ((reinterpret_cast<uintptr_t>(h_actions) & (e->cfg.variant == FLOCK_UWD ? 3u : 7u)) || (reinterpret_cast<uintptr_t>(h_noise) & 7u))
        return fail(FLOCK_E_INVALID, "h_actions / h_noise must be 8-byte aligned (4-byte for uw_discrete action ids)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t EN = (size_t)e->cfg.num_envs * e->cfg.num_agents;
    // Zero-copy fast path: when every host buffer is pinned (device-visible under UVA) the kernel
    // reads the actions from and writes the results to host memory itself -- no staging copies, the
    // PCIe traffic overlaps the kernel, one synchronisation per step.
    if (h_obs && h_reward && h_agent_done && h_env_done && (h_noise == nullptr || e->cfg.variant != FLOCK_UWD)) {
        const void* hp[5] = {h_actions, h_obs, h_reward, h_agent_done, h_env_done};
        bool same = true;
        for (int i = 0; i < 5; ++i) same = same && (hp[i] == e->zc_host[i]);
        if (!same) {
            e->zc_ok = true;
            for (int i = 0; i < 5; ++i) {
                cudaPointerAttributes a;
                cudaError_t pe = cudaPointerGetAttributes(&a, hp[i]);
                if (pe != cudaSuccess || a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) {
                    cudaGetLastError();
                    e->zc_ok = false;
                    break;
                }
                e->zc_dev[i] = a.devicePointer;
            }
            for (int i = 0; i < 5; ++i) e->zc_host[i] = hp[i];
        }
        // Policy (measured on B200, profiles/README.md): SM-posted writes to host memory beat a DMA
        // copy only while the result set is small (latency bound); large result sets move faster by
        // DMA. The tiled path re-reads actions once per CTA, so it always stages.
        static const int zc_mode = [] {   // FLOCK_ZEROCOPY = 0 (never) | 1 (always) | unset (auto)
            const char* v = getenv("FLOCK_ZEROCOPY");
            return v == nullptr ? 2 : (v[0] == '0' ? 0 : 1);
        }();
        const size_t out_bytes = EN * e->cfg.obs_hist * e->cfg.k * sizeof(float) + EN * 5 + (size_t)e->cfg.num_envs;
        // async (pipelined) callers keep several steps in flight, which hides the fixed cost of the copy engines:
        // there both directions go by DMA (measured, cfg2, four batches in flight: 1.87-1.97e9 agent-steps/s,
        // against 1.78e9 with the zero-copy kernel and 1.65e9 with zero-copy inputs + DMA results);
        // FLOCK_ASYNC_ZC_INPUTS=1 lets the kernel read the actions from host memory itself
        static const bool async_zc_inputs = [] {
            const char* v = getenv("FLOCK_ASYNC_ZC_INPUTS");
            return v != nullptr && v[0] == '1';
        }();
        const bool zc_inputs = e->zc_ok && e->path == 0 && zc_mode != 0 && (sync || async_zc_inputs);
        const bool ring = e->cfg.obs_hist > 1 && e->b.obs_head != nullptr;   // the host mirrors hold the window, not the ring
        const bool zc_outputs = sync && zc_inputs && e->cfg.range_noise_std == 0.0f && !e->auto_reset && !ring &&   // post-step launches
                                (zc_mode == 1 || out_bytes <= (size_t)3 << 20);
        if (zc_outputs) {
            HostMirrors mir;
            mir.obs = static_cast<float*>(e->zc_dev[1]);
            mir.reward = static_cast<float*>(e->zc_dev[2]);
            mir.agent_done = static_cast<uint8_t*>(e->zc_dev[3]);
            mir.env_done = static_cast<uint8_t*>(e->zc_dev[4]);
            rc = step_device(e, static_cast<const float*>(e->zc_dev[0]), dt, nullptr, s, &mir);
            if (rc != FLOCK_OK) return rc;
            if (!sync) return FLOCK_OK;
            cudaError_t se = cudaStreamSynchronize(s);
            if (se != cudaSuccess) return cuda_fail(se, "stream synchronize");
            return FLOCK_OK;
        }
        if (zc_inputs) {   // actions read in place, results by one packed DMA copy below
            rc = step_device(e, static_cast<const float*>(e->zc_dev[0]), dt, nullptr, s);
            if (rc != FLOCK_OK) return rc;
            return copy_results_to_host(e, h_obs, h_reward, h_agent_done, h_env_done, s, sync);
        }
    }
    cudaError_t err = cudaMemcpyAsync(e->stage_actions, h_actions, e->action_floats * sizeof(float),
                                      cudaMemcpyHostToDevice, s);
    if (err != cudaSuccess) return cuda_fail(err, "H2D actions");
    const float* d_noise = nullptr;
    if (h_noise != nullptr && e->stage_noise != nullptr) {
        err = cudaMemcpyAsync(e->stage_noise, h_noise, EN * 2 * sizeof(float), cudaMemcpyHostToDevice, s);
        if (err != cudaSuccess) return cuda_fail(err, "H2D noise");
        d_noise = e->stage_noise;
    }
    rc = step_device(e, e->stage_actions, dt, d_noise, s);
    if (rc != FLOCK_OK) return rc;
    return copy_results_to_host(e, h_obs, h_reward, h_agent_done, h_env_done, s, sync);
}

int flock_step_host(flock_env_t* e, const float* h_actions, float dt, const float* h_noise, float* h_obs,
                    float* h_reward, uint8_t* h_agent_done, uint8_t* h_env_done, void* stream) {
    return step_host_impl(e, h_actions, dt, h_noise, h_obs, h_reward, h_agent_done, h_env_done, stream, true);
}

int flock_step_host_async(flock_env_t* e, const float* h_actions, float dt, const float* h_noise, float* h_obs,
                          float* h_reward, uint8_t* h_agent_done, uint8_t* h_env_done, void* stream) {
    // same transfer policy as flock_step_host (zero-copy kernel for small result sets, copy engines otherwise)
    int rc = step_host_impl(e, h_actions, dt, h_noise, h_obs, h_reward, h_agent_done, h_env_done, stream, false);
    if (rc != FLOCK_OK) return rc;
    cudaError_t err;
    if (!e->host_event_live) {
        err = cudaEventCreateWithFlags(&e->host_event, cudaEventDisableTiming);
        if (err != cudaSuccess) return cuda_fail(err, "event create");
        e->host_event_live = true;
    }
    err = cudaEventRecord(e->host_event, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "event record");
}

int flock_wait_host(flock_env_t* e) {
    if (e == nullptr) return fail(FLOCK_E_INVALID, "null handle");
    if (!e->host_event_live) return fail(FLOCK_E_INVALID, "flock_wait_host without a flock_step_host_async");
    cudaError_t err = cudaEventSynchronize(e->host_event);
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "event synchronize");
}

uint32_t flock_get_step_index(const flock_env_t* e) { return e ? e->step_index : 0u; }
int flock_set_step_index(flock_env_t* e, uint32_t step_index) {
    if (e == nullptr) return fail(FLOCK_E_INVALID, "null handle");
    e->step_index = step_index;
    return FLOCK_OK;
}
uint64_t flock_launch_count(const flock_env_t* e) { return e ? e->launches : 0ULL; }
int flock_path(const flock_env_t* e) { return e ? e->path : 0; }
uint64_t flock_pairs_evaluated(flock_env_t* e, int reset) {
    if (e == nullptr) return 0ULL;
    if (e->pair_counter == nullptr) {   // counting costs one atomic per warp: enabled by the first query
        if (cudaMalloc(&e->pair_counter, sizeof(unsigned long long)) != cudaSuccess) return 0ULL;
        cudaMemset(e->pair_counter, 0, sizeof(unsigned long long));
        return 0ULL;
    }
    unsigned long long v = 0;
    if (cudaMemcpy(&v, e->pair_counter, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return 0ULL;
    if (reset) cudaMemset(e->pair_counter, 0, sizeof(v));
    return v;
}
int flock_set_tiled_mode(flock_env_t* e, int mode) {
    if (e == nullptr) return fail(FLOCK_E_INVALID, "null handle");
    if (mode < 0 || mode > 2) return fail(FLOCK_E_INVALID, "tiled mode %d not in {0,1,2}", mode);
    e->tiled_mode = mode;
    return FLOCK_OK;
}

size_t flock_actor_packed_bytes(int num_agents) {
    return num_agents > 0 ? (size_t)num_agents * flock::actor_blob_bytes() : 0;
}

static int actor_check_dims(int num_agents, int in_dims, int fc1_dims, int fc2_dims, int n_actions) {
    int fc1, fc2, na;
    flock::actor_dims(&fc1, &fc2, &na);
    if (num_agents < 1 || num_agents > 65535) return fail(FLOCK_E_INVALID, "num_agents %d not in [1, 65535]", num_agents);
    if (in_dims < 1 || in_dims > flock::actor_max_in_dims())
        return fail(FLOCK_E_INVALID, "actor input_dims %d not in [1, %d]", in_dims, flock::actor_max_in_dims());
    if (fc1_dims != fc1 || fc2_dims != fc2 || n_actions != na)
        return fail(FLOCK_E_INVALID, "fused actor is built for %d-%d-%d (got %d-%d-%d)", fc1, fc2, na, fc1_dims, fc2_dims,
                    n_actions);
    return FLOCK_OK;
}

int flock_actor_pack(int num_agents, int in_dims, int fc1_dims, int fc2_dims, int n_actions, const float* const* params,
                     void* packed, void* stream) {
    int rc = actor_check_dims(num_agents, in_dims, fc1_dims, fc2_dims, n_actions);
    if (rc != FLOCK_OK) return rc;
    if (params == nullptr || packed == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    for (int i = 0; i < 10; ++i)
        if (params[i] == nullptr) return fail(FLOCK_E_INVALID, "actor parameter %d is NULL", i);
    cudaError_t err = flock::launch_actor_pack(num_agents, in_dims, params, packed, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "actor pack kernel launch");
}

static flock::NoiseCounters noise_counters(const flock_noise_counters_t* c) {
    flock::NoiseCounters n;
    if (c != nullptr) {
        n.env_step = c->env_step;
        n.env_epoch = c->env_epoch;
    }
    return n;
}

static int actor_forward_impl(const void* packed, const float* obs, float* actions, int num_envs, int num_agents, int in_dims,
                              float* ou_state, float theta, float mu, float sigma, float dt, uint64_t seed, uint32_t step,
                              int env_offset, const flock_noise_counters_t* counters, void* stream,
                              const int32_t* obs_head = nullptr, int ring_h = 0, int ring_k = 0) {
    int rc = actor_check_dims(num_agents, in_dims, 400, 300, 2);
    if (rc != FLOCK_OK) return rc;
    if (packed == nullptr || obs == nullptr || actions == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    if (num_envs < 1 || (num_envs + 127) / 128 > 65535) return fail(FLOCK_E_INVALID, "num_envs %d out of range", num_envs);
    if ((reinterpret_cast<uintptr_t>(packed) & 15u) || (reinterpret_cast<uintptr_t>(obs) & 15u) ||
        (reinterpret_cast<uintptr_t>(actions) & 7u) || (reinterpret_cast<uintptr_t>(ou_state) & 7u))
        return fail(FLOCK_E_INVALID, "actor buffers must be 16-byte aligned");
    if (obs_head != nullptr && (ring_h < 1 || ring_h > 4 || ring_k < 1 || ring_k > 3 || ring_h * ring_k != in_dims))
        return fail(FLOCK_E_INVALID, "observation ring %d x %d does not match input_dims %d (obs_hist <= 4, k <= 3)", ring_h, ring_k, in_dims);
    cudaError_t err = flock::launch_actor_forward(packed, obs, actions, num_envs, num_agents, in_dims, ou_state, theta, mu, sigma,
                                                  dt, seed, step, env_offset, noise_counters(counters), obs_head, ring_h, ring_k,
                                                  static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "actor forward kernel launch");
}

int flock_actor_forward(const void* packed, const float* obs, float* actions, int num_envs, int num_agents, int in_dims,
                        void* stream) {
    return actor_forward_impl(packed, obs, actions, num_envs, num_agents, in_dims, nullptr, 0.f, 0.f, 0.f, 0.f, 0ULL, 0u, 0, nullptr,
                              stream);
}

int flock_actor_forward_ou(const void* packed, const float* obs, float* actions, int num_envs, int num_agents, int in_dims,
                           float* ou_state, float theta, float mu, float sigma, float dt, uint64_t seed, uint32_t step,
                           int env_offset, const flock_noise_counters_t* counters, void* stream) {
    if (ou_state == nullptr) return fail(FLOCK_E_INVALID, "ou_state is NULL");
    if (!(dt >= 0.0f) || !(sigma >= 0.0f)) return fail(FLOCK_E_INVALID, "OU dt and sigma must be >= 0");
    return actor_forward_impl(packed, obs, actions, num_envs, num_agents, in_dims, ou_state, theta, mu, sigma, dt, seed, step,
                              env_offset, counters, stream);
}

int flock_actor_forward_ring(const void* packed, const float* obs_ring, const int32_t* obs_head, float* actions, int num_envs,
                             int num_agents, int obs_hist, int k, float* ou_state, float theta, float mu, float sigma, float dt,
                             uint64_t seed, uint32_t step, int env_offset, const flock_noise_counters_t* counters, void* stream) {
    if (obs_head == nullptr) return fail(FLOCK_E_INVALID, "obs_head is NULL");
    if (ou_state != nullptr && (!(dt >= 0.0f) || !(sigma >= 0.0f))) return fail(FLOCK_E_INVALID, "OU dt and sigma must be >= 0");
    return actor_forward_impl(packed, obs_ring, actions, num_envs, num_agents, obs_hist * k, ou_state, theta, mu, sigma, dt, seed,
                              step, env_offset, counters, stream, obs_head, obs_hist, k);
}

size_t flock_rnn_actor_packed_bytes(int num_agents) {
    return num_agents > 0 ? (size_t)num_agents * flock::rnn_actor_blob_bytes() : 0;
}

int flock_rnn_actor_pack(int num_agents, int hidden_rnn, int hidden1, int hidden2, int n_actions, const float* const* params,
                         void* packed, void* stream) {
    if (num_agents < 1 || num_agents > 65535) return fail(FLOCK_E_INVALID, "num_agents %d not in [1, 65535]", num_agents);
    if (hidden_rnn != 32 || hidden1 != 400 || hidden2 != 300 || n_actions != 2)
        return fail(FLOCK_E_INVALID, "fused recurrent actor is built for 32-400-300-2 (got %d-%d-%d-%d)", hidden_rnn, hidden1,
                    hidden2, n_actions);
    if (params == nullptr || packed == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    for (int i = 0; i < 8; ++i)
        if (params[i] == nullptr) return fail(FLOCK_E_INVALID, "recurrent actor parameter %d is NULL", i);
    cudaError_t err = flock::launch_rnn_actor_pack(num_agents, params, packed, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "rnn actor pack kernel launch");
}

static int rnn_actor_forward_impl(const void* packed, const float* const* front_params, const float* obs, const float* hidden_in,
                                  float* hidden_out, float* actions, int num_envs, int num_agents, int n_obs, float* ou_state,
                                  float theta, float mu, float sigma, float dt, uint64_t seed, uint32_t step, int env_offset,
                                  const flock_noise_counters_t* counters, void* stream) {
    if (packed == nullptr || front_params == nullptr || obs == nullptr || hidden_in == nullptr || hidden_out == nullptr ||
        actions == nullptr)
        return fail(FLOCK_E_INVALID, "null argument");
    if (num_agents < 1 || num_agents > 65535 || num_envs < 1) return fail(FLOCK_E_INVALID, "bad num_envs / num_agents");
    if (n_obs < 1 || n_obs > flock::rnn_actor_max_obs())
        return fail(FLOCK_E_INVALID, "n_obs %d not in [1, %d]", n_obs, flock::rnn_actor_max_obs());
    for (int i = 0; i < 6; ++i) {
        if (front_params[i] == nullptr) return fail(FLOCK_E_INVALID, "front-end parameter %d is NULL", i);
        if (reinterpret_cast<uintptr_t>(front_params[i]) & 15u)
            return fail(FLOCK_E_INVALID, "front-end parameter %d is not 16-byte aligned", i);
    }
    if ((reinterpret_cast<uintptr_t>(packed) & 15u) || (reinterpret_cast<uintptr_t>(hidden_in) & 15u) ||
        (reinterpret_cast<uintptr_t>(hidden_out) & 15u) || (reinterpret_cast<uintptr_t>(actions) & 7u) ||
        (reinterpret_cast<uintptr_t>(ou_state) & 7u))
        return fail(FLOCK_E_INVALID, "recurrent actor buffers must be 16-byte aligned");
    cudaError_t err = flock::launch_rnn_actor_forward(packed, front_params, obs, hidden_in, hidden_out, actions, num_envs,
                                                      num_agents, n_obs, ou_state, theta, mu, sigma, dt, seed, step, env_offset,
                                                      noise_counters(counters), static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "rnn actor kernel launch");
}

int flock_rnn_actor_forward(const void* packed, const float* const* front_params, const float* obs, const float* hidden_in,
                            float* hidden_out, float* actions, int num_envs, int num_agents, int n_obs, void* stream) {
    return rnn_actor_forward_impl(packed, front_params, obs, hidden_in, hidden_out, actions, num_envs, num_agents, n_obs, nullptr,
                                  0.f, 0.f, 0.f, 0.f, 0ULL, 0u, 0, nullptr, stream);
}

int flock_rnn_actor_forward_ou(const void* packed, const float* const* front_params, const float* obs, const float* hidden_in,
                               float* hidden_out, float* actions, int num_envs, int num_agents, int n_obs, float* ou_state,
                               float theta, float mu, float sigma, float dt, uint64_t seed, uint32_t step, int env_offset,
                               const flock_noise_counters_t* counters, void* stream) {
    if (ou_state == nullptr) return fail(FLOCK_E_INVALID, "ou_state is NULL");
    if (!(dt >= 0.0f) || !(sigma >= 0.0f)) return fail(FLOCK_E_INVALID, "OU dt and sigma must be >= 0");
    return rnn_actor_forward_impl(packed, front_params, obs, hidden_in, hidden_out, actions, num_envs, num_agents, n_obs, ou_state,
                                  theta, mu, sigma, dt, seed, step, env_offset, counters, stream);
}

int flock_qnet_forward(const float* const* params, int recurrent, const float* obs, const float* hidden_in, float* q_out,
                       float* hidden_out, float* actions, int num_envs, int num_agents, int n_obs, int n_actions,
                       float epsilon, uint64_t seed, uint32_t step, int env_offset, const flock_noise_counters_t* counters,
                       void* stream) {
    if (params == nullptr || obs == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    if (num_envs < 1 || num_agents < 1 || num_agents > 65535) return fail(FLOCK_E_INVALID, "bad num_envs / num_agents");
    if (n_obs < 1 || n_obs > flock::qnet_max_obs()) return fail(FLOCK_E_INVALID, "n_obs %d not in [1, %d]", n_obs, flock::qnet_max_obs());
    if (n_actions < 1 || n_actions > flock::qnet_max_actions())
        return fail(FLOCK_E_INVALID, "n_actions %d not in [1, %d]", n_actions, flock::qnet_max_actions());
    for (int i = 0; i < (recurrent ? 10 : 6); ++i) {
        if (params[i] == nullptr) return fail(FLOCK_E_INVALID, "Q-network parameter %d is NULL", i);
        if (reinterpret_cast<uintptr_t>(params[i]) & 15u)
            return fail(FLOCK_E_INVALID, "Q-network parameter %d is not 16-byte aligned", i);
    }
    if (recurrent && hidden_in == nullptr) return fail(FLOCK_E_INVALID, "recurrent Q-network needs hidden_in");
    if (recurrent && ((reinterpret_cast<uintptr_t>(hidden_in) & 15u) || (reinterpret_cast<uintptr_t>(hidden_out) & 15u)))
        return fail(FLOCK_E_INVALID, "hidden state buffers must be 16-byte aligned");
    if (!(epsilon >= 0.0f)) return fail(FLOCK_E_INVALID, "epsilon must be >= 0");
    cudaError_t err = flock::launch_qnet(params, recurrent, obs, hidden_in, q_out, hidden_out, actions, num_envs, num_agents,
                                         n_obs, n_actions, epsilon, seed, step, env_offset, noise_counters(counters),
                                         static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "qnet kernel launch");
}

size_t flock_gru_tc_packed_bytes(int mode, int num_agents) {
    return (num_agents > 0 && (mode == 0 || mode == 1)) ? (size_t)num_agents * flock::gru_tc_blob_bytes(mode) : 0;
}

int flock_gru_tc_pack(int mode, int num_agents, int n_obs, int n_actions, const float* const* params, void* packed, void* stream) {
    if (mode != 0 && mode != 1) return fail(FLOCK_E_INVALID, "mode %d not in {0 (recurrent actor front), 1 (VDN QNet)}", mode);
    if (num_agents < 1 || num_agents > 65535) return fail(FLOCK_E_INVALID, "num_agents %d not in [1, 65535]", num_agents);
    if (n_obs < 1 || n_obs > flock::gru_tc_max_obs()) return fail(FLOCK_E_INVALID, "n_obs %d not in [1, %d]", n_obs, flock::gru_tc_max_obs());
    if (mode == 1 && (n_actions < 1 || n_actions > flock::gru_tc_max_actions()))
        return fail(FLOCK_E_INVALID, "n_actions %d not in [1, %d]", n_actions, flock::gru_tc_max_actions());
    if (params == nullptr || packed == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    for (int i = 0; i < (mode == 1 ? 10 : 6); ++i)
        if (params[i] == nullptr) return fail(FLOCK_E_INVALID, "parameter %d is NULL", i);
    if (reinterpret_cast<uintptr_t>(packed) & 15u) return fail(FLOCK_E_INVALID, "packed must be 16-byte aligned");
    cudaError_t err = flock::launch_gru_tc_pack(mode, num_agents, n_obs, n_actions, params, packed, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "gru pack kernel launch");
}

static int gru_tc_check(const void* packed, const float* obs, const float* hidden_in, const float* hidden_out, int num_envs,
                        int num_agents, int n_obs) {
    if (packed == nullptr || obs == nullptr || hidden_in == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    if (num_envs < 1 || num_agents < 1 || num_agents > 65535) return fail(FLOCK_E_INVALID, "bad num_envs / num_agents");
    if (n_obs < 1 || n_obs > flock::gru_tc_max_obs()) return fail(FLOCK_E_INVALID, "n_obs %d not in [1, %d]", n_obs, flock::gru_tc_max_obs());
    if ((reinterpret_cast<uintptr_t>(packed) & 15u) || (reinterpret_cast<uintptr_t>(hidden_in) & 15u) ||
        (reinterpret_cast<uintptr_t>(hidden_out) & 15u) || (reinterpret_cast<uintptr_t>(obs) & 3u))
        return fail(FLOCK_E_INVALID, "packed / hidden state buffers must be 16-byte aligned");
    return FLOCK_OK;
}

int flock_qnet_forward_tc(const void* packed, const float* obs, const float* hidden_in, float* q_out, float* hidden_out,
                          float* actions, int num_envs, int num_agents, int n_obs, int n_actions, float epsilon, uint64_t seed,
                          uint32_t step, int env_offset, const flock_noise_counters_t* counters, void* stream) {
    int rc = gru_tc_check(packed, obs, hidden_in, hidden_out, num_envs, num_agents, n_obs);
    if (rc != FLOCK_OK) return rc;
    if (n_actions < 1 || n_actions > flock::gru_tc_max_actions())
        return fail(FLOCK_E_INVALID, "n_actions %d not in [1, %d]", n_actions, flock::gru_tc_max_actions());
    if (!(epsilon >= 0.0f)) return fail(FLOCK_E_INVALID, "epsilon must be >= 0");
    cudaError_t err = flock::launch_gru_tc_forward(1, packed, obs, hidden_in, hidden_out, q_out, actions, num_envs, num_agents, n_obs,
                                                   n_actions, epsilon, seed, step, env_offset, noise_counters(counters),
                                                   static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "qnet tensor-core kernel launch");
}

int flock_rnn_actor_forward_tc(const void* packed_mlp, const void* packed_front, const float* obs, const float* hidden_in,
                               float* hidden_out, float* actions, int num_envs, int num_agents, int n_obs, float* ou_state,
                               float theta, float mu, float sigma, float dt, uint64_t seed, uint32_t step, int env_offset,
                               const flock_noise_counters_t* counters, void* stream) {
    int rc = gru_tc_check(packed_front, obs, hidden_in, hidden_out, num_envs, num_agents, n_obs);
    if (rc != FLOCK_OK) return rc;
    if (packed_mlp == nullptr || hidden_out == nullptr || actions == nullptr) return fail(FLOCK_E_INVALID, "null argument");
    if ((reinterpret_cast<uintptr_t>(packed_mlp) & 15u) || (reinterpret_cast<uintptr_t>(actions) & 7u) ||
        (reinterpret_cast<uintptr_t>(ou_state) & 7u))
        return fail(FLOCK_E_INVALID, "recurrent actor buffers must be 16-byte aligned");
    if (ou_state != nullptr && (!(dt >= 0.0f) || !(sigma >= 0.0f))) return fail(FLOCK_E_INVALID, "OU dt and sigma must be >= 0");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t err = flock::launch_gru_tc_forward(0, packed_front, obs, hidden_in, hidden_out, nullptr, nullptr, num_envs, num_agents,
                                                   n_obs, 0, 0.0f, 0ULL, 0u, 0, flock::NoiseCounters(), s);
    if (err != cudaSuccess) return cuda_fail(err, "rnn front tensor-core kernel launch");
    err = flock::launch_rnn_actor_forward(packed_mlp, nullptr, obs, hidden_out, hidden_out, actions, num_envs, num_agents, n_obs,
                                          ou_state, theta, mu, sigma, dt, seed, step, env_offset, noise_counters(counters), s);
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "rnn actor MLP kernel launch");
}

#ifdef FLOCK_TIMELINE
// developer build only (tools/cta_timeline.py): copies the first n CTA records of the handle's timeline out and clears it
extern "C" __attribute__((visibility("default"))) int flock_debug_timeline(flock_env* e, unsigned long long* out, int n) {
    if (e == nullptr || e->timeline == nullptr || out == nullptr || n < 1 || n > flock::kTimelineCtas) return FLOCK_E_INVALID;
    const size_t bytes = (size_t)flock::kTimelineSlots * flock::kTimelineCtas * sizeof(unsigned long long);
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess)
        err = cudaMemcpy(out, e->timeline, (size_t)n * flock::kTimelineSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (err == cudaSuccess) err = cudaMemset(e->timeline, 0, bytes);
    return err == cudaSuccess ? FLOCK_OK : FLOCK_E_CUDA;
}
#endif

int flock_debug_sincos(const float* h, int n, float* sn, float* cs, void* stream) {
    cudaError_t err = flock::launch_debug_sincos(h, n, sn, cs, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "debug_sincos");
}
int flock_debug_normal2(const uint32_t* words, int n_pairs, float* z, void* stream) {
    cudaError_t err = flock::launch_debug_normal2(words, n_pairs, z, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "debug_normal2");
}
int flock_debug_philox(const uint32_t* ctr4_key2, int n, uint32_t* out4, void* stream) {
    cudaError_t err = flock::launch_debug_philox(ctr4_key2, n, out4, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? FLOCK_OK : cuda_fail(err, "debug_philox");
}

}  // extern "C"
