// flock_device.cuh -- device-side building blocks shared by the sm_100a flocking kernels.
//
// Canonical arithmetic (DESIGN.md section "Canonical arithmetic"): every parity-critical
// expression is a single IEEE binary32 operation; the translation units are compiled with
// -fmad=false so nvcc never contracts a*b+c, and fused operations appear only as explicit
// fmaf()/fma() calls that the CPU oracle restates one for one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flock_b200.h"

namespace flock {

constexpr int kWarp = 32;
#define kInf (__int_as_float(0x7f800000))
constexpr float kFltMax = 3.4028234663852886e38f;

enum : uint32_t { kTagReset = 0u, kTagNoise = 1u, kTagAction = 2u, kTagRange = 3u };

// Kernel parameters (passed by value as a __grid_constant__).
struct Params {
    int E, N, k, H;
    int env_offset;
    int G;        // envs per warp (small path)
    int g_magic;  // ceil(65536 / N): lane / N == (lane * g_magic) >> 16 for lane < 32
    int num_tasks;  // ceil(E / G)
    int sstride;  // shared-memory stride of one env group, floats (small path)
    float B, sensor_range, cd, cd4, vmax, noise_std, dt;
    float range_lo, reset_hi, heading_hi, reset_cd;
    float fill_hi, fill_lo;   // check_boundary replacement values (see wrap_coord)
    float range_noise_std;
    float inv_n;      // 1/N when N is a power of two (x * inv_n == x / N bit for bit), else 0
    uint32_t seed_lo, seed_hi, step_offset;   // step_offset: flock_random_actions look-ahead
    int num_steps;      // step_n
    int max_attempts;   // reset
    int reset_flags;    // reset: FLOCK_RESET_*
    int fused_auto_reset;   // step (small path): restart finished envs in the same launch
    const float* actions;
    const float* noise;
    const float* init_state;
    const uint8_t* env_mask;
    const float *x, *y, *h;   // state read this step
    float *xo, *yo, *ho;      // state written this step (== x,y,h on the small path)
    float* prev_h;
    float *vx, *vy;
    float* obs;
    int32_t* obs_head;  // uw ring layout: [E] slot of the newest row (nullptr = materialised window layout)
    int32_t* nn;
    float* reward;
    uint8_t* agent_done;
    uint8_t* env_done;
    uint32_t* reset_epoch;
    long long* ep_return_fx;
    int32_t* ep_len;
    unsigned long long* stats;
    unsigned int* tile_scratch;   // tiled path: [E] CTA arrival counters, [E] collision counters (self-resetting)
    uint8_t* reset_need;          // tiled reset: [E] envs that still collide after the attempt launches so far
    const int* perm;              // tiled thread-per-row path: [E][N] row order (spatially sorted), nullable = identity
    unsigned long long* pair_counter;   // pruned kernel: row x neighbour pairs actually evaluated (nullable)
    const int* inv;               // inverse of perm (agent -> slot)
    float* sorted_xy;             // pruned path: per env x | y by slot, boxes, agent ids (3 * PS floats)
    unsigned short* hint_slots;   // pruned path: [E][PS][8] slots of last step's neighbours (0xffff: none)
    uint8_t* far_rows;            // pruned path: [E][PS] rows (by slot) that wrapped around the world since the last row-order refresh
#ifdef FLOCK_TIMELINE
    unsigned long long* timeline; // developer build: per-CTA timeline records (see below)
#endif
    float* env_sums;              // tiled path, uw / uwd: [E][2] sequential (agent order) sums of the NEW x, y (uw) or h (uwd)
    // host-call path: device-visible HOST mirrors of the step results (zero-copy), nullable
    float* m_obs;
    float* m_reward;
    uint8_t* m_agent_done;
    uint8_t* m_env_done;
    // streamed rollout (flock_rollout_n): time-major trajectory buffers, slice t = step t of the launch
    const float* traj_actions;   // [T][E][N][2] or [T][E][N] (uwd); nullptr = canonical Philox random actions
    float* traj_obs;             // [T][E][N][k] newest range row of every step
    float* traj_reward;          // [T][E][N]
    uint8_t* traj_agent_done;    // [T][E][N]
    uint8_t* traj_env_done;      // [T][E]
    int32_t* traj_nn;            // [T][E][N][k], nullable
};

// uw observation history (gym_flock_uw.py:120-123), two layouts:
//   window (obs_head == nullptr): obs[E][N][H][k], newest row first, shifted by one row every step -- the
//          reference's materialised view (36 B read + 48 B written per agent-step at H = 4, k = 3);
//   ring:  obs[E][H][N][k] + obs_head[E]: row r of the window lives in slot (head + r) % H. A step moves the head back
//          by one and writes ONLY the new row: k floats per agent, one contiguous run per env, nothing read.
__device__ __forceinline__ int ring_prev_slot(int head, int H) { return head == 0 ? H - 1 : head - 1; }
__device__ __forceinline__ float* ring_row(const Params& p, int env, int slot, int a) {
    return p.obs + (((size_t)env * p.H + slot) * p.N + a) * (size_t)p.k;
}

// Programmatic dependent launch (PDL): a step kernel lets the NEXT kernel of the stream start
// launching immediately (its index prologue overlaps our execution) and itself waits for the
// previous kernel's memory before the first global access. No-ops when the launch does not carry
// the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }


// Developer instrumentation (compiled in only with -DFLOCK_TIMELINE, see tools/cta_timeline.py): per-CTA wall-clock
// timeline of a kernel -- %globaltimer at entry / after griddepcontrol.wait / at exit, SM id, and a few per-kernel
// counters -- written to Params::timeline (allocated by flock_create in such a build) and read back through
// flock_debug_timeline().
#ifdef FLOCK_TIMELINE
constexpr int kTimelineSlots = 8, kTimelineCtas = 65536;      // Params::timeline: [kTimelineCtas][kTimelineSlots]
__device__ __forceinline__ unsigned long long timeline_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned timeline_smid() {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}
#endif

// ---------------------------------------------------------------------------------------------
// Canonical transcendental functions (same DEFINITION as oracle/flock_oracle.c, separate code).
// ---------------------------------------------------------------------------------------------

// sin/cos of a binary32 angle: fp64 Cody-Waite reduction to |r| <= pi/4, fp32 polynomials with
// explicit fmaf. Replaces torch.cos / torch.sin of gym_flock_v2.py:335-336 and
// gym_flock_uw_discrete.py:351-352 (<= 2 ulp from the correctly rounded value).
__device__ __forceinline__ void sincos_canon(float h, float& sn, float& cs) {
    const double hd = (double)h;
    if (!(fabs(hd) < 1.0e9)) {
        sn = __int_as_float(0x7fc00000);
        cs = sn;
        return;
    }
    const double q = rint(hd * 0.63661977236758134308);
    double r = fma(-q, 1.57079632673412561417, hd);
    r = fma(-q, 6.07710050650619224932e-11, r);
    const float x = (float)r;
    const float z = x * x;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    const float xz = x * z;
    const float s = fmaf(xz, ps, x);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    const float zz = z * z;
    const float c = fmaf(zz, pc, fmaf(z, -0.5f, 1.0f));
    const int quad = (int)(__double2ll_rn(q) & 3LL);
    const bool swap = quad & 1;
    float so = swap ? c : s;
    float co = swap ? s : c;
    so = (quad & 2) ? -so : so;               // quadrants 2,3 negate sin
    co = ((quad + 1) & 2) ? -co : co;         // quadrants 1,2 negate cos
    sn = so;
    cs = co;
}

// ln(m * 2^-24), m integer in [1, 2^24], binary32 with explicit fmaf (Box-Muller radius; ~1e-7
// absolute error, irrelevant for a noise source but bit-identical to the oracle's restatement):
// u = f * 2^e with f in [sqrt(1/2), sqrt(2)), s = (f-1)/(f+1), ln f = 2s(1 + s^2/3 + ... + s^8/9).
__device__ __forceinline__ float log_u24(uint32_t m) {
    const float u = (float)m * (1.0f / 16777216.0f);
    uint32_t bits = __float_as_uint(u);
    int e = (int)((bits >> 23) & 0xffu) - 126;
    bits = (bits & 0x007fffffu) | 0x3f000000u;
    float f = __uint_as_float(bits);                 // [0.5, 1)
    if (f < 0.70710678118654752440f) {
        f = f * 2.0f;
        e -= 1;
    }
    const float s = __fdiv_rn(f - 1.0f, f + 1.0f);
    const float s2 = s * s;
    float p = 1.0f / 9.0f;
    p = fmaf(p, s2, 1.0f / 7.0f);
    p = fmaf(p, s2, 1.0f / 5.0f);
    p = fmaf(p, s2, 1.0f / 3.0f);
    p = fmaf(p, s2, 1.0f);
    const float two_s = 2.0f * s;
    const float lnf = two_s * p;
    return fmaf((float)e, 0.69314718055994530942f, lnf);
}

// Philox4x32-10, counter-based (Salmon et al. SC'11).
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u24(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// two standard normals from two 32-bit words (Box-Muller on the canonical log / sincos)
__device__ __forceinline__ void normal2(uint32_t r0, uint32_t r1, float& z0, float& z1) {
    const float ln = log_u24((r0 >> 8) + 1u);          // u1 in (0, 1]  =>  ln <= 0
    const float rad = __fsqrt_rn(-2.0f * ln);
    const float theta = 6.28318530717958647692f * u24(r1);
    float sn, cs;
    sincos_canon(theta, sn, cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

// ---------------------------------------------------------------------------------------------
// Reference semantics, per agent.
// ---------------------------------------------------------------------------------------------

// torch.clamp propagates NaN (gym_flock_v2.py:327,331)
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
    const float t = fminf(fmaxf(v, lo), hi);
    return (v != v) ? v : t;
}

// torch.nan_to_num defaults (gym_flock_v2.py:346)
__device__ __forceinline__ float nan_to_num(float v) {
    const float t = fminf(fmaxf(v, -kFltMax), kFltMax);   // fmaxf(NaN, a) = a: fixed below
    return (v != v) ? 0.0f : t;
}

// check_boundary per coordinate (gym_flock_v2.py:271-304): both modes are
//   c = (c < B) ? c : fill_hi;  c = (c > 0) ? c : fill_lo
// with (fill_hi, fill_lo) = (0.001, B) for the default wrap (:292-304) and (B, 0) for the rigid
// boundary (:273-289); the host passes the two fill values, so the kernel has no mode branch.
__device__ __forceinline__ float wrap_coord(float c, float B, float fill_hi, float fill_lo) {
    c = (c < B) ? c : fill_hi;
    c = (c > 0.0f) ? c : fill_lo;
    return c;
}

// _updateState + check_boundary for one agent. V: 0 v2 (gym_flock_v2.py:317-350), 1 uw
// (gym_flock_uw.py:269-302, heading=False), 2 uwd (gym_flock_uw_discrete.py:324-366).
// a0/a1: the agent's action (uwd: a0 = float-coded id); nzu/nzw: additive actuation noise.
template <int V>
__device__ __forceinline__ void integrate_agent(const Params& p, float a0, float a1, float nzu, float nzw,
                                                float& x, float& y, float& h, float& vx, float& vy) {
    const float dt = p.dt;
    if (V == FLOCK_V2) {
        const float w = clamp_nan(a1, -1.57079632679489661923f, 1.57079632679489661923f);
        const float wd = w * dt;
        h = h + wd;
        const float u = clamp_nan(a0, 0.005f, p.vmax);
        float sn, cs;
        sincos_canon(h, sn, cs);
        vx = u * cs;
        vy = u * sn;
    } else if (V == FLOCK_UW) {
        float n2 = a0 * a0;
        n2 = fmaf(a1, a1, n2);                            // torch.norm (gym_flock_uw.py:294)
        const float n = __fsqrt_rn(n2);
        vx = __fdiv_rn(a0, n);
        vy = __fdiv_rn(a1, n);
    } else {
        float a = (a0 == a0) ? a0 : 0.0f;                 // int(act) with a clamp to the dictionary
        a = fminf(fmaxf(a, 0.0f), 9.0f);
        const int id = (int)a;
        const float mu_u = id < 5 ? 0.2f : 0.6f;
        const int m = id < 5 ? id : id - 5;
        const float mu_w = m == 0 ? -1.2f : m == 1 ? -0.5f : m == 2 ? 0.0f : m == 3 ? 0.5f : 1.2f;
        float u = mu_u + nzu;
        float w = mu_w + nzw;
        w = clamp_nan(w, -0.025f, 0.025f);
        const float wd = w * dt;
        h = h + wd;
        u = clamp_nan(u, 5e-6f, p.vmax);
        float sn, cs;
        sincos_canon(h, sn, cs);
        vx = u * cs;
        vy = u * sn;
        float n2 = vx * vx;
        n2 = fmaf(vy, vy, n2);                            // torch.norm (gym_flock_uw_discrete.py:358)
        const float n = __fsqrt_rn(n2);
        vx = __fdiv_rn(vx, n);
        vy = __fdiv_rn(vy, n);
    }
    vx = nan_to_num(vx);
    vy = nan_to_num(vy);
    vx = vx * dt;
    vy = vy * dt;
    x = x + vx;
    y = y + vy;
    x = wrap_coord(x, p.B, p.fill_hi, p.fill_lo);
    y = wrap_coord(y, p.B, p.fill_hi, p.fill_lo);
}

// Philox stream layout: key = seed; counter = (global env, agent, episode step, tag + 4*reset_epoch).
// (episode step, reset_epoch) are per-env DEVICE counters, so the stream is unique over the env's
// life, independent of the GPU count and safe to replay from a CUDA graph.
__device__ __forceinline__ uint32_t stream_word(uint32_t tag, uint32_t reset_epoch) { return tag + (reset_epoch << 2); }

// actuation noise of uwd (gym_flock_uw_discrete.py:333-334): 2 normals * std
__device__ __forceinline__ void act_noise(const Params& p, int env_global, int agent, uint32_t step, uint32_t repoch,
                                          float& nzu, float& nzw) {
    const uint4 r = philox4x32_10((uint32_t)env_global, (uint32_t)agent, step, stream_word(kTagNoise, repoch),
                                  p.seed_lo, p.seed_hi);
    float z0, z1;
    normal2(r.x, r.y, z0, z1);
    nzu = p.noise_std * z0;
    nzw = p.noise_std * z1;
}

// canonical random action of (env, agent, episode step): tag 2
template <int V>
__device__ __forceinline__ void random_action(const Params& p, int env_global, int agent, uint32_t step,
                                              uint32_t repoch, float& a0, float& a1) {
    const uint4 r = philox4x32_10((uint32_t)env_global, (uint32_t)agent, step, stream_word(kTagAction, repoch),
                                  p.seed_lo, p.seed_hi);
    if (V == FLOCK_V2) {
        const float t0 = u24(r.x) * 3.0f, t1 = u24(r.y) * 3.0f;
        a0 = t0 - 1.5f;
        a1 = t1 - 1.5f;
    } else if (V == FLOCK_UW) {
        const float t0 = u24(r.x) * 2.0f, t1 = u24(r.y) * 2.0f;
        a0 = t0 - 1.0f;
        a1 = t1 - 1.0f;
    } else {
        a0 = (float)__umulhi(r.x, (uint32_t)p.k);
        a1 = 0.0f;
    }
}

// squared pair distance: min-image (gym_flock_v2.py:140-144) or Euclidean (v2:166-168).
// For wrapped coordinates |dx| is in [0, B], where (dx > B/2 ? B - dx : dx) == fminf(dx, B - dx)
// bit for bit (B/2 is exact).
template <bool PER>
__device__ __forceinline__ float pair_d2(float xi, float yi, float xj, float yj, float B) {
    float dx = fabsf(xi - xj);
    float dy = fabsf(yi - yj);
    if (PER) {
        dx = fminf(dx, B - dx);
        dy = fminf(dy, B - dy);
    }
    const float a = dx * dx;
    // Euclidean sites are torch.norm upstream, which accumulates the squares like an fma
    // (fma(dy, dy, dx*dx): with this form the uw oracle reproduces the reference bit for bit);
    // the periodic metric is an explicit multiply / add chain (gym_flock_v2.py:144): unfused.
    if (!PER) return fmaf(dy, dy, a);
    const float b = dy * dy;
    return a + b;
}

// Per-row running k-smallest list, ascending by (d2, j). Candidates must be offered in
// ascending j, so a strict `<` keeps the lower index on equal d2.
template <int K>
struct TopK {
    float d[K];
    int idx[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            d[s] = kInf;
            idx[s] = -1;
        }
    }
    // branch-free sorted insertion: values by a min/max network, indices by selects
    __device__ __forceinline__ void insert(float c, int j) {
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            const bool lt_prev = c < d[s - 1];
            const bool lt_cur = c < d[s];
            idx[s] = lt_prev ? idx[s - 1] : (lt_cur ? j : idx[s]);
            d[s] = fminf(d[s], fmaxf(d[s - 1], c));
        }
        idx[0] = (c < d[0]) ? j : idx[0];
        d[0] = fminf(d[0], c);
    }
    // insertion by the full (d2, j) key, for candidates that do NOT arrive in ascending j: d2 >= 0
    // (never NaN), so (bits(d2) << 32 | j) orders like (d2, j); empty slots hold (+inf, 0xffffffff)
    __device__ __forceinline__ void insert_lex(float c, int j) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(c) << 32) | (unsigned)j;
        bool lt[K];
#pragma unroll
        for (int s = 0; s < K; ++s)
            lt[s] = key < (((unsigned long long)__float_as_uint(d[s]) << 32) | (unsigned)idx[s]);
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            idx[s] = lt[s - 1] ? idx[s - 1] : (lt[s] ? j : idx[s]);
            d[s] = lt[s - 1] ? d[s - 1] : (lt[s] ? c : d[s]);
        }
        idx[0] = lt[0] ? j : idx[0];
        d[0] = lt[0] ? c : d[0];
    }
    __device__ __forceinline__ float worst() const { return d[K - 1]; }
};

// Values-only variant: the k smallest d2 WITH multiplicity (a min/max network, 2 FMNMX per slot), no
// index tracking. The reference discards the neighbour indices in uw / uwd
// (gym_flock_uw.py:141-144, gym_flock_uw_discrete.py:189-192), and the ranges -- hence obs,
// collisions, dones, rewards -- are identical to the indexed list.
template <int K>
struct TopKValues {
    float d[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < K; ++s) d[s] = kInf;
    }
    __device__ __forceinline__ void insert(float c, int) {
#pragma unroll
        for (int s = K - 1; s > 0; --s) d[s] = fminf(d[s], fmaxf(d[s - 1], c));
        d[0] = fminf(d[0], c);
    }
    __device__ __forceinline__ float worst() const { return d[K - 1]; }
};

// sqrt + clamp(0, sensor_range) of the k winners (gym_flock_v2.py:151) and the collision flag
// (gym_flock_v2.py:212-215, 314).
template <int K, typename List>
__device__ __forceinline__ bool finish_row(const List& t, int k, float sensor_range, float cd, float* dist) {
    bool coll = false;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        float d = __fsqrt_rn(t.d[s]);
        d = fminf(fmaxf(d, 0.0f), sensor_range);
        dist[s] = d;
        coll = coll || (s < k && d < cd);
    }
    return coll;
}

// Reward as a function of the agent's threshold flags, so that the per-agent value and the
// per-class constants used by the episode-return sum come from the same float operations.
//   v2  (gym_flock_v2.py:217-220,268):           coll ? -5 : 0.01
//   uw  (gym_flock_uw.py:186-221):  f1 = near the centre of mass, f2 = heading changed > 0.27
//   uwd (gym_flock_uw_discrete.py:234-276):      f1 = heading farther than 0.20 from the mean
template <int V>
__device__ __forceinline__ float reward_from_flags(bool coll, bool f1, bool f2) {
    if (V == FLOCK_V2) {
        return coll ? -5.0f : 0.01f;
    } else if (V == FLOCK_UW) {
        const float pen = coll ? -5.0f : 0.01f;
        const float rcom = f1 ? 0.01f : 0.0f;
        const float rang = f2 ? -0.01f : 0.001f;
        float r = pen + rcom;
        r = r + rang;
        return r;
    } else {
        const float pen = coll ? -9.0f : 0.0f;
        const float ral = f1 ? 0.0f : 0.1f;
        return pen + ral;
    }
}

// threshold flags of one agent; prev_h is the value BEFORE this step's update
template <int V>
__device__ __forceinline__ void reward_flags(const Params& p, float x, float y, float h, float prev_h, float comx,
                                             float comy, float hmean, bool& f1, bool& f2) {
    f1 = false;
    f2 = false;
    if (V == FLOCK_UW) {
        const float ddx = x - comx, ddy = y - comy;
        float q = ddx * ddx;
        q = fmaf(ddy, ddy, q);                            // torch.norm (gym_flock_uw.py:194)
        const float dc = __fsqrt_rn(q);
        f1 = dc < p.cd4;
        f2 = fabsf(prev_h - h) > 0.27f;
    } else if (V == FLOCK_UWD) {
        f1 = fabsf(hmean - h) > 0.20f;
    }
}

template <int V>
__device__ __forceinline__ float agent_reward(const Params& p, bool coll, float x, float y, float h,
                                              float prev_h, float comx, float comy, float hmean) {
    bool f1, f2;
    reward_flags<V>(p, x, y, h, prev_h, comx, comy, hmean, f1, f2);
    return reward_from_flags<V>(coll, f1, f2);
}

// mean over the env's agents: sum / N (torch.mean, gym_flock_uw.py:193; sum/N, uwd:256). For a
// power-of-two N the division is an exact scaling, so the multiply by 1/N gives the identical
// correctly rounded result without the IEEE division sequence.
__device__ __forceinline__ float mean_of_sum(const Params& p, float sum) {
    return p.inv_n != 0.0f ? sum * p.inv_n : __fdiv_rn(sum, (float)p.N);
}

// reward in 2^-32 fixed point (order-free integer accumulation of episode returns)
__device__ __forceinline__ long long reward_fx(float r) { return __double2ll_rn((double)r * 4294967296.0); }

}  // namespace flock
