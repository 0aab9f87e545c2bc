/*
 * flock_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * A plain-C restatement of the environment step / reset / obs / reward path of
 *   /root/reference/environments/gym_flock_v2.py           (variant 0, "v2")
 *   /root/reference/environments/gym_flock_uw.py           (variant 1, "uw")
 *   /root/reference/environments/gym_flock_uw_discrete.py  (variant 2, "uwd")
 * batched over independent env instances. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.
 *
 * PARITY STATUS: the reference ships no golden vectors or numeric tests for this path
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the reference ITSELF,
 * executed unmodified on CPU under oracle/ref_shim.py; the recorded vectors live in
 * tests/golden/ (*.npz; generator: tests/golden/make_golden.py) and are checked by
 * tests/test_oracle_golden.py.
 *
 * Arithmetic: every expression is evaluated in IEEE binary32 exactly as written (compile with
 * -ffp-contract=off, no fast-math). Where the reference calls into torch for a transcendental
 * (cos/sin) whose last bit is platform dependent, the oracle uses the repo's CANONICAL
 * definition (flock_sincosf below: fp64 Cody-Waite reduction + fixed fp32 minimax polynomials
 * evaluated with fmaf), which the CUDA kernels restate independently, so CUDA == oracle
 * bit-for-bit while oracle == reference within 1e-5 relative (measured: <= 2 ulp per call).
 * k-NN order is the canonical (d2, j) ascending with j != i; the
 * reference's torch.topk agrees wherever it is well defined.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_K 16

typedef struct {
    int32_t variant;          /* 0 v2, 1 uw, 2 uwd */
    int32_t num_envs;
    int32_t num_agents;
    int32_t k;
    int32_t rigid_boundary;   /* check_boundary: gym_flock_v2.py:273 */
    int32_t periodic;         /* 1: min-image distances in step (gym_flock_v2.py:135), 0: Euclidean */
    int32_t obs_hist;         /* uw: memory_size = 4 (gym_flock_uw.py:59); else 1 */
    int32_t env_offset;       /* global id of local env 0 (multi-GPU sharding) */
    float boundary;           /* range_start[1] (gym_flock_v2.py:46) */
    float range_lo;           /* range_start[0] */
    float reset_hi;           /* v2/uwd: range_start[1]; uw: range_start[1]//2 (gym_flock_uw.py:87-89) */
    float heading_hi;         /* 1.5pi / 2pi / pi/1.2 (v2:96, uw:92, uwd:133) */
    float sensor_range;
    float collision_distance;
    float reset_collision_distance; /* uwd: 4 (gym_flock_uw_discrete.py:145), else = collision_distance */
    float max_linear_velocity;
    float act_noise_std;      /* uwd: 0.1 (gym_flock_uw_discrete.py:333-334) */
    float range_noise_std;    /* optional sensing noise (extension, reference has none): default 0 */
    uint64_t seed;
} orc_cfg_t;

/* ------------------------------------------------------------------------------------------
 * Canonical transcendental helpers (shared DEFINITION with the CUDA kernels, separate code).
 * ---------------------------------------------------------------------------------------- */

/* sin/cos of a binary32 angle. Reduction in fp64: q = rint(h*2/pi), r = h - q*pi/2 with a
 * two-term Cody-Waite split; then fp32 polynomials on |r| <= pi/4. |h| >= 1e9 or non-finite
 * gives NaN (callers pass it through nan_to_num like the reference does with torch's NaN). */
static void flock_sincosf(float h, float *sn, float *cs)
{
    double hd = (double)h;
    if (!(fabs(hd) < 1.0e9)) { *sn = NAN; *cs = NAN; return; }
    double q = rint(hd * 0.63661977236758134308);
    double r = fma(-q, 1.57079632673412561417, hd);
    r = fma(-q, 6.07710050650619224932e-11, r);
    float x = (float)r;
    float z = x * x;
    /* sin(x) = x + x*z*(S1 + z*(S2 + z*S3)) */
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    float xz = x * z;
    float s = fmaf(xz, ps, x);
    /* cos(x) = 1 - z/2 + z*z*(C1 + z*(C2 + z*C3)) */
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    float zz = z * z;
    float c = fmaf(zz, pc, fmaf(z, -0.5f, 1.0f));
    int quad = (int)(((long long)q) & 3);
    float so, co;
    switch (quad) {
        case 0: so = s;  co = c;  break;
        case 1: so = c;  co = -s; break;
        case 2: so = -s; co = -c; break;
        default: so = -c; co = s; break;
    }
    *sn = so; *cs = co;
}

/* natural log of m * 2^-24 for an integer m in [1, 2^24], binary32 with explicit fmaf (canonical
 * definition shared with the CUDA kernels; accuracy ~1e-7 absolute, plenty for a noise source). */
static float flock_log_u24(uint32_t m)
{
    float u = (float)m * (1.0f / 16777216.0f);
    uint32_t bits; memcpy(&bits, &u, 4);
    int e = (int)((bits >> 23) & 0xffu) - 126;                /* u = f * 2^e, f in [0.5,1) */
    bits = (bits & 0x007fffffu) | 0x3f000000u;
    float f; memcpy(&f, &bits, 4);
    if (f < 0.70710678118654752440f) { f = f * 2.0f; e -= 1; } /* f in [sqrt.5, sqrt2) */
    float num = f - 1.0f, den = f + 1.0f;
    float s = num / den;
    float s2 = s * s;
    float p = 1.0f / 9.0f;
    p = fmaf(p, s2, 1.0f / 7.0f);
    p = fmaf(p, s2, 1.0f / 5.0f);
    p = fmaf(p, s2, 1.0f / 3.0f);
    p = fmaf(p, s2, 1.0f);
    float two_s = 2.0f * s;
    float lnf = two_s * p;
    return fmaf((float)e, 0.69314718055994530942f, lnf);
}

/* Philox4x32-10 (Salmon et al., SC'11; constants of Random123 philox.h). */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* Stream layout (canonical, GPU-count invariant): key = seed; counter =
 * (global env id, agent id, epoch, tag word). tag 0: reset draw (epoch = per-env attempt counter
 * reset_epoch, tag word 0); tag 1: actuation noise, tag 2: random actions (epoch = ep_len[env] =
 * steps since the env's last reset, tag word = tag + 4 * reset_epoch[env]). */
enum { TAG_RESET = 0, TAG_NOISE = 1, TAG_ACTION = 2, TAG_RANGE = 3 };

static float u24(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); } /* [0,1), torch.rand grid */

/* two independent standard normals (Box-Muller) from two 32-bit words */
static void flock_normal2(uint32_t r0, uint32_t r1, float *z0, float *z1)
{
    float ln = flock_log_u24((r0 >> 8) + 1u);               /* u1 in (0,1] => ln <= 0 */
    float m2 = -2.0f * ln;
    float rad = sqrtf(m2);
    float theta = 6.28318530717958647692f * u24(r1);
    float sn, cs;
    flock_sincosf(theta, &sn, &cs);
    *z0 = rad * cs;
    *z1 = rad * sn;
}

void orc_sincosf(const float *h, int n, float *sn, float *cs)
{
    for (int i = 0; i < n; ++i) flock_sincosf(h[i], &sn[i], &cs[i]);
}

void orc_normal2(const uint32_t *r, int n, float *z)
{
    for (int i = 0; i < n; ++i) flock_normal2(r[2 * i], r[2 * i + 1], &z[2 * i], &z[2 * i + 1]);
}

/* ------------------------------------------------------------------------------------------
 * Reference semantics, one env at a time.
 * ---------------------------------------------------------------------------------------- */

/* torch.clamp propagates NaN (gym_flock_v2.py:327,331) */
static float clamp_nan(float v, float lo, float hi)
{
    if (v != v) return v;
    float t = v < lo ? lo : v;
    return t > hi ? hi : t;
}

/* torch.nan_to_num defaults (gym_flock_v2.py:346): NaN->0, +-inf -> +-FLT_MAX */
static float nan_to_num(float v)
{
    if (v != v) return 0.0f;
    if (v > 3.4028234663852886e38f) return 3.4028234663852886e38f;
    if (v < -3.4028234663852886e38f) return -3.4028234663852886e38f;
    return v;
}

/* check_boundary, per coordinate (gym_flock_v2.py:271-304; uw :223-256; uwd :278-311) */
static float wrap_coord(float c, float B, int rigid)
{
    if (rigid) {
        c = (c < B) ? c : B;
        c = (c > 0.0f) ? c : 0.0f;
    } else {
        c = (c < B) ? c : 0.001f;
        c = (c > 0.0f) ? c : B;
    }
    return c;
}

static const float UWD_MU_W[5] = { -1.2f, -0.5f, 0.0f, 0.5f, 1.2f }; /* gym_flock_uw_discrete.py:59-75 */

/* _updateState + check_boundary for one agent.
 * v2: gym_flock_v2.py:317-350; uw: gym_flock_uw.py:269-302 (heading=False branch);
 * uwd: gym_flock_uw_discrete.py:324-366. `act` points at this agent's action
 * (2 floats for v2/uw, 1 float-coded id for uwd); nz = additive actuation noise (uwd). */
static void integrate_agent(const orc_cfg_t *c, const float *act, float nzu, float nzw, float dt,
                            float *x, float *y, float *h, float *vx_out, float *vy_out)
{
    float vx, vy;
    if (c->variant == 0) {
        float w = clamp_nan(act[1], -1.57079632679489661923f, 1.57079632679489661923f);
        float wd = w * dt;
        *h = *h + wd;
        float u = clamp_nan(act[0], 0.005f, c->max_linear_velocity);
        float sn, cs;
        flock_sincosf(*h, &sn, &cs);
        vx = u * cs;
        vy = u * sn;
    } else if (c->variant == 1) {
        float ax = act[0], ay = act[1];
        float n2 = ax * ax;
        n2 = fmaf(ay, ay, n2);                     /* torch.norm accumulates like fma */
        float n = sqrtf(n2);
        vx = ax / n;
        vy = ay / n;
    } else {
        float a = act[0];
        /* int(act): truncation (uwd :329). Ids outside the dictionary (KeyError upstream) are
         * clamped to it, NaN -> 0; the clamp happens in float so the conversion is always defined. */
        a = (a == a) ? a : 0.0f;
        a = a < 0.0f ? 0.0f : a;
        a = a > 9.0f ? 9.0f : a;
        int id = (int)a;
        float mu_u = id < 5 ? 0.2f : 0.6f;
        float mu_w = UWD_MU_W[id % 5];
        float u = mu_u + nzu;                      /* torch.normal(mean, 0.1) = mean + noise */
        float w = mu_w + nzw;
        w = clamp_nan(w, -0.025f, 0.025f);
        float wd = w * dt;
        *h = *h + wd;
        u = clamp_nan(u, 5e-6f, c->max_linear_velocity);
        float sn, cs;
        flock_sincosf(*h, &sn, &cs);
        vx = u * cs;
        vy = u * sn;
        float n2 = vx * vx;
        n2 = fmaf(vy, vy, n2);
        float n = sqrtf(n2);
        vx = vx / n;
        vy = vy / n;
    }
    vx = nan_to_num(vx);
    vy = nan_to_num(vy);
    vx = vx * dt;
    vy = vy * dt;
    *x = *x + vx;
    *y = *y + vy;
    *x = wrap_coord(*x, c->boundary, c->rigid_boundary);
    *y = wrap_coord(*y, c->boundary, c->rigid_boundary);
    *vx_out = vx;
    *vy_out = vy;
}

/* squared distance of a pair: periodic min-image (gym_flock_v2.py:140-144) or Euclidean
 * (gym_flock_v2.py:166-168, gym_flock_uw.py:136-138, gym_flock_uw_discrete.py:184-186) */
static float pair_d2(float xi, float yi, float xj, float yj, float B, float halfB, int periodic)
{
    float dx = fabsf(xi - xj);
    float dy = fabsf(yi - yj);
    if (periodic) {
        if (dx > halfB) dx = B - dx;
        if (dy > halfB) dy = B - dy;
    }
    float a = dx * dx;
    /* Euclidean sites are torch.norm upstream, which accumulates like fma(dy, dy, dx*dx) (SURVEY 8c:
     * with this form the uw golden trajectories are reproduced bit for bit); the periodic metric is an
     * explicit multiply / add chain (gym_flock_v2.py:144) and stays unfused. */
    if (!periodic) return fmaf(dy, dy, a);
    float b = dy * dy;
    return a + b;
}

/* k smallest (d2, j), j != i, ascending; then sqrt + clamp(0, sensor_range)
 * (gym_flock_v2.py:147-151). */
static void knn_row(const orc_cfg_t *c, const float *x, const float *y, int i, int periodic,
                    float *d_out, int32_t *j_out)
{
    const int N = c->num_agents, k = c->k;
    float bd[ORC_MAX_K]; int32_t bj[ORC_MAX_K];
    for (int s = 0; s < k; ++s) { bd[s] = INFINITY; bj[s] = -1; }
    const float B = c->boundary, halfB = (float)((double)c->boundary / 2.0);
    for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        float d2 = pair_d2(x[i], y[i], x[j], y[j], B, halfB, periodic);
        if (!(d2 < bd[k - 1])) continue;           /* strict: equal d2 keeps the lower index */
        int s = k - 1;
        while (s > 0 && d2 < bd[s - 1]) { bd[s] = bd[s - 1]; bj[s] = bj[s - 1]; --s; }
        bd[s] = d2; bj[s] = j;
    }
    for (int s = 0; s < k; ++s) {
        float d = sqrtf(bd[s]);
        d = d < 0.0f ? 0.0f : d;
        d = d > c->sensor_range ? c->sensor_range : d;
        d_out[s] = d;
        j_out[s] = bj[s];
    }
}

/* Everything after positions are final: distances, k-NN, collisions, obs, dones, reward.
 * `in_reset`: use Euclidean metric + reset collision distance, skip reward/prev_h update
 * (gym_flock_v2.py:99-106). Returns env_done. */
static int sense_env(const orc_cfg_t *c, int in_reset, const float *x, const float *y, const float *h,
                     float *prev_h, float *obs /* [N][H][k] */, int32_t *nn /* [N][k] */,
                     float *reward /* [N] */, uint8_t *agent_done /* [N] */)
{
    const int N = c->num_agents, k = c->k, H = c->obs_hist;
    const int periodic = in_reset ? 0 : c->periodic;
    const float cd = in_reset ? c->reset_collision_distance : c->collision_distance;
    int env_done = 0;
    float comx = 0.0f, comy = 0.0f, hmean = 0.0f;
    if (!in_reset && c->variant == 1) {            /* torch.mean(positions, 0): gym_flock_uw.py:193 */
        float sx = 0.0f, sy = 0.0f;
        for (int j = 0; j < N; ++j) { sx = sx + x[j]; sy = sy + y[j]; }
        comx = sx / (float)N; comy = sy / (float)N;
    }
    if (!in_reset && c->variant == 2) {            /* torch.sum(headings)/N: uwd :256 */
        float sh = 0.0f;
        for (int j = 0; j < N; ++j) sh = sh + h[j];
        hmean = sh / (float)N;
    }
    for (int i = 0; i < N; ++i) {
        float d[ORC_MAX_K]; int32_t jj[ORC_MAX_K];
        knn_row(c, x, y, i, periodic, d, jj);
        int coll = 0;
        for (int s = 0; s < k; ++s) coll |= (d[s] < cd);         /* _computeCollisions v2:212 */
        agent_done[i] = (uint8_t)coll;                             /* _computeDone v2:306 */
        env_done |= coll;
        float *o = obs + (size_t)i * H * k;
        if (H > 1) {                                               /* _computeObs uw:120-123 (roll, newest first) */
            for (int t = H - 1; t > 0; --t)
                for (int s = 0; s < k; ++s) o[t * k + s] = o[(t - 1) * k + s];
        }
        for (int s = 0; s < k; ++s) { o[s] = d[s]; nn[(size_t)i * k + s] = jj[s]; }
        if (in_reset) continue;
        float r;
        if (c->variant == 0) {
            r = coll ? -5.0f : 0.01f;                              /* v2:217-220,268 */
        } else if (c->variant == 1) {
            float pen = coll ? -5.0f : 0.01f;                      /* uw:186-189 */
            float ddx = x[i] - comx, ddy = y[i] - comy;            /* uw:191-197 */
            float q = ddx * ddx; q = fmaf(ddy, ddy, q);
            float dc = sqrtf(q);
            float thr = (float)((double)c->collision_distance * 4.0);
            float rcom = dc < thr ? 0.01f : 0.0f;
            float diff = fabsf(prev_h[i] - h[i]);                  /* uw:201-204 */
            float rang = diff > 0.27f ? -0.01f : 0.001f;
            prev_h[i] = h[i];
            r = pen + rcom;
            r = r + rang;                                          /* uw:218 */
        } else {
            float pen = coll ? -9.0f : 0.0f;                       /* uwd:234-237 */
            float err = fabsf(hmean - h[i]);                       /* uwd:254-258 */
            float ral = err > 0.20f ? 0.0f : 0.1f;
            prev_h[i] = h[i];                                      /* side effect of uwd:249-252 */
            r = pen + ral;                                         /* uwd:275 */
        }
        reward[i] = r;
    }
    return env_done;
}

typedef struct {
    float *x, *y, *h, *prev_h;     /* [E][N] */
    float *vx, *vy;                /* [E][N] displacement of the last step (reference `velocities`) */
    float *obs;                    /* [E][N][H][k] */
    int32_t *nn;                   /* [E][N][k] */
    float *reward;                 /* [E][N] */
    uint8_t *agent_done;           /* [E][N] */
    uint8_t *env_done;             /* [E] */
    uint32_t *reset_epoch;         /* [E] Philox attempt counter */
    int64_t *ep_return_fx;         /* [E] episode return, sum_i reward_i in 2^-32 fixed point (logging, main.py:44) */
    int32_t *ep_len;               /* [E] steps since reset */
    uint64_t *stats;               /* [8] episodes, ep steps, ep return fx, reset attempts, gave up */
} orc_buf_t;

/* step(): gym_flock_v2.py:71-83, gym_flock_uw.py:69-81, gym_flock_uw_discrete.py:110-122.
 * actions: [E][N][2] (v2/uw) or [E][N] (uwd). noise: NULL -> Philox (tag 1, epoch step_index)
 * when act_noise_std > 0; else [E][N][2] additive (n_u, n_w). */
void orc_step(const orc_cfg_t *c, orc_buf_t *b, const float *actions, const float *noise,
              float dt, int nthreads)
{
    const int E = c->num_envs, N = c->num_agents, k = c->k, H = c->obs_hist;
    const int aw = c->variant == 2 ? 1 : 2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int e = 0; e < E; ++e) {
        size_t o = (size_t)e * N;
        for (int i = 0; i < N; ++i) {
            float nzu = 0.0f, nzw = 0.0f;
            if (c->variant == 2) {
                if (noise) { nzu = noise[(o + i) * 2]; nzw = noise[(o + i) * 2 + 1]; }
                else if (c->act_noise_std > 0.0f) {
                    uint32_t r[4];
                    philox4x32_10((uint32_t)(c->env_offset + e), (uint32_t)i, (uint32_t)b->ep_len[e],
                                  TAG_NOISE + (b->reset_epoch[e] << 2),
                                  (uint32_t)c->seed, (uint32_t)(c->seed >> 32), r);
                    float z0, z1;
                    flock_normal2(r[0], r[1], &z0, &z1);
                    nzu = c->act_noise_std * z0;
                    nzw = c->act_noise_std * z1;
                }
            }
            integrate_agent(c, actions + (o + i) * aw, nzu, nzw, dt,
                            &b->x[o + i], &b->y[o + i], &b->h[o + i], &b->vx[o + i], &b->vy[o + i]);
        }
        b->env_done[e] = (uint8_t)sense_env(c, 0, b->x + o, b->y + o, b->h + o, b->prev_h + o,
                                            b->obs + o * H * k, b->nn + o * k, b->reward + o,
                                            b->agent_done + o);
        for (int i = 0; i < N; ++i)
            b->ep_return_fx[e] += (int64_t)llrint((double)b->reward[o + i] * 4294967296.0);
        b->ep_len[e] += 1;
    }
}

/* canonical random actions for step_n (tag 2): v2 U[-1.5,1.5)^2 (action_space v2:58),
 * uw U[-1,1)^2 (uw:57), uwd id = floor(U*k) (Discrete(k) uwd:98). */
void orc_random_actions(const orc_cfg_t *c, const orc_buf_t *b, uint32_t step_offset, float *actions)
{
    const int E = c->num_envs, N = c->num_agents;
    for (int e = 0; e < E; ++e)
        for (int i = 0; i < N; ++i) {
            uint32_t r[4];
            philox4x32_10((uint32_t)(c->env_offset + e), (uint32_t)i, (uint32_t)b->ep_len[e] + step_offset,
                          TAG_ACTION + (b->reset_epoch[e] << 2),
                          (uint32_t)c->seed, (uint32_t)(c->seed >> 32), r);
            size_t a = (size_t)e * N + i;
            if (c->variant == 0) {
                float t0 = u24(r[0]) * 3.0f, t1 = u24(r[1]) * 3.0f;
                actions[a * 2] = t0 - 1.5f; actions[a * 2 + 1] = t1 - 1.5f;
            } else if (c->variant == 1) {
                float t0 = u24(r[0]) * 2.0f, t1 = u24(r[1]) * 2.0f;
                actions[a * 2] = t0 - 1.0f; actions[a * 2 + 1] = t1 - 1.0f;
            } else {
                actions[a] = (float)(uint32_t)(((uint64_t)r[0] * (uint32_t)c->k) >> 32);
            }
        }
}

/* reset(): gym_flock_v2.py:85-108, gym_flock_uw.py:83-111, gym_flock_uw_discrete.py:124-156.
 * mask NULL -> all envs. init NULL -> Philox draws with bounded rejection (max_attempts; the
 * reference recurses without bound); init = [3][E][N] (x, y, h) -> injected state, one pass.
 * env_done reports whether the accepted/injected start still collides. Returns #envs that
 * exhausted max_attempts. */
int orc_reset(const orc_cfg_t *c, orc_buf_t *b, const uint8_t *mask, const float *init,
              int max_attempts, int keep_outputs, int nthreads)
{
    const int E = c->num_envs, N = c->num_agents, k = c->k, H = c->obs_hist;
    int gave_up = 0;
    (void)nthreads;
    for (int e = 0; e < E; ++e) {
        if (mask && !mask[e]) continue;
        size_t o = (size_t)e * N;
        int done = 1;
        int att = 0;
        /* auto-reset mode keeps reward / dones of the step that ended the episode */
        float *keep_r = NULL; uint8_t *keep_d = NULL;
        if (keep_outputs) {
            keep_r = (float *)malloc(sizeof(float) * N); keep_d = (uint8_t *)malloc(N);
            memcpy(keep_r, b->reward + o, sizeof(float) * N); memcpy(keep_d, b->agent_done + o, N);
        }
        for (; att < (init ? 1 : max_attempts) && done; ++att) {
            for (int i = 0; i < N; ++i) {
                float px, py, ph;
                if (init) {
                    px = init[o + i]; py = init[(size_t)E * N + o + i]; ph = init[(size_t)2 * E * N + o + i];
                } else {
                    uint32_t r[4];
                    philox4x32_10((uint32_t)(c->env_offset + e), (uint32_t)i, b->reset_epoch[e], TAG_RESET,
                                  (uint32_t)c->seed, (uint32_t)(c->seed >> 32), r);
                    float span = c->range_lo - c->reset_hi;            /* (r0 - r1) * U + r1, v2:87-89 */
                    float tx = span * u24(r[0]); px = tx + c->reset_hi;
                    float ty = span * u24(r[1]); py = ty + c->reset_hi;
                    float th = (0.0f - c->heading_hi) * u24(r[2]);     /* v2:96 */
                    ph = th + c->heading_hi;
                }
                b->x[o + i] = wrap_coord(px, c->boundary, c->rigid_boundary);   /* v2:99 */
                b->y[o + i] = wrap_coord(py, c->boundary, c->rigid_boundary);
                b->h[o + i] = ph;
                b->prev_h[o + i] = 0.0f;                                /* v2:95 */
                b->vx[o + i] = 0.0f; b->vy[o + i] = 0.0f;               /* v2:94 */
                b->reward[o + i] = 0.0f;
            }
            if (!init) b->reset_epoch[e] += 1u;
            if (H > 1) memset(b->obs + o * H * k, 0, sizeof(float) * (size_t)N * H * k); /* uw:100-102 */
            done = sense_env(c, 1, b->x + o, b->y + o, b->h + o, b->prev_h + o,
                             b->obs + o * H * k, b->nn + o * k, b->reward + o, b->agent_done + o);
        }
        if (keep_outputs) {
            memcpy(b->reward + o, keep_r, sizeof(float) * N); memcpy(b->agent_done + o, keep_d, N);
            free(keep_r); free(keep_d);
        } else {
            b->env_done[e] = (uint8_t)done;
        }
        if (done && !init) gave_up += 1;
        if (b->ep_len[e] > 0) {                    /* close the running episode (logging only) */
            b->stats[0] += 1;
            b->stats[1] += (uint64_t)b->ep_len[e];
            b->stats[2] += (uint64_t)b->ep_return_fx[e];
        }
        if (!init) {
            b->stats[3] += (uint64_t)att;
            if (done) b->stats[4] += 1;
        }
        b->ep_len[e] = 0;
        b->ep_return_fx[e] = 0;
    }
    return gave_up;
}

/* Optional sensing noise (extension of the north star, not in the reference): newest obs row
 * becomes clamp(d + std*z, 0, sensor_range); Philox (global env, agent + 65536*call, ep_len, tag 3
 * + 4*reset_epoch). Applied after step / reset, only to masked envs. */
void orc_range_noise(const orc_cfg_t *c, orc_buf_t *b, const uint8_t *mask)
{
    const int E = c->num_envs, N = c->num_agents, k = c->k, H = c->obs_hist;
    if (!(c->range_noise_std > 0.0f)) return;
    for (int e = 0; e < E; ++e) {
        if (mask && !mask[e]) continue;
        for (int a = 0; a < N; ++a) {
            float *o = b->obs + ((size_t)e * N + a) * H * k;
            for (int cc = 0; cc * 4 < k; ++cc) {
                uint32_t r[4];
                philox4x32_10((uint32_t)(c->env_offset + e), (uint32_t)a + 65536u * (uint32_t)cc, (uint32_t)b->ep_len[e],
                              TAG_RANGE + (b->reset_epoch[e] << 2), (uint32_t)c->seed, (uint32_t)(c->seed >> 32), r);
                float z[4];
                flock_normal2(r[0], r[1], &z[0], &z[1]);
                flock_normal2(r[2], r[3], &z[2], &z[3]);
                for (int s = 0; s < 4 && cc * 4 + s < k; ++s) {
                    float nz = c->range_noise_std * z[s];
                    float d = o[cc * 4 + s] + nz;
                    d = d < 0.0f ? 0.0f : d;
                    d = d > c->sensor_range ? c->sensor_range : d;
                    o[cc * 4 + s] = d;
                }
            }
        }
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
