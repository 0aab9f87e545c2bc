/*
 * flock_b200.h -- C ABI of libflock_b200.so: the B200 (sm_100a) batched range-only flocking
 * environment. Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (RetamalVictor/marl-range-flocking) has no FFI: its boundary is the Python
 * class `MultiAgentEnv` of environments/gym_flock_v2.py, gym_flock_uw.py and
 * gym_flock_uw_discrete.py. Each entry point below states which reference method(s) it replaces;
 * the Python mirror of that class API lives in marl_range_flocking_b200/ and the binding a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative FLOCK_E_* code; the message is available
 *     from flock_last_error() (thread local). Nothing throws across the ABI.
 *   - device buffers are OWNED BY THE CALLER (torch tensors in the Python host); the library
 *     borrows raw device pointers via flock_bind() and never frees them.
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*); only the
 *     *_host entry points synchronise that stream before returning.
 *   - one handle per device; a handle is not thread safe.
 *   - state layout: structure of arrays, float32, `[E][N]` row-major (E envs, N agents).
 */
#ifndef FLOCK_B200_H
#define FLOCK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FLOCK_API __attribute__((visibility("default")))
#else
#define FLOCK_API
#endif

#define FLOCK_ABI_VERSION 2
#define FLOCK_MAX_K 8          /* neighbours per agent (reference uses 3, 4; BASELINE cfg 5 uses 8) */
#define FLOCK_MAX_AGENTS 8192  /* per env; whole env is staged in shared memory */

enum {
    FLOCK_OK = 0,
    FLOCK_E_INVALID = -1,   /* bad argument / configuration */
    FLOCK_E_UNBOUND = -2,   /* flock_bind() not called or a required pointer is NULL */
    FLOCK_E_CUDA = -3,      /* CUDA runtime error (message has the cudaError string) */
    FLOCK_E_NO_DEVICE = -4  /* no sm_100 device: there is NO CPU fallback */
};

enum { FLOCK_V2 = 0, FLOCK_UW = 1, FLOCK_UWD = 2 };

/* Mirrors the constructor arguments of MultiAgentEnv.__init__ (gym_flock_v2.py:21-69,
 * gym_flock_uw.py:20-67, gym_flock_uw_discrete.py:20-108) plus the per-variant constants that the
 * reference hard-codes inside reset()/step(). */
typedef struct flock_cfg_t {
    int32_t variant;            /* FLOCK_V2 / FLOCK_UW / FLOCK_UWD */
    int32_t num_envs;           /* E: env instances on this device */
    int32_t num_agents;         /* N: `agents` */
    int32_t k;                  /* `k` nearest neighbours, 1..FLOCK_MAX_K, N >= k+1 */
    int32_t rigid_boundary;     /* `rigid_boundary` (check_boundary, gym_flock_v2.py:271-304) */
    int32_t periodic;           /* 1: min-image metric in step (_computePeriodicDistances, v2:135) */
    int32_t obs_hist;           /* 4 for uw (memory_size, gym_flock_uw.py:59), else 1 */
    int32_t env_offset;         /* global index of local env 0 (sharding across GPUs) */
    float boundary;             /* range_start[1] */
    float range_lo;             /* range_start[0] */
    float reset_hi;             /* upper end of the reset box: r1 (v2, uwd) or r1//2 (uw:87-89) */
    float heading_hi;           /* 1.5*pi (v2:96), 2*pi (uw:92), pi/1.2 (uwd:133) */
    float sensor_range;
    float collision_distance;
    float reset_collision_distance; /* uwd: 4 (gym_flock_uw_discrete.py:145); else collision_distance */
    float max_linear_velocity;
    float act_noise_std;        /* uwd: 0.1 (gym_flock_uw_discrete.py:333-334); else 0 */
    float range_noise_std;      /* optional sensing noise on the observed ranges, N(0, std); the reference has
                                   none, default 0 (parity). Applied by a follow-up kernel only when > 0 */
    uint64_t seed;              /* Philox4x32-10 key */
} flock_cfg_t;

/* Device pointers of the caller-owned buffers. Nullable outputs are skipped by the kernels. */
typedef struct flock_buffers_t {
    float *x, *y, *h;               /* [E][N] positions and headings, updated in place (both kernel paths) */
    float *prev_h;                  /* [E][N] `prev_headings` (uw reward term, gym_flock_uw.py:201) */
    float *vx, *vy;                 /* [E][N] last displacement = reference `velocities`; nullable */
    float *obs;                     /* window layout (obs_head == NULL): [E][N][obs_hist][k] newest first
                                       (gym_flock_uw.py:120-123); ring layout (uw, obs_head != NULL):
                                       [E][obs_hist][N][k], see obs_head */
    int32_t *obs_head;              /* [E], uw only, nullable. Non-NULL selects the RING layout of the observation
                                       history: row r of the newest-first window lives in slot (obs_head[e] + r) %
                                       obs_hist; a step moves the head back by one slot and writes ONLY the new row
                                       (4k bytes per agent instead of reading 36 and writing 48 at H = 4, k = 3).
                                       flock_obs_window() materialises the window on request */
    int32_t *nn_idx;                /* [E][N][k] `nearest_neighbors` (v2:150); nullable */
    float *reward;                  /* [E][N] (reference shape (N,1)) */
    uint8_t *agent_done;            /* [E][N] `dones[0]` (v2:314) */
    uint8_t *env_done;              /* [E]    `dones[1]` (v2:315), stays on the device */
    uint32_t *reset_epoch;          /* [E] Philox attempt counter of reset() */
    int64_t *ep_return_fx;          /* [E] sum over the episode of sum_i reward_i, fixed point 2^-32; nullable */
    int32_t *ep_len;                /* [E] steps since the last reset */
    uint64_t *stats;                /* [8] FLOCK_STAT_* accumulators flushed by reset; nullable */
} flock_buffers_t;

enum {
    FLOCK_STAT_EPISODES = 0,     /* episodes closed by reset() (ep_len > 0) */
    FLOCK_STAT_EP_STEPS = 1,     /* sum of their lengths */
    FLOCK_STAT_EP_RETURN_FX = 2, /* sum of their returns (two's complement int64, 2^-32 units) */
    FLOCK_STAT_RESET_ATTEMPTS = 3,
    FLOCK_STAT_RESET_GAVE_UP = 4,
    FLOCK_STAT_COUNT = 8
};

typedef struct flock_env flock_env_t;

/* MultiAgentEnv.__init__ (gym_flock_v2.py:21). Validates the configuration, selects the device,
 * allocates only the small host-call staging buffers. */
FLOCK_API int flock_create(const flock_cfg_t *cfg, int device, flock_env_t **out);
FLOCK_API void flock_destroy(flock_env_t *env);

/* Attach the caller-owned device buffers. */
FLOCK_API int flock_bind(flock_env_t *env, const flock_buffers_t *bufs);

/* MultiAgentEnv.reset (gym_flock_v2.py:85-108, gym_flock_uw.py:83-111,
 * gym_flock_uw_discrete.py:124-156), batched and masked.
 *   env_mask   device uint8[E], nullable (NULL = every env). Pass the env_done buffer to reset
 *              exactly the finished envs without a host round trip.
 *   init_state device float[3][E][N] (x, y, heading), nullable. NULL: Philox draws with at most
 *              `max_attempts` rejection rounds per env (the reference recurses without bound);
 *              non-NULL: the given state is installed (parity injection), one pass.
 *   flags      FLOCK_RESET_KEEP_OUTPUTS: auto-reset mode, leave reward / agent_done / env_done
 *              of the step that finished the episode untouched (env_mask may then alias the
 *              env_done buffer); 0: write reward = 0 and the done flags of the new start
 *              (env_done = the start still collides).
 * Always writes the state, obs and nn_idx of the reset envs and closes their episode counters. */
#define FLOCK_RESET_KEEP_OUTPUTS 1
FLOCK_API int flock_reset(flock_env_t *env, const uint8_t *env_mask, const float *init_state,
                          int max_attempts, int flags, void *stream);

/* MultiAgentEnv.step (gym_flock_v2.py:71-83, gym_flock_uw.py:69-81,
 * gym_flock_uw_discrete.py:110-122): ONE fused kernel = _updateState + check_boundary +
 * _compute(Periodic)Distances + _computeCollisions + _computeObs + _computeDone + _computeReward.
 *   actions device float[E][N][2] (v2: linear, angular; uw: velocity direction) or float[E][N]
 *           (uwd: float-coded action id 0..9).
 *   noise   uwd only, nullable: device float[E][N][2] additive (linear, angular) actuation noise
 *           replacing the in-kernel Philox N(0, act_noise_std) draw (reference: torch.normal,
 *           gym_flock_uw_discrete.py:333-334). */
FLOCK_API int flock_step(flock_env_t *env, const float *actions, float dt, const float *noise, void *stream);

/* Auto-reset mode (the batched form of "caller resets on done[1]", main.py:31-51): when enabled,
 * every flock_step also restarts the envs whose step ended with a collision -- inside the same
 * kernel launch when N <= 32, by a follow-up masked reset otherwise. reward / agent_done / env_done
 * keep the values of the finishing step; obs (and nn_idx) of a restarted env are those of its new
 * first state. Equivalent to flock_step followed by flock_reset(env_done, NULL, max_attempts,
 * FLOCK_RESET_KEEP_OUTPUTS). */
FLOCK_API int flock_set_auto_reset(flock_env_t *env, int enabled, int max_attempts);

/* T consecutive steps with the canonical in-kernel random actions (action_space sampling:
 * v2 U[-1.5,1.5)^2 gym_flock_v2.py:58, uw U[-1,1)^2 gym_flock_uw.py:57, uwd id in [0,k)
 * gym_flock_uw_discrete.py:98). N <= 32: one persistent launch with the state in registers;
 * N > 32: T launches. Outputs hold the last step. */
FLOCK_API int flock_step_n(flock_env_t *env, int num_steps, float dt, void *stream);

/* Streamed rollout: `num_steps` consecutive steps in ONE launch (N <= 32, no sensing noise; otherwise num_steps launches
 * plus copies), the batched form of the reference's episode loop (main.py:24-51) feeding a replay buffer
 * (learners/maddpg_official_rnn/memory_rnn.py:53-67). The state stays in registers; per step the kernel reads that step's
 * actions and writes that step's results into TIME-MAJOR trajectory buffers (all device pointers):
 *   actions_T    float[T][E][N][2] (float[T][E][N] for uwd), or NULL = the canonical Philox random actions of flock_step_n
 *   obs_T        float[T][E][N][k]   the new range row of every step (for uw the (4, k) window of step t is rows t, t-1,
 *                                    t-2, t-3 of this buffer: a time-major history needs no separate window)
 *   reward_T     float[T][E][N],  agent_done_T uint8[T][E][N],  env_done_T uint8[T][E]
 *   nn_T         int32[T][E][N][k] or NULL (needs a bound nn_idx buffer)
 * With flock_set_auto_reset enabled, an env whose step ends with a collision is re-drawn inside the kernel before the
 * next step, exactly as flock_step does: reward / dones of slice t are those of the finishing step, obs_T[t] (and
 * nn_T[t]) of a restarted env are those of its new first state. Afterwards the env's own buffers hold the state and
 * the outputs of the last step, as after flock_step_n. Bit-identical to num_steps calls of flock_step. */
FLOCK_API int flock_rollout_n(flock_env_t *env, int num_steps, const float *actions_T, float dt, float *obs_T,
                              float *reward_T, uint8_t *agent_done_T, uint8_t *env_done_T, int32_t *nn_T, void *stream);

/* uw ring layout only: write the newest-first window [E][N][obs_hist][k] (the reference's observation,
 * gym_flock_uw.py:120-123) into `out` (device). The fused actor kernel and flock_rollout_n never need it. */
FLOCK_API int flock_obs_window(flock_env_t *env, float *out, void *stream);

/* Fill `actions` (device, layout as in flock_step) with the canonical random actions that
 * flock_step_n would use `step_offset` steps from now. Philox4x32-10 stream layout: key = seed,
 * counter = (global env, agent, ep_len[env] + offset, tag + 4 * reset_epoch[env]); the epoch words
 * are per-env device counters, so streams do not depend on the GPU count or on host state. */
FLOCK_API int flock_random_actions(flock_env_t *env, uint32_t step_offset, float *actions, void *stream);

/* Host-buffer form of step(): host actions in, host obs / reward / agent_done / env_done out, stream
 * synchronised before returning. This is the end-to-end path timed as `e2e`.
 *   - pinned (device-visible) buffers, N <= 32: zero-copy -- the kernel reads the actions from host
 *     memory and, when the result set is <= 3 MiB, also writes the results there itself;
 *   - otherwise (pageable memory, N > 32, large result sets, auto-reset / sensing noise on): the
 *     actions are staged with one copy and the results come back with one packed copy when the
 *     caller laid obs | reward | agent_done | env_done out back to back on both sides.
 * Any output pointer may be NULL (staged path). FLOCK_ZEROCOPY=0/1 in the environment overrides the
 * policy. */
FLOCK_API int flock_step_host(flock_env_t *env, const float *h_actions, float dt, const float *h_noise,
                    float *h_obs, float *h_reward, uint8_t *h_agent_done, uint8_t *h_env_done,
                    void *stream);

/* flock_step_host without the final synchronisation: one staged H2D copy of the actions, the fused step,
 * one packed D2H copy of the results (copy engines on both sides: with several steps in flight they beat
 * the zero-copy kernel of flock_step_host); everything is enqueued on `stream` and the call returns at once. The host buffers must be pinned and must stay untouched until
 * the caller has synchronised `stream` (or an event recorded on it). Two handles driven on two streams
 * overlap one env batch's result copy with the other's action copy and step: the pipelined `e2e` leg of
 * bench.py. Same results as flock_step_host. */
FLOCK_API int flock_step_host_async(flock_env_t *env, const float *h_actions, float dt, const float *h_noise,
                          float *h_obs, float *h_reward, uint8_t *h_agent_done, uint8_t *h_env_done,
                          void *stream);
/* Blocks until the results of this handle's last flock_step_host_async have landed in the host buffers
 * (an event recorded behind the result copy; other work on the stream is not waited for). */
FLOCK_API int flock_wait_host(flock_env_t *env);

/* Host-side count of steps issued through this handle (informational). */
FLOCK_API uint32_t flock_get_step_index(const flock_env_t *env);
FLOCK_API int flock_set_step_index(flock_env_t *env, uint32_t step_index);

/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
FLOCK_API uint64_t flock_launch_count(const flock_env_t *env);

/* 0: warp-per-env-group path (N <= 32), 1: tiled path (N > 32). */
FLOCK_API int flock_path(const flock_env_t *env);

/* Row x neighbour pairs actually evaluated so far by the pruned large-swarm kernel (v2, N > 32, many
 * envs: blocks of neighbours that provably cannot contain a k-nearest neighbour are skipped; results
 * are identical to the all-pairs scan). Synchronises the device; `reset` != 0 zeroes the counter.
 * bench.py uses it to report the FP32 roofline on the work really done. */
FLOCK_API uint64_t flock_pairs_evaluated(flock_env_t *env, int reset);

/* Kernel choice on the tiled path: 0 = automatic (default), 1 = one thread per row (many envs),
 * 2 = one warp per row with a warp-shuffle bitonic top-k merge (few envs x large swarm). Results are
 * identical bit for bit; only speed differs. */
FLOCK_API int flock_set_tiled_mode(flock_env_t *env, int mode);

FLOCK_API const char *flock_last_error(void);
FLOCK_API int flock_abi_version(void);

/* Device-side step counters for the exploration noise fused into the policy kernels below. The host scalar `step`
 * of those calls is baked into a captured CUDA graph, so on replay every launch would redraw the same normals /
 * uniforms; with counters the Philox counter words become
 *     (env_offset + env, agent, step + env_step[env], tag + 16 * env_epoch[env])
 * read on the DEVICE at launch time. Pass the env's own per-env counters -- ep_len (steps since the last reset) and
 * reset_epoch (flock_buffers_t) -- and the draws are unique over the env's life, replay-safe and invariant under
 * sharding, exactly like the env's own streams. NULL (or NULL members) = host `step` only: do not capture that. */
typedef struct flock_noise_counters_t {
    const int32_t *env_step;    /* [E] device, nullable */
    const uint32_t *env_epoch;  /* [E] device, nullable */
} flock_noise_counters_t;

/* Fused per-agent actor MLP for the batched rollout ("MADDPG actor rollout", BASELINE configs[2]): the
 * N ActorNetworks of learners/maddpg_shared_critic/ddpg_network.py:85-141 (fc1 -> LayerNorm -> ReLU ->
 * fc2 -> LayerNorm -> ReLU -> mu -> tanh, 400 / 300 hidden units, 2 actions, input_dims <= 14) evaluated
 * for all envs in one launch on the tensor cores (tcgen05, bf16 operands, fp32 accumulation), replacing
 * the per-agent Python loop of learners/maddpg_shared_critic/train_flock.py:114-115.
 *   flock_actor_packed_bytes: size of the packed parameter image for `num_agents` agents.
 *   flock_actor_pack: `params` = 10 DEVICE pointers {w1 [A][in][400], b1 [A][400], ln1.weight, ln1.bias,
 *     w2 [A][400][300], b2 [A][300], ln2.weight, ln2.bias, w3 [A][300][2], b3 [A][2]}, float32, weights
 *     stored input-major (the transpose of torch.nn.Linear.weight); writes `packed` (device, 16-byte
 *     aligned). Call again after every parameter update.
 *   flock_actor_forward: obs [E][A][input_dims] float32 (the env's observation buffer) -> actions
 *     [E][A][2] float32 (what flock_step consumes). Asynchronous on `stream`. */
FLOCK_API size_t flock_actor_packed_bytes(int num_agents);
FLOCK_API int flock_actor_pack(int num_agents, int input_dims, int fc1_dims, int fc2_dims, int n_actions,
                     const float *const *params, void *packed, void *stream);
FLOCK_API int flock_actor_forward(const void *packed, const float *obs, float *actions, int num_envs, int num_agents,
                        int input_dims, void *stream);
/* The same with the exploration noise of the learner fused into the output stage: one Ornstein-Uhlenbeck process per
 * (env, agent, action) -- OUActionNoiseGPU, learners/maddpg_shared_critic/utils.py:6-21 (defaults theta 0.2, mu 0,
 * sigma 0.15, dt 1e-2), applied as mu' = mu + noise() (agent_simple_shared_critic.py:104):
 *     x <- x + theta (mu - x) dt + sigma sqrt(dt) N(0, 1),   actions = tanh(...) + x.
 * ou_state [E][A][2] float32 is read and updated in place (zero it to reset(), utils.py:20-21). The normals are
 * Philox4x32-10 with counter (env_offset + env, agent, step, tag) and key = seed: pass the rollout step as `step`. */
FLOCK_API int flock_actor_forward_ou(const void *packed, const float *obs, float *actions, int num_envs, int num_agents,
                           int input_dims, float *ou_state, float theta, float mu, float sigma, float dt,
                           uint64_t seed, uint32_t step, int env_offset, const flock_noise_counters_t *counters,
                           void *stream);

/* The same actor reading the env's observation history IN PLACE from the uw ring layout (flock_buffers_t.obs_head):
 * obs_ring [E][obs_hist][N][k], obs_head [E]; input feature r * k + c of (env, agent) is row r of the newest-first
 * window = slot (obs_head[env] + r) % obs_hist, exactly what obs.reshape(N, -1) feeds the reference actor
 * (train_flock.py:115). input_dims = obs_hist * k (obs_hist <= 4, k <= 3). ou_state NULL = no exploration noise. */
FLOCK_API int flock_actor_forward_ring(const void *packed, const float *obs_ring, const int32_t *obs_head, float *actions,
                             int num_envs, int num_agents, int obs_hist, int k, float *ou_state, float theta, float mu,
                             float sigma, float dt, uint64_t seed, uint32_t step, int env_offset,
                             const flock_noise_counters_t *counters, void *stream);

/* Fused recurrent MADDPG actor -- the policy of the reference's default main.py loop: `Actor` of
 * learners/maddpg_official_rnn/net.py:14-72 (fce(in,32) - GRUCell(32,32) - fc1(32,400) - ReLU - fc2(400,300) - ReLU -
 * [linear_speed: (tanh+1)/2 | angular_speed: 1.5 tanh]), one weight set per agent, replacing the per-agent Python loop
 * of MADDPG.get_actions (learners/maddpg_official_rnn/MADDPG.py:24-33). Two launches: fce + GRU in fp32, then the MLP
 * on the tensor cores (bf16 operands, fp32 accumulation).
 *   flock_rnn_actor_pack: `params` = 8 DEVICE pointers {w1 [A][32][400], b1 [A][400], w2 [A][400][300], b2 [A][300],
 *     w_linear [A][300], b_linear [A], w_angular [A][300], b_angular [A]}, float32, weights input-major; writes
 *     `packed` (flock_rnn_actor_packed_bytes(A) bytes, 16-byte aligned). Call again after a parameter update.
 *   flock_rnn_actor_forward: `front_params` = 6 DEVICE pointers {w_e [A][n_obs][32], b_e [A][32], w_ih [A][32][96],
 *     b_ih [A][96], w_hh [A][32][96], b_hh [A][96]} (used as they are, gate order r|z|n); obs [E][A][n_obs],
 *     hidden_in / hidden_out [E][A][32] (may alias) -> actions [E][A][2] = (linear, angular). n_obs <= 16. */
FLOCK_API size_t flock_rnn_actor_packed_bytes(int num_agents);
FLOCK_API int flock_rnn_actor_pack(int num_agents, int hidden_rnn, int hidden1, int hidden2, int n_actions,
                         const float *const *params, void *packed, void *stream);
FLOCK_API int flock_rnn_actor_forward(const void *packed, const float *const *front_params, const float *obs,
                            const float *hidden_in, float *hidden_out, float *actions, int num_envs, int num_agents,
                            int n_obs, void *stream);
/* The same with the learner's exploration noise (mu' = mu + noise.sample(), learners/maddpg_official_rnn/agent.py:61;
 * OrnsteinUhlenbeckProcess.sample, utils.py:43-47) fused into the output stage, one process per (env, agent, action):
 * ou_state [E][A][2] float32 read and updated in place; see flock_actor_forward_ou for the recurrence and the draws. */
FLOCK_API int flock_rnn_actor_forward_ou(const void *packed, const float *const *front_params, const float *obs,
                               const float *hidden_in, float *hidden_out, float *actions, int num_envs,
                               int num_agents, int n_obs, float *ou_state, float theta, float mu, float sigma,
                               float dt, uint64_t seed, uint32_t step, int env_offset,
                               const flock_noise_counters_t *counters, void *stream);

/* Fused VDN action selection ("VDN action selection", BASELINE configs[3]): QNet.forward + QNet.sample_action of
 * learners/vdn/net.py:11-58 for all envs and agents in one fp32 launch, replacing the per-agent Python loop
 * (net.py:30-35). Per agent: Linear(n_obs,64)-ReLU-Linear(64,32)-ReLU-[GRUCell(32,32)]-Linear(32,n_actions).
 *   params: 6 (10 if recurrent) DEVICE pointers {w1 [A][n_obs][64], b1 [A][64], w2 [A][64][32], b2 [A][32],
 *     wq [A][32][n_actions], bq [A][n_actions], w_ih [A][32][96], b_ih [A][96], w_hh [A][32][96], b_hh [A][96]},
 *     float32, weights input-major (transposed torch.nn.Linear / GRUCell weights, gate order r|z|n).
 *   obs [E][A][n_obs]; hidden_in / hidden_out [E][A][32] (recurrent; may alias); q_out [E][A][n_actions] or NULL;
 *   actions [E][A] float-coded ids or NULL: argmax (first maximum), except that with probability epsilon an env
 *   explores as a whole (one decision per env, net.py:54) and every agent draws a uniform id. The draws are
 *   Philox4x32-10 with counter (env_offset + env, agent, step, tag) and key = seed: invariant under sharding.
 *   n_obs <= 16, n_actions <= 16. Asynchronous on `stream`. */
FLOCK_API int flock_qnet_forward(const float *const *params, int recurrent, const float *obs, const float *hidden_in,
                       float *q_out, float *hidden_out, float *actions, int num_envs, int num_agents, int n_obs,
                       int n_actions, float epsilon, uint64_t seed, uint32_t step, int env_offset,
                       const flock_noise_counters_t *counters, void *stream);

/* The recurrent front ends of the two policies above ON THE TENSOR CORES (csrc/flock_gru_tc.cu): every layer is a
 * tcgen05.mma over tiles of 128 env rows with split fp16 operands (v = hi + lo, three MMAs per product: fp32-level
 * accuracy, hidden state and Q-values within 1e-5 of the fp32 modules), accumulators in TMEM, activations never leave
 * the SM. `mode` 0 = recurrent MADDPG actor front (fce + GRUCell, net.py:53-58), 1 = recurrent VDN QNet (net.py:27-58).
 *   flock_gru_tc_packed_bytes / flock_gru_tc_pack: packed parameter image; `params` as in flock_rnn_actor_forward
 *     (mode 0: the 6 front pointers) resp. flock_qnet_forward (mode 1: the 10 pointers). Call again after an update.
 *   flock_qnet_forward_tc: flock_qnet_forward (recurrent) with the packed image: same outputs, same exploration stream.
 *   flock_rnn_actor_forward_tc: flock_rnn_actor_forward[_ou] with the front end on the tensor cores (two launches:
 *     front, MLP); ou_state NULL = no exploration noise. */
FLOCK_API size_t flock_gru_tc_packed_bytes(int mode, int num_agents);
FLOCK_API int flock_gru_tc_pack(int mode, int num_agents, int n_obs, int n_actions, const float *const *params,
                                void *packed, void *stream);
FLOCK_API int flock_qnet_forward_tc(const void *packed, const float *obs, const float *hidden_in, float *q_out,
                                    float *hidden_out, float *actions, int num_envs, int num_agents, int n_obs,
                                    int n_actions, float epsilon, uint64_t seed, uint32_t step, int env_offset,
                                    const flock_noise_counters_t *counters, void *stream);
FLOCK_API int flock_rnn_actor_forward_tc(const void *packed_mlp, const void *packed_front, const float *obs,
                                         const float *hidden_in, float *hidden_out, float *actions, int num_envs,
                                         int num_agents, int n_obs, float *ou_state, float theta, float mu, float sigma,
                                         float dt, uint64_t seed, uint32_t step, int env_offset,
                                         const flock_noise_counters_t *counters, void *stream);

/* Debug / test hooks for the canonical arithmetic (device arrays, n elements). */
FLOCK_API int flock_debug_sincos(const float *h, int n, float *sn, float *cs, void *stream);
FLOCK_API int flock_debug_normal2(const uint32_t *words, int n_pairs, float *z, void *stream);
FLOCK_API int flock_debug_philox(const uint32_t *ctr4_key2, int n, uint32_t *out4, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOCK_B200_H */
