// flock_launch.h -- host-side launch entry points of the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

namespace flock {
struct Params;

// Per-device one-time kernel configuration (dynamic shared memory opt-in, SM count): a process may drive
// several GPUs (VecEnv / the policies take device=...), and a function attribute belongs to the device
// that was current when it was set.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
    bool done[kMaxDevices] = {};
    int sm_count[kMaxDevices] = {};
    // returns the current device's SM count; runs `configure` the first time this device is seen
    template <typename F>
    cudaError_t get(F configure, int* sms) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDevices) dev = kMaxDevices - 1, done[dev] = false;
        if (!done[dev]) {
            e = configure();
            if (e != cudaSuccess) return e;
            int n = 0;
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            sm_count[dev] = n > 0 ? n : 148;
            done[dev] = true;
        }
        *sms = sm_count[dev];
        return cudaSuccess;
    }
};

// Launch `kernel` as a programmatic dependent of the previous kernel of the stream (FLOCK_PDL=0 disables): it may start
// while that kernel drains and must execute griddepcontrol.wait before it reads anything an earlier kernel wrote.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    static const bool use_pdl = [] {
        const char* v = getenv("FLOCK_PDL");
        return v == nullptr || v[0] != '0';
    }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// device-side step counters of the fused exploration noise (flock_noise_counters_t)
struct NoiseCounters {
    const int32_t* env_step = nullptr;
    const uint32_t* env_epoch = nullptr;
};

// flock_small.cu (N <= 32); kernel modes of flock_small_impl.cuh as seen by the API
enum : int { kSmallModeStep = 0, kSmallModeMirror = 1, kSmallModeAutoReset = 2, kSmallModeMulti = 3, kSmallModeRollout = 4,
             kSmallModeNoise = 5 };
// mode: the kernel modes of flock_small_impl.cuh (0 step, 1 host mirror, 2 fused auto-reset, 3 step_n, 4 rollout_n, 5 fused sensing noise)
cudaError_t launch_step_small(int variant, bool periodic, const Params& p, int mode, int sm_count, cudaStream_t s);
cudaError_t launch_newest_row(const Params& p, float* out, int sm_count, cudaStream_t s);    // [E][N][k] newest range row
cudaError_t launch_obs_window(const Params& p, float* out, int sm_count, cudaStream_t s);   // uw ring -> newest-first window
cudaError_t launch_reset_small(const Params& p, int sm_count, cudaStream_t s);
cudaError_t launch_random_actions(int variant, const Params& p, float* out, int sm_count, cudaStream_t s);
cudaError_t launch_range_noise(const Params& p, int sm_count, cudaStream_t s);
cudaError_t launch_debug_sincos(const float* h, int n, float* sn, float* cs, cudaStream_t s);
cudaError_t launch_debug_normal2(const uint32_t* w, int n, float* z, cudaStream_t s);
cudaError_t launch_debug_philox(const uint32_t* ck, int n, uint32_t* out, cudaStream_t s);

// flock_tiled.cu (N > 32)
cudaError_t tiled_configure(int num_agents);   // opt in to the dynamic shared memory the env needs
size_t tiled_smem_bytes(int num_agents, int rows, bool need_sh);
cudaError_t launch_step_tiled(int variant, bool periodic, const Params& p, int sm_count, int tiled_mode, cudaStream_t s);
cudaError_t launch_reset_tiled(const Params& p, uint8_t* need, int* launches, cudaStream_t s);   // need: [E] device scratch
cudaError_t launch_perm_refresh(const Params& p, int* perm, int* inv, cudaStream_t s);   // spatial row order
size_t pruned_scratch_floats(int N, int E);
size_t pruned_hint_bytes(int N, int E);
int tiled_step_launches(int variant, const Params& p, int sm_count, int tiled_mode);   // kernels per tiled step
cudaError_t launch_perm_identity(int* perm, int* inv, int N, int E, cudaStream_t s);
bool tiled_uses_row_order(const Params& p, int sm_count, int tiled_mode);

// flock_actor.cu (fused per-agent actor MLP on the tensor cores)
size_t actor_blob_bytes();
int actor_max_in_dims();
void actor_dims(int* fc1, int* fc2, int* n_actions);
cudaError_t launch_actor_pack(int agents, int in_dims, const float* const* ptrs, void* blobs, cudaStream_t s);
cudaError_t launch_actor_forward(const void* blobs, const float* obs, float* actions, int E, int N, int in_dims,
                                 float* ou_state, float ou_theta, float ou_mu, float ou_sigma, float ou_dt, uint64_t seed,
                                 uint32_t step, int env_offset, NoiseCounters ctr, const int32_t* obs_head, int ring_h, int ring_k,
                                 cudaStream_t s);

// flock_qnet.cu (fused VDN Q-network forward + epsilon-greedy action selection, fp32)
int qnet_max_obs();
int qnet_max_actions();
cudaError_t launch_qnet(const float* const* params, int recurrent, const float* obs, const float* hidden_in, float* q_out,
                        float* hidden_out, float* actions, int E, int A, int n_obs, int n_act, float epsilon, uint64_t seed,
                        uint32_t step, int env_offset, NoiseCounters ctr, cudaStream_t s);

// flock_rnn_actor.cu (fused recurrent MADDPG actor: fp32 GRU front end + tensor-core MLP)
size_t rnn_actor_blob_bytes();
int rnn_actor_max_obs();
cudaError_t launch_rnn_actor_pack(int agents, const float* const* ptrs, void* blobs, cudaStream_t s);
cudaError_t launch_rnn_actor_forward(const void* blobs, const float* const* front, const float* obs, const float* hidden_in,
                                     float* hidden_out, float* actions, int E, int N, int n_obs, float* ou_state,
                                     float ou_theta, float ou_mu, float ou_sigma, float ou_dt, uint64_t seed, uint32_t step,
                                     int env_offset, NoiseCounters ctr, cudaStream_t s);

// flock_gru_tc.cu (the recurrent front ends on the tensor cores: mode 0 = recurrent MADDPG actor front, 1 = VDN QNet)
size_t gru_tc_blob_bytes(int mode);
int gru_tc_max_obs();
int gru_tc_max_actions();
cudaError_t launch_gru_tc_pack(int mode, int agents, int n_obs, int n_act, const float* const* params, void* blobs, cudaStream_t s);
cudaError_t launch_gru_tc_forward(int mode, const void* blobs, const float* obs, const float* hidden_in, float* hidden_out,
                                  float* q_out, float* actions, int E, int A, int n_obs, int n_act, float epsilon, uint64_t seed,
                                  uint32_t step, int env_offset, NoiseCounters ctr, cudaStream_t s);
}  // namespace flock
