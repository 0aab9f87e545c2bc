// flock_small_uw.cu -- explicit instantiation of the small-path step kernels for one variant
// (V = FLOCK_UW, periodic metric = false, neighbour indices tracked = true); see flock_small_impl.cuh.
#include "flock_small_impl.cuh"

namespace flock {
template cudaError_t launch_step_small_vpi<FLOCK_UW, false, true>(const Params&, int, int, cudaStream_t);
}  // namespace flock
