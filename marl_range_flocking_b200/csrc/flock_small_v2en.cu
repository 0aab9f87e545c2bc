// flock_small_v2en.cu -- explicit instantiation of the small-path step kernels for one variant
// (V = FLOCK_V2, periodic metric = false, neighbour indices tracked = false); see flock_small_impl.cuh.
#include "flock_small_impl.cuh"

namespace flock {
template cudaError_t launch_step_small_vpi<FLOCK_V2, false, false>(const Params&, int, int, cudaStream_t);
}  // namespace flock
