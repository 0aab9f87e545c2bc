// flock_small_v2p.cu -- explicit instantiation of the small-path step kernels for one variant
// (V = FLOCK_V2, periodic metric = true, neighbour indices tracked = true); see flock_small_impl.cuh.
#include "flock_small_impl.cuh"

namespace flock {
template cudaError_t launch_step_small_vpi<FLOCK_V2, true, true>(const Params&, int, int, cudaStream_t);
}  // namespace flock
