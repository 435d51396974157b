// pgm_inst_fast_soft_b.cu - instantiates pgm_fast_step_kernel<*, *, 2, radius group b> (see pgm_fast_launch.cuh)
#include "pgm_fast_launch.cuh"
namespace pgm {
int launch_fast_soft_b(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_fast_variant<2, 1>(d, a, s); }
}  // namespace pgm
