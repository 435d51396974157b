// pgm_inst_fast_priority_b.cu - instantiates pgm_fast_step_kernel<*, *, 0, radius group b> (see pgm_fast_launch.cuh)
#include "pgm_fast_launch.cuh"
namespace pgm {
int launch_fast_priority_b(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_fast_variant<0, 1>(d, a, s); }
}  // namespace pgm
