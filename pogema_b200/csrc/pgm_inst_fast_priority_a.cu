// pgm_inst_fast_priority_a.cu - instantiates pgm_fast_step_kernel<*, *, 0, radius group a> (see pgm_fast_launch.cuh)
#include "pgm_fast_launch.cuh"
namespace pgm {
int launch_fast_priority_a(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_fast_variant<0, 0>(d, a, s); }
}  // namespace pgm
