// pgm_inst_step_priority_g1.cu - instantiates pgm_step_kernel<*, 0, OP_STEP, radius group 1, *> (see pgm_launch.cuh)
#include "pgm_launch.cuh"
namespace pgm {
int launch_step_priority_g1(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_variant<0, OP_STEP, 1>(d, a, s); }
}  // namespace pgm
