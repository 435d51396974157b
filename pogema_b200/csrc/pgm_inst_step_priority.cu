// pgm_inst_step_priority.cu - instantiates pgm_step_kernel<*, 0, OP_STEP, *> (see pgm_launch.cuh)
#include "pgm_launch.cuh"
namespace pgm {
int launch_step_priority(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_variant<0, OP_STEP>(d, a, s); }
}  // namespace pgm
