// pgm_inst_step_soft.cu - instantiates pgm_step_kernel<*, 2, OP_STEP, *> (see pgm_launch.cuh)
#include "pgm_launch.cuh"
namespace pgm {
int launch_step_soft(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_variant<2, OP_STEP>(d, a, s); }
}  // namespace pgm
