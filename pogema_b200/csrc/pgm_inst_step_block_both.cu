// pgm_inst_step_block_both.cu - instantiates pgm_step_kernel<*, 1, OP_STEP, *> (see pgm_launch.cuh)
#include "pgm_launch.cuh"
namespace pgm {
int launch_step_block_both(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_variant<1, OP_STEP>(d, a, s); }
}  // namespace pgm
