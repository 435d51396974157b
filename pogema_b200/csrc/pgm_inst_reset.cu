// pgm_inst_reset.cu - instantiates pgm_step_kernel<*, 0, OP_RESET, *> (see pgm_launch.cuh)
#include "pgm_launch.cuh"
namespace pgm {
int launch_reset(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_variant<0, OP_RESET>(d, a, s); }
}  // namespace pgm
