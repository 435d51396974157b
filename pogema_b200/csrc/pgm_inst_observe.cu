// pgm_inst_observe.cu - instantiates pgm_step_kernel<*, 0, OP_OBSERVE, *> (see pgm_launch.cuh)
#include "pgm_launch.cuh"
namespace pgm {
int launch_observe(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_variant<0, OP_OBSERVE>(d, a, s); }
}  // namespace pgm
