// pgm_inst_fast_block_both_a.cu - instantiates pgm_fast_step_kernel<*, *, 1, radius group a> (see pgm_fast_launch.cuh)
#include "pgm_fast_launch.cuh"
namespace pgm {
int launch_fast_block_both_a(const LaunchDims& d, const StepArgs& a, cudaStream_t s) { return launch_fast_variant<1, 0>(d, a, s); }
}  // namespace pgm
