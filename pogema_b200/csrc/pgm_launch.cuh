// pgm_launch.cuh - kernel instantiation + launch dispatch shared by the per-variant
// translation units (pgm_inst_*.cu).  Each TU instantiates one (COLL, OP) pair for
// every TEAM size and every compile-time radius, so the variants build in parallel.
#pragma once
#include "pgm_kernels.cuh"

namespace pgm {

struct LaunchDims {
  int team, rt, grid, block, smem, device, pdl, occ, og;
};

// returns a cudaError_t as int
template <int TEAM, int COLL, int OP, int RT, int OCC, int OG = 0>
int launch_exact(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  auto kern = pgm_step_kernel<TEAM, COLL, OP, RT, OCC, OG>;
  static thread_local int configured_dev = -1;
  static thread_local int configured_smem = -1;
  if (configured_dev != d.device || configured_smem < d.smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem);
    if (e != cudaSuccess) return (int)e;
    configured_dev = d.device;
    configured_smem = d.smem;
  }
  // programmatic dependent launch: this grid may be scheduled while the previous kernel of
  // the stream drains; the kernel's griddepcontrol.wait orders every read of mutable state.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(d.grid);
  cfg.blockDim = dim3(d.block);
  cfg.dynamicSmemBytes = d.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = d.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, a);
}

template <int TEAM, int COLL, int OP, int OCC>
int launch_rt_occ(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  switch (d.rt) {
    case 3: return launch_exact<TEAM, COLL, OP, 3, OCC>(d, a, s);
    case 5: return launch_exact<TEAM, COLL, OP, 5, OCC>(d, a, s);
    case 7: return launch_exact<TEAM, COLL, OP, 7, OCC>(d, a, s);
    default: return launch_exact<TEAM, COLL, OP, 0, OCC>(d, a, s);
  }
}

template <int TEAM, int COLL, int OP>
int launch_rt(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  // the occupancy structure only exists in the step path
  if constexpr (OP == OP_STEP) {
    if (d.occ == 1) return launch_rt_occ<TEAM, COLL, OP, 1>(d, a, s);
  }
  return launch_rt_occ<TEAM, COLL, OP, 0>(d, a, s);
}

// obstacle bitmap in global memory (huge maps): 1024-thread teams, tile buckets only
template <int COLL, int OP>
int launch_og(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  constexpr int OCC = (OP == OP_STEP) ? 1 : 0;
  switch (d.rt) {
    case 3: return launch_exact<1024, COLL, OP, 3, OCC, 1>(d, a, s);
    case 5: return launch_exact<1024, COLL, OP, 5, OCC, 1>(d, a, s);
    case 7: return launch_exact<1024, COLL, OP, 7, OCC, 1>(d, a, s);
    default: return launch_exact<1024, COLL, OP, 0, OCC, 1>(d, a, s);
  }
}

template <int COLL, int OP>
int launch_variant(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  if (d.og) return launch_og<COLL, OP>(d, a, s);
  switch (d.team) {
    case 32: return launch_rt<32, COLL, OP>(d, a, s);
    case 64: return launch_rt<64, COLL, OP>(d, a, s);
    case 128: return launch_rt<128, COLL, OP>(d, a, s);
    case 256: return launch_rt<256, COLL, OP>(d, a, s);
    case 512: return launch_rt<512, COLL, OP>(d, a, s);
    default: return launch_rt<1024, COLL, OP>(d, a, s);
  }
}

// defined in pgm_inst_*.cu
int launch_step_priority(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_block_both(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_soft(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_observe(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_reset(const LaunchDims& d, const StepArgs& a, cudaStream_t s);

// radii with a compile-time specialisation
inline int static_radius(int r) { return (r == 3 || r == 5 || r == 7) ? r : 0; }

}  // namespace pgm
