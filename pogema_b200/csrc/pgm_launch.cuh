// pgm_launch.cuh - kernel instantiation + launch dispatch shared by the per-variant
// translation units (pgm_inst_*.cu).  Each TU instantiates one (COLL, OP) pair for
// every TEAM size and every compile-time radius, so the variants build in parallel.
#pragma once
#include <atomic>

#include "pgm_kernels.cuh"

namespace pgm {

struct LaunchDims {
  int team, rt, grid, block, smem, device, pdl, occ, og;
  int apt = 0;  // agents per thread of the fast kernel (pgm_fast.cuh)
};

// returns a cudaError_t as int
template <typename Kern>
int launch_kernel(Kern kern, std::atomic<unsigned long long>& configured, const LaunchDims& d, const StepArgs& a,
                  cudaStream_t s) {
  // The attribute is per (function, device) and process-wide: set it once per device to the hardware maximum, so
  // engines of different shapes driven from different threads never lower each other's limit.
  constexpr int kMaxSmem = 227 * 1024;
  const unsigned long long bit = d.device < 64 ? (1ull << d.device) : 0ull;  // devices >= 64: every launch
  if (!(configured.load(std::memory_order_acquire) & bit) || bit == 0ull) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e != cudaSuccess) return (int)e;
    configured.fetch_or(bit, std::memory_order_release);
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

template <int TEAM, int COLL, int OP, int RT, int OCC, int OG = 0>
int launch_exact(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  static std::atomic<unsigned long long> configured{0ull};
  return launch_kernel(pgm_step_kernel<TEAM, COLL, OP, RT, OCC, OG>, configured, d, a, s);
}

// Radii with a compile-time specialisation are split in two groups (RTG) so that the variants of one
// (COLL, OP) pair compile in two translation units: group 0 = generic + r 1..3, group 1 = r 4..7.
template <int TEAM, int COLL, int OP, int OCC, int RTG>
int launch_rt_occ(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  if constexpr (RTG == 0) {
    switch (d.rt) {
      case 1: return launch_exact<TEAM, COLL, OP, 1, OCC>(d, a, s);
      case 2: return launch_exact<TEAM, COLL, OP, 2, OCC>(d, a, s);
      case 3: return launch_exact<TEAM, COLL, OP, 3, OCC>(d, a, s);
      default: return launch_exact<TEAM, COLL, OP, 0, OCC>(d, a, s);
    }
  } else {
    switch (d.rt) {
      case 4: return launch_exact<TEAM, COLL, OP, 4, OCC>(d, a, s);
      case 5: return launch_exact<TEAM, COLL, OP, 5, OCC>(d, a, s);
      case 6: return launch_exact<TEAM, COLL, OP, 6, OCC>(d, a, s);
      default: return launch_exact<TEAM, COLL, OP, 7, OCC>(d, a, s);
    }
  }
}

template <int TEAM, int COLL, int OP, int RTG>
int launch_rt(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  // the occupancy structure only exists in the step path
  if constexpr (OP == OP_STEP) {
    if (d.occ == 1) return launch_rt_occ<TEAM, COLL, OP, 1, RTG>(d, a, s);
  }
  return launch_rt_occ<TEAM, COLL, OP, 0, RTG>(d, a, s);
}

// obstacle bitmap in global memory (huge maps): 1024-thread teams, tile buckets only
template <int COLL, int OP, int RTG>
int launch_og(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  constexpr int OCC = (OP == OP_STEP) ? 1 : 0;
  if constexpr (RTG == 0) {
    return launch_exact<1024, COLL, OP, 0, OCC, 1>(d, a, s);  // huge maps: generic radius path only
  } else {
    return launch_exact<1024, COLL, OP, 5, OCC, 1>(d, a, s);  // ... and the default radius 5
  }
}

template <int COLL, int OP, int RTG>
int launch_variant(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  if (d.og) return launch_og<COLL, OP, RTG>(d, a, s);
  switch (d.team) {
    case 32: return launch_rt<32, COLL, OP, RTG>(d, a, s);
    case 64: return launch_rt<64, COLL, OP, RTG>(d, a, s);
    case 128: return launch_rt<128, COLL, OP, RTG>(d, a, s);
    case 256: return launch_rt<256, COLL, OP, RTG>(d, a, s);
    case 512: return launch_rt<512, COLL, OP, RTG>(d, a, s);
    default: return launch_rt<1024, COLL, OP, RTG>(d, a, s);
  }
}

// defined in pgm_inst_*.cu (suffix = radius group)
int launch_step_priority_g0(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_priority_g1(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_block_both_g0(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_block_both_g1(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_soft_g0(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_step_soft_g1(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_observe_g0(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_observe_g1(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_reset_g0(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_reset_g1(const LaunchDims& d, const StepArgs& a, cudaStream_t s);

// pgm_fast_step_kernel variants, defined in pgm_inst_fast_*.cu (suffix = radius group: a = 2..4, b = 5..7)
int launch_fast_priority_a(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_fast_priority_b(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_fast_block_both_a(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_fast_block_both_b(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_fast_soft_a(const LaunchDims& d, const StepArgs& a, cudaStream_t s);
int launch_fast_soft_b(const LaunchDims& d, const StepArgs& a, cudaStream_t s);

// radii with a compile-time specialisation (1..7); every other radius runs the generic path
inline int static_radius(int r) { return (r >= 1 && r <= 7) ? r : 0; }
inline int radius_group(int rt) { return rt >= 4 ? 1 : 0; }

}  // namespace pgm
