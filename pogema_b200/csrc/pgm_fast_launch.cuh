// pgm_fast_launch.cuh - instantiation + dispatch of pgm_fast_step_kernel<TEAM, APT, COLL, RT> (pgm_fast.cuh).
// One translation unit per (collision system, radius group) so that the variants compile in parallel.
#pragma once
#include "pgm_fast.cuh"
#include "pgm_launch.cuh"

namespace pgm {

template <int TEAM, int APT, int COLL, int RT>
int launch_fast_exact(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  static std::atomic<unsigned long long> configured{0ull};
  return launch_kernel(pgm_fast_step_kernel<TEAM, APT, COLL, RT>, configured, d, a, s);
}

template <int TEAM, int APT, int COLL, int RTG>
int launch_fast_rt(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  if constexpr (RTG == 0) {
    switch (d.rt) {
      case 2: return launch_fast_exact<TEAM, APT, COLL, 2>(d, a, s);
      case 3: return launch_fast_exact<TEAM, APT, COLL, 3>(d, a, s);
      default: return launch_fast_exact<TEAM, APT, COLL, 4>(d, a, s);
    }
  } else {
    switch (d.rt) {
      case 5: return launch_fast_exact<TEAM, APT, COLL, 5>(d, a, s);
      case 6: return launch_fast_exact<TEAM, APT, COLL, 6>(d, a, s);
      default: return launch_fast_exact<TEAM, APT, COLL, 7>(d, a, s);
    }
  }
}

template <int TEAM, int COLL, int RTG>
int launch_fast_apt(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  switch (d.apt) {
    case 1: return launch_fast_rt<TEAM, 1, COLL, RTG>(d, a, s);
    case 2: return launch_fast_rt<TEAM, 2, COLL, RTG>(d, a, s);
    default: return launch_fast_rt<TEAM, 4, COLL, RTG>(d, a, s);
  }
}

// (team, agents per thread) pairs the planner may choose: see pgm_plan.cu :: plan_fast
template <int COLL, int RTG>
int launch_fast_variant(const LaunchDims& d, const StepArgs& a, cudaStream_t s) {
  switch (d.team) {
    case 32: return launch_fast_apt<32, COLL, RTG>(d, a, s);
    case 64: return launch_fast_apt<64, COLL, RTG>(d, a, s);
    case 128: return launch_fast_apt<128, COLL, RTG>(d, a, s);
    default: return launch_fast_apt<256, COLL, RTG>(d, a, s);
  }
}

}  // namespace pgm
