// pgm_devgen.h - device-side task generation (see pgm_devgen.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pgm_rng.h"

namespace pgm {

struct DevGenArgs {
  int first, count;        // instances [first, first+count) unless `index` is given
  const int* index;        // optional device array [count] of instance indices (regeneration mode)
  const int* count_ptr;    // regeneration mode: number of entries of `index` (device scalar)
  const uint64_t* seeds;   // device [count] (direct mode)
  const uint64_t* cur_seeds;  // regeneration mode: seed of instance n = cur_seeds[n]
  int slots;               // scratch slots = warps of the launch; warp w handles entries w, w+slots, ...
  int smem_labels;         // set by launch_devgen: component labelling runs in shared memory (small maps)
  int* err_flag;           // regeneration mode: sticky engine error flag (bit 2: task could not be rebuilt)
  int H, W, A, r, lifelong;
  const uint8_t* map;      // optional device map [H][W] shared by all instances
  // numpy random_binomial(p, n=1) constants computed on the host (libm exp/log)
  int binom_zero, binom_flip;
  double binom_qn, binom_px1;
  int* scratch;            // devgen_scratch_bytes(...)
  int* fail;               // device [count], zero on entry; != 0 -> regenerate on the host
  // engine arrays
  uint32_t* obst;
  int obst_stride;
  uint2* state;
  uint2* state0;
  int32_t* elapsed;
  uint8_t* episode_done;
  uint8_t* was_on_goal;
  int32_t* metric_acc;
  int32_t* metric_last;
  int32_t* solve;  // null unless on_target == nothing (pgm_kernels.cuh :: StepArgs::solve)
  Pcg64* rng;
  Pcg64* rng0;
  int32_t* comp_start;
  int32_t* comp_size;
  uint32_t* cells;
  long long cells_stride;
};

long long devgen_scratch_bytes(int HW, int A, int count);
int launch_devgen(const DevGenArgs& a, cudaStream_t s);  // returns a cudaError_t
// flags[n] != 0 -> index[count++] = n, cur_seeds[n] += stride
int launch_regen_compact(int N, const uint8_t* flags, uint64_t* cur_seeds, uint64_t stride, int* index, int* count,
                         cudaStream_t s);

}  // namespace pgm
