// pgm_engine.h - the engine object behind the C-ABI (include/pgm_b200.h) and the pieces its three
// translation units share:
//   pgm_plan.cu       planner (shared-memory layouts, team / CTA geometry, fast-kernel eligibility) + kernel launch
//   pgm_transport.cu  host-buffer calls: pgm_step_host / pgm_observe_host, packed transport, staging buffers
//   pgm_capi.cu       everything else of the ABI: create / destroy, task generation, step / reset / observe with
//                     device pointers, state access, checkpoints, errors
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pgm_b200.h"
#include "pgm_devgen.h"
#include "pgm_gen.h"
#include "pgm_hostexpand.h"
#include "pgm_launch.cuh"

namespace pgm_impl {

// Sets the thread's pgm_last_error() text and returns `code` (never throws across the ABI).
int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return pgm_impl::fail(PGM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int obs_elem_size(int fmt) { return fmt == PGM_OBS_F32 ? 4 : (fmt == PGM_OBS_F16 ? 2 : 1); }
inline int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
inline int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p *= 2;
  return p;
}

}  // namespace pgm_impl

struct pgm_engine {
  pgm_config cfg{};
  int PH = 0, PW = 0, WPR = 0, D = 0;
  int obst_stride = 0;  // words
  int bits_per_agent = 0, stage_bpa = 0;
  int64_t obs_inst_stride = 0, obs_bytes = 0;
  int64_t cells_stride = 0;
  bool lifelong = false;
  bool tasks_ready = false;
  int sm_count = 148;
  // plan
  int team = 32, tpc = 1, cta_threads = 32, smem_cta = 0, grid = 0, batch_agents = 1, occ_mode = 0, obst_global = 0;
  int batch_single = 1;  // observation batch of single-step launches (pgm_step), <= batch_agents
  pgm::StepArgs layout{};  // offsets only
  // fast path (pgm_fast.cuh): chosen by plan_fast() for the common shapes; step launches use it, reset / observe /
  // odd caller pointers go through the generic kernel with the plan above
  bool fast = false;
  int f_team = 0, f_apt = 0, f_tpc = 1, f_cta_threads = 0, f_smem_cta = 0, f_grid = 0;
  pgm::StepArgs f_layout{};
  // device state
  uint32_t* d_obst = nullptr;
  uint2 *d_state = nullptr, *d_state0 = nullptr;  // see pgm_kernels.cuh: x | active<<15 | y<<16 , target
  uint8_t *d_was = nullptr, *d_done = nullptr;
  int32_t *d_elapsed = nullptr, *d_macc = nullptr, *d_mlast = nullptr;
  int32_t* d_solve = nullptr;  // [N][A][2], on_target == nothing only (StepArgs::solve)
  pgm::Pcg64 *d_rng = nullptr, *d_rng0 = nullptr;
  int32_t *d_cstart = nullptr, *d_csize = nullptr;
  uint32_t* d_cells = nullptr;
  int* d_err = nullptr;
  long long* d_debug = nullptr;  // caller-owned, see pgm_set_debug_buffer
  // step_host scratch
  uint8_t *d_act_h = nullptr, *d_obs_h = nullptr, *d_term_h = nullptr, *d_trunc_h = nullptr;
  float* d_rew_h = nullptr;
  int act_h_itemsize = 0;
  // d_obs_h / d_rew_h / d_term_h / d_trunc_h are parts of ONE device block (d_obs_h is its base); small results
  // (single instances behind the list API) come back with one copy into pinned staging and one synchronisation
  int64_t out_block_bytes = 0, off_rew = 0, off_term = 0, off_trunc = 0;
  uint8_t* h_small = nullptr;  // pinned: [block | state NA*8 | was NA | pad to 16 | actions NA*8], only if the block is <= kSmallBlock
  uint8_t* h_small_dev = nullptr;  // the same memory as the device sees it (zero-copy results of tiny engines)
  std::vector<uint2> h_state_tmp;
  // host mirrors
  std::vector<uint32_t> h_obst;
  bool h_obst_valid = true;  // false after a device-side generation (obstacles are read back on demand)
  // device generator buffers
  uint64_t* d_gen_seeds = nullptr;
  int* d_gen_fail = nullptr;
  int* d_gen_index = nullptr;
  uint8_t* d_gen_map = nullptr;
  int* d_gen_scratch = nullptr;
  long long gen_scratch_bytes = 0;
  // auto_reset == 2 (rebuild the task from a new seed when an episode ends)
  uint64_t* d_cur_seeds = nullptr;
  uint8_t* d_regen_flag = nullptr;
  int* d_regen_count = nullptr;
  double gen_density = -1.0;  // parameters of the last pgm_generate*, reused by the rebuilds
  bool gen_has_map = false;
  bool gen_explicit = false;
  int regen_slots = 0;
  int64_t launches = 0;
  bool use_pdl = true;
  bool serialize_next = false;  // the next launch follows a kernel that rewrote d_obst (device generator): it must not
                                // start its bulk copy of the obstacle bitmap before that kernel has completed
  // packed host transport (pgm_step_host / pgm_observe_host): device bit stream -> pinned staging -> host threads
  int host_transport = -1;      // -1 auto, 0 plain (DMA of the final tensor), 1 packed
  int host_threads = 0;         // 0 = hardware concurrency (at most 32)
  int64_t stream_unit_bytes = 0, stream_batch_bytes = 0, stream_bytes = 0;
  uint8_t* d_stream = nullptr;  // device: [N][batches][stream_batch_bytes]
  uint8_t* h_stream = nullptr;  // pinned host copy
  uint32_t* d_flags = nullptr;  // device: [chunks] the step's epoch byte, copied to h_flags[c] right after chunk c
  uint32_t* h_flags = nullptr;  // pinned: polled by the host threads
  uint32_t epoch = 0;
  pgm::ExpandPool* pool = nullptr;
  bool ovr_stream = false;      // make_args: write the raw stream instead of cfg.obs_format
  int64_t last_d2h_bytes = 0, last_h2d_bytes = 0;
  int64_t last_us[5] = {0, 0, 0, 0, 0};  // packed pgm_step_host: enqueue done, first chunk landed, last chunk landed, widening done, stream idle
  std::chrono::steady_clock::time_point t_call;
  int stream_chunks = 8;
  cudaStream_t expand_stream = nullptr;  // stream of the packed host call in flight (stream_failed)
};

namespace pgm_impl {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- pgm_plan.cu
int compute_plan(pgm_engine* e);                                               // fills the plan fields of *e
pgm::StepArgs make_args(pgm_engine* e);                                        // kernel arguments without I/O pointers
int launch(pgm_engine* e, const pgm::StepArgs& a, int op, cudaStream_t s);     // one kernel launch (generic or fast)

// ---- pgm_transport.cu
int ensure_host_scratch(pgm_engine* e, int itemsize);
void free_transport(pgm_engine* e);                                            // staging buffers + widening pool

}  // namespace pgm_impl
