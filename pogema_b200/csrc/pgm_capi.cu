// pgm_capi.cu - engine object + the C-ABI declared in include/pgm_b200.h.
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

using namespace pgm;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(PGM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
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

}  // namespace

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
  StepArgs layout{};  // offsets only
  // fast path (pgm_fast.cuh): chosen by plan_fast() for the common shapes; step launches use it, reset / observe /
  // odd caller pointers go through the generic kernel with the plan above
  bool fast = false;
  int f_team = 0, f_apt = 0, f_tpc = 1, f_cta_threads = 0, f_smem_cta = 0, f_grid = 0;
  StepArgs f_layout{};
  uint8_t* d_fast_fill = nullptr;  // constant template the fast kernel's prologue copies into shared memory (TMA engine)
  int fast_fill_bytes = 0;
  int stagger_ns = 0;              // tuning knob PGM_STAGGER_NS (single-step launches of the fast kernel)
  // device state
  uint32_t* d_obst = nullptr;
  uint2 *d_state = nullptr, *d_state0 = nullptr;  // see pgm_kernels.cuh: x | active<<15 | y<<16 , target
  uint8_t *d_was = nullptr, *d_done = nullptr;
  int32_t *d_elapsed = nullptr, *d_macc = nullptr, *d_mlast = nullptr;
  Pcg64 *d_rng = nullptr, *d_rng0 = nullptr;
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

namespace {

// Shared-memory layout of one instance for a given occupancy structure and observation batch.
struct Layout {
  StepArgs L{};
  int occ_mode = 0, batch_agents = 0, team_smem = 0, obst_global = 0;
};

bool make_layout(const pgm_engine* e, int occ_mode, int want_resident, Layout* out, bool obst_global = false,
                 int force_batch = 0) {
  const int A = e->cfg.num_agents;
  const int smem_max = 227 * 1024;
  int tiles = 0, tiles_w = 0, tshift = 0, occ_bytes;
  if (occ_mode == 0) {
    occ_bytes = round_up(e->PH * e->PW * 2 + 4, 16);
  } else {
    // tile buckets: 4x4 tiles, coarser while the head array is larger than 16 KB
    tshift = 2;
    for (;;) {
      tiles_w = (e->PW + (1 << tshift) - 1) >> tshift;
      tiles = round_up(((e->PH + (1 << tshift) - 1) >> tshift) * tiles_w, 4);
      if (tiles * 4 <= 16 * 1024 || tshift >= 6) break;
      tshift++;
    }
    occ_bytes = round_up(tiles * 4 + A * 2, 16);
  }
  const int bitmap_bytes = round_up((e->PH * e->WPR + 1) * 4, 16);
  // obst_global: no staged obstacle bitmap, and the pre-move bitmap aliases the post-move one
  const int fixed = (obst_global ? 0 : e->obst_stride * 4) + bitmap_bytes * ((occ_mode == 1 && !obst_global) ? 2 : 1) +
                    4 * round_up(A * 4, 16) + 2 * round_up(A, 16) + 16;
  if (fixed + occ_bytes > smem_max) return false;
  // observation stage: aliases the occupancy region, so at least that much is free; beyond it take what
  // still lets `want_resident` instances share an SM, but never less than 32 agents (or all of them)
  const long long per_agent_bits = e->stage_bpa;
  auto stage_bytes_for = [&](long long g) { return (long long)round_up((int)(((g * per_agent_bits + 31) / 32 + 2) * 4), 16); };
  // an SM has 228 KB; every resident CTA costs 1 KB of it on top of its own allocation
  const long long target = (228 * 1024) / std::max(1, want_resident) - 1024;
  long long budget = std::max<long long>(occ_bytes, target - fixed);
  budget = std::max<long long>(budget, stage_bytes_for(std::min(A, 32)));
  budget = std::min<long long>(budget, (long long)smem_max - fixed);
  budget = std::min<long long>(budget, std::max<long long>(stage_bytes_for(A), occ_bytes));
  long long g = (budget >= stage_bytes_for(A)) ? A : ((budget - 16) * 8) / per_agent_bits;
  if (g < 1) return false;
  g = std::min<long long>(g, A);
  if (g < A && g > 32) g = g / 32 * 32;  // whole warps of agents per batch
  if (const char* v = getenv("PGM_OBS_BATCH")) g = std::max<long long>(1, std::min<long long>(g, atoi(v)));  // tuning knob
  if (force_batch > 0) {
    // the fast step kernel writes the packed stream in batches of its team size: the generic observe / reset
    // launches of the same engine must use the same batch
    g = std::min<long long>(A, force_batch);
    if (fixed + std::max<long long>(occ_bytes, stage_bytes_for(g)) > smem_max) return false;
  }
  const int stage_bytes = (int)stage_bytes_for(g);
  StepArgs& L = out->L;
  int off = 0;
  L.off_obst = off;
  if (!obst_global) off += e->obst_stride * 4;
  L.off_abits = off;
  off += bitmap_bytes;
  L.off_pbits = obst_global ? L.off_abits : off;
  if (occ_mode == 1 && !obst_global) off += bitmap_bytes;
  L.off_occ = off;
  off += std::max(occ_bytes, stage_bytes);
  L.off_pos = off;
  off += round_up(A * 4, 16);
  L.off_tgt = off;
  off += round_up(A * 4, 16);
  L.off_npos = off;
  off += round_up(A * 4, 16);
  L.off_link = off;
  off += round_up(A * 4, 16);
  L.off_act = off;
  off += round_up(A, 16);
  L.off_flag = off;
  off += round_up(A, 16);
  L.off_misc = off;
  off += 16;
  L.team_smem = round_up(off, 16);
  L.occ_tiles = tiles;
  L.occ_tiles_w = tiles_w;
  L.occ_tshift = tshift;
  if (L.team_smem > smem_max) return false;
  out->occ_mode = occ_mode;
  out->obst_global = obst_global ? 1 : 0;
  out->batch_agents = (int)g;
  out->team_smem = L.team_smem;
  return true;
}

// teams per CTA: balance the busiest SM (CTAs are dealt round-robin, every SM should host the same
// number of instances), prefer CTAs of 192..512 threads (measured: smaller CTAs cost ~15 %)
int choose_tpc(const pgm_engine* e, int team, int team_smem) {
  const int smem_max = 227 * 1024;
  const pgm_config& c = e->cfg;
  int max_tpc = std::min(1024 / team, std::max(1, smem_max / team_smem));
  if (team > 32) max_tpc = std::min(max_tpc, 15);
  int tpc = 1;
  double best = -1.0;
  const double ideal = (double)c.num_envs / e->sm_count;
  for (int t = 1; t <= max_tpc; ++t) {
    const int grid = (c.num_envs + t - 1) / t;
    const int per_sm_ctas = (grid + e->sm_count - 1) / e->sm_count;
    const double busiest = (double)per_sm_ctas * t;
    double score = ideal / busiest;
    const int threads = t * team;
    if (threads < 192) score *= 0.85;
    if (threads > 512) score *= 0.95;
    score -= 1e-4 * std::abs(threads - 256) / 256.0;  // tie-break: closest to 256 threads
    if (score > best) {
      best = score;
      tpc = t;
    }
  }
  if (const char* v = getenv("PGM_TPC")) tpc = std::max(1, std::min(max_tpc, atoi(v)));  // tuning knob
  return tpc;
}

// The fast step kernel (pgm_fast.cuh) for the common shapes: compile-time radius 2..7, uint8 / bits observations
// whose per-instance block is a multiple of 16 bytes, at most 4 agents per thread, at most 8190 agents, both bitmaps
// (and for priority / soft the uint16 cell grid) in shared memory at the residency the job wants.
bool plan_fast(pgm_engine* e, int team, int want) {
  const pgm_config& c = e->cfg;
  const int A = c.num_agents;
  if (const char* v = getenv("PGM_FAST")) {
    if (v[0] == '0') return false;
  }
  if (c.obs_radius < 2 || c.obs_radius > 7) return false;
  if (c.obs_format != PGM_OBS_U8 && c.obs_format != PGM_OBS_BITS) return false;
  if (c.obs_format == PGM_OBS_U8 && ((int64_t)A * e->bits_per_agent) % 16 != 0) return false;
  if (A > 8190 || e->obst_global) return false;
  team = std::max(32, std::min(team, 256));
  while ((A + team - 1) / team > 4 && team < 256) team *= 2;
  int apt = (A + team - 1) / team;
  if (apt > 4) return false;
  if (apt == 3) apt = 4;
  if (const char* v = getenv("PGM_FAST_TEAM")) {  // tuning knob
    const int t = atoi(v);
    if ((t == 32 || t == 64 || t == 128 || t == 256) && (A + t - 1) / t <= 4) {
      team = t;
      apt = (A + t - 1) / t;
      if (apt == 3) apt = 4;
    }
  }
  const int smem_max = 227 * 1024;
  const int bitmap_bytes = round_up((e->PH * e->WPR + 1) * 4, 16);
  const bool bb = c.collision_system == PGM_COLLISION_BLOCK_BOTH;
  const int stage_one = round_up((team * e->stage_bpa + 31) / 32 * 4 + 16, 16);
  auto build = [&](int bufs, StepArgs* L) {
    int off = 0;
    L->off_obst = off;
    off += e->obst_stride * 4;
    L->off_abits = off;
    off += bitmap_bytes;
    L->off_pbits = off;  // block_both: the second agent bitmap
    if (bb) off += bitmap_bytes;
    L->off_occ = off;
    L->off_stage = off;  // block_both: the stream buffers lie over the claim planes (zeroed at the start of a step)
    if (bb) {
      off += std::max(2 * bitmap_bytes, bufs * stage_one);
    } else {
      off += round_up(e->PH * e->PW * 2, 16);
      L->off_stage = off;
      off += bufs * stage_one;
    }
    L->off_link = off;
    if (!bb) off += round_up(apt * team * 4, 16);
    L->off_npos = off;
    if (!bb) off += round_up(apt * team * 4, 16);
    L->off_misc = off;
    off += 16;
    L->team_smem = round_up(off, 16);
    L->stage_bufs = bufs;
    L->stage_words = stage_one / 4;
    L->plane_words = bitmap_bytes / 4;
    L->narrow = (e->WPR == 2 && c.width <= 32) ? 1 : 0;
    return L->team_smem;
  };
  auto fit = [](int team_smem) { return (228 * 1024) / (team_smem + 1024); };
  StepArgs L{};
  int bufs = apt > 1 ? 2 : 1;
  if (const char* v = getenv("PGM_FAST_BUFS")) bufs = atoi(v) > 1 ? 2 : 1;  // tuning knob
  int sm = build(bufs, &L);
  const int need = std::min(want, std::max(1, 1024 / team));
  if (bufs == 2 && (sm > smem_max || fit(sm) < need)) sm = build(1, &L);
  if (sm > smem_max) return false;
  if (fit(sm) < std::min(need, 2) && want > 1) return false;  // the generic kernel's leaner layouts keep more instances resident
  e->f_layout = L;
  e->f_team = team;
  e->f_apt = apt;
  return true;
}

int compute_plan(pgm_engine* e) {
  const pgm_config& c = e->cfg;
  const int A = c.num_agents;
  const int smem_max = 227 * 1024;
  int per_sm = (c.num_envs + e->sm_count - 1) / e->sm_count;  // instances an SM has to host
  if (const char* v = getenv("PGM_RESIDENT")) per_sm = std::max(1, atoi(v));  // tuning knob
  // Instances an SM should host at a time: what the job needs, but not so many that a team drops
  // below a quarter of a thread per agent (measured on 512 instances of 1024 agents, 256x256 map: 4 x 256
  // threads with 128-agent observation batches 42.7 us per step, 2 x 512 threads 48.0, 1 x 1024 56.6 -
  // four teams per SM interleave their move and store phases, two mostly alternate; 512-agent instances
  // keep 4 x 256: 7 x 128 threads gain 4 % with 16 steps per launch but lose 16 % with one).  The dense cell->agent grid is used when it reaches that residency (one LDS per
  // lookup), otherwise the tile buckets (memory ~ agents instead of cells).
  int want = std::max(1, std::min(per_sm, std::max(A <= 1024 ? 4 : 2, 2048 / pow2_ceil(A))));
  if (const char* v = getenv("PGM_WANT")) want = std::max(1, atoi(v));  // tuning knob
  Layout dense, buckets, *use = nullptr;
  const bool ok_d = make_layout(e, 0, want, &dense);
  const bool ok_h = make_layout(e, 1, want, &buckets);
  auto fit = [](int team_smem) { return (228 * 1024) / (team_smem + 1024); };  // 1 KB per resident CTA is reserved
  const int res_d = ok_d ? std::min(want, fit(dense.team_smem)) : 0;
  const int res_h = ok_h ? std::min(want, fit(buckets.team_smem)) : 0;
  int force = -1;
  if (const char* v = getenv("PGM_OCC")) force = atoi(v);  // tuning knob: 0 dense grid, 1 tile buckets
  if (force == 0 && ok_d) use = &dense;
  else if (force == 1 && ok_h) use = &buckets;
  else if (ok_d && res_d >= res_h) use = &dense;
  else if (ok_h) use = &buckets;
  Layout huge;
  if (!use && make_layout(e, 1, 1, &huge, true)) use = &huge;  // bitmaps too large: obstacles stay in global memory
  if (!use)
    return fail(PGM_ERR_UNSUPPORTED,
                "one instance does not fit in 227 KB of shared memory: map %dx%d (padded %dx%d), %d agents, r=%d",
                c.height, c.width, e->PH, e->PW, A, c.obs_radius);
  e->layout = use->L;
  e->occ_mode = use->occ_mode;
  e->obst_global = use->obst_global;
  e->batch_agents = use->batch_agents;
  StepArgs& L = e->layout;
  int team = use->obst_global ? 1024 : c.team_threads;
  if (team == 0) {
    // ~1024 threads per SM (64 registers each) shared by the instances an SM hosts at a time
    const int resident = std::max(1, std::min(want, fit(L.team_smem)));
    team = pow2_floor(std::max(32, 1024 / resident));
    team = std::min(team, std::max(32, pow2_ceil(A)));
    team = std::min(team, 1024);
  }
  if (team != 32 && team != 64 && team != 128 && team != 256 && team != 512 && team != 1024)
    return fail(PGM_ERR_INVALID, "team_threads must be 0 or a power of two in [32,1024], got %d", team);
  e->team = team;
  // Single-step launches (pgm_step, closed loop): all teams reach the store phase together, so splitting the
  // observation phase in two lets the first half's stores drain under the second half's bit assembly
  // (measured: configs[1] 22.6 -> 21.7 us, configs[2] 26.6 -> 24.7 us per step; 512-thread teams lose).
  // Teams of 64 / 128 threads do the same in multi-step launches (configs[2], 128 threads x 256 agents: 18.8 -> 18.1 us
  // per step with 16 steps per launch); single warps lose 1 % there and keep one batch.
  if (!getenv("PGM_OBS_BATCH") && e->batch_agents == A && team >= 64 && team <= 128 && A >= 2 * team)
    e->batch_agents = std::max(team, A / 2);
  e->batch_single = e->batch_agents;
  if (!getenv("PGM_OBS_BATCH") && e->batch_agents == A && team <= 128 && A >= 2 * team) e->batch_single = std::max(team, A / 2);
  // the fast step kernel, if this shape has one; the generic launches of the engine then use its batch size
  e->fast = false;
  if (plan_fast(e, team, want)) {
    Layout forced;
    if (make_layout(e, use->occ_mode, want, &forced, use->obst_global != 0, e->f_team)) {
      e->layout = forced.L;
      e->batch_agents = forced.batch_agents;
      e->batch_single = forced.batch_agents;
      e->fast = true;
      e->f_tpc = choose_tpc(e, e->f_team, e->f_layout.team_smem);
      e->f_layout.teams_per_cta = e->f_tpc;
      e->f_cta_threads = e->f_tpc * e->f_team;
      e->f_smem_cta = e->f_tpc * e->f_layout.team_smem;
      e->f_grid = (c.num_envs + e->f_tpc - 1) / e->f_tpc;
    }
  }
  const int tpc = choose_tpc(e, team, L.team_smem);
  e->tpc = tpc;
  L.teams_per_cta = tpc;
  e->cta_threads = tpc * team;
  e->smem_cta = tpc * L.team_smem;
  e->grid = (c.num_envs + tpc - 1) / tpc;
  return PGM_OK;
}

int launch(pgm_engine* e, const StepArgs& a, int op, cudaStream_t s) {
  LaunchDims d{e->team, static_radius(e->cfg.obs_radius), e->grid, e->cta_threads, e->smem_cta, e->cfg.device,
               (e->use_pdl && !e->serialize_next) ? 1 : 0, e->occ_mode, e->obst_global};
  e->serialize_next = false;
  // huge maps (obstacles in global memory) only have the generic and the r=5 variants
  if (d.og && d.rt != 5) d.rt = 0;
  const int g = d.og ? (d.rt == 5 ? 1 : 0) : radius_group(d.rt);
  int err;
  // step launches of the common shapes: the register-resident kernel (uint8 observation blocks must be 16-byte aligned)
  const bool fast = op == OP_STEP && e->fast &&
                    (a.obs == nullptr || a.obs_format != 0 ||
                     ((reinterpret_cast<uintptr_t>(a.obs) & 15u) == 0 && (a.obs_slot_stride & 15) == 0));
  if (fast) {
    StepArgs f = a;
    const StepArgs& L = e->f_layout;
    f.off_obst = L.off_obst;
    f.off_abits = L.off_abits;
    f.off_pbits = L.off_pbits;
    f.off_occ = L.off_occ;
    f.off_stage = L.off_stage;
    f.off_link = L.off_link;
    f.off_npos = L.off_npos;
    f.off_misc = L.off_misc;
    f.team_smem = L.team_smem;
    f.teams_per_cta = L.teams_per_cta;
    f.stage_bufs = L.stage_bufs;
    f.stage_words = L.stage_words;
    f.plane_words = L.plane_words;
    f.narrow = L.narrow;
    f.fill_src = e->d_fast_fill;
    f.fill_bytes = e->fast_fill_bytes;
    f.stagger_ns = a.num_steps == 1 ? e->stagger_ns : 0;
    d.team = e->f_team;
    d.apt = e->f_apt;
    d.grid = e->f_grid;
    d.block = e->f_cta_threads;
    d.smem = e->f_smem_cta;
    const int fg = d.rt >= 5 ? 1 : 0;
    if (e->cfg.collision_system == PGM_COLLISION_PRIORITY)
      err = fg ? launch_fast_priority_b(d, f, s) : launch_fast_priority_a(d, f, s);
    else if (e->cfg.collision_system == PGM_COLLISION_BLOCK_BOTH)
      err = fg ? launch_fast_block_both_b(d, f, s) : launch_fast_block_both_a(d, f, s);
    else
      err = fg ? launch_fast_soft_b(d, f, s) : launch_fast_soft_a(d, f, s);
  } else if (op == OP_OBSERVE) err = g ? launch_observe_g1(d, a, s) : launch_observe_g0(d, a, s);
  else if (op == OP_RESET) err = g ? launch_reset_g1(d, a, s) : launch_reset_g0(d, a, s);
  else if (e->cfg.collision_system == PGM_COLLISION_PRIORITY)
    err = g ? launch_step_priority_g1(d, a, s) : launch_step_priority_g0(d, a, s);
  else if (e->cfg.collision_system == PGM_COLLISION_BLOCK_BOTH)
    err = g ? launch_step_block_both_g1(d, a, s) : launch_step_block_both_g0(d, a, s);
  else
    err = g ? launch_step_soft_g1(d, a, s) : launch_step_soft_g0(d, a, s);
  if (err != 0)
    return fail(PGM_ERR_CUDA, "kernel launch failed: %s (grid %d, block %d, smem %d)",
                cudaGetErrorString((cudaError_t)err), d.grid, d.block, d.smem);
  e->launches++;
  return PGM_OK;
}

StepArgs make_args(pgm_engine* e) {
  StepArgs a = e->layout;
  const pgm_config& c = e->cfg;
  a.N = c.num_envs;
  a.A = c.num_agents;
  a.PH = e->PH;
  a.PW = e->PW;
  a.WPR = e->WPR;
  a.r = c.obs_radius;
  a.D = e->D;
  a.obst_stride = e->obst_stride;
  a.bits_per_agent = e->bits_per_agent;
  a.stage_bpa = e->stage_bpa;
  a.obs_format = c.obs_format == PGM_OBS_F16 ? 4 : c.obs_format;  // kernel numbering: 3 is the raw stream
  a.max_steps = c.max_episode_steps;
  a.auto_reset = c.auto_reset;
  a.on_target = c.on_target;
  a.batch_agents = e->batch_agents;
  int lg = 0;
  while ((1 << lg) < c.num_agents) lg++;
  a.max_rounds = lg + 2;
  a.obst = e->d_obst;
  a.state = e->d_state;
  a.state0 = e->d_state0;
  a.elapsed = e->d_elapsed;
  a.rng = e->d_rng;
  a.rng0 = e->d_rng0;
  a.comp_start = e->d_cstart;
  a.comp_size = e->d_csize;
  a.cells = e->d_cells;
  a.cells_stride = e->cells_stride;
  a.was_on_goal = e->d_was;
  a.episode_done = e->d_done;
  a.metric_acc = e->d_macc;
  a.metric_last = e->d_mlast;
  a.actions = nullptr;
  a.act_itemsize = 1;
  a.num_steps = 1;
  a.act_step_stride = 0;
  a.out_step_stride = 0;
  a.obs_ring = 1;
  a.obs_slot_stride = 0;
  a.obs = nullptr;
  a.obs_inst_stride = e->obs_inst_stride;
  if (e->ovr_stream) {
    a.obs_format = 3;
    a.obs_inst_stride = e->stream_unit_bytes;
  }
  a.rewards = nullptr;
  a.terminated = nullptr;
  a.truncated = nullptr;
  a.err_flag = e->d_err;
  a.debug = e->d_debug;
  a.regen_flag = e->d_regen_flag;
  a.mask = nullptr;
  return a;
}

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

// Upload generated instances [first, first+count) and make them the current state.
int upload_instances(pgm_engine* e, int first, int count, std::vector<GenInstance>& inst, cudaStream_t s) {
  const int A = e->cfg.num_agents;
  std::vector<uint32_t> obst((size_t)count * e->obst_stride, 0u);
  std::vector<uint2> st((size_t)count * A);
  for (int k = 0; k < count; ++k) {
    memcpy(&obst[(size_t)k * e->obst_stride], inst[k].obst_bits.data(), inst[k].obst_bits.size() * 4);
    for (int a = 0; a < A; ++a) st[(size_t)k * A + a] = make_uint2(inst[k].pos[a] | 0x8000u, inst[k].tgt[a]);
  }
  memcpy(&e->h_obst[(size_t)first * e->obst_stride], obst.data(), obst.size() * 4);
  CUDA_TRY(cudaMemcpyAsync(e->d_obst + (size_t)first * e->obst_stride, obst.data(), obst.size() * 4,
                           cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(e->d_state0 + (size_t)first * A, st.data(), st.size() * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(e->d_state + (size_t)first * A, st.data(), st.size() * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(e->d_was + (size_t)first * A, 0, (size_t)count * A, s));
  CUDA_TRY(cudaMemsetAsync(e->d_done + first, 0, (size_t)count, s));
  CUDA_TRY(cudaMemsetAsync(e->d_elapsed + first, 0, (size_t)count * 4, s));
  CUDA_TRY(cudaMemsetAsync(e->d_macc + (size_t)first * 4, 0, (size_t)count * 16, s));
  CUDA_TRY(cudaMemsetAsync(e->d_mlast + (size_t)first * 4, 0, (size_t)count * 16, s));
  std::vector<Pcg64> rng;
  std::vector<int32_t> cstart, csize;
  std::vector<uint32_t> cells;
  if (e->lifelong) {
    rng.resize((size_t)count * A);
    cstart.resize((size_t)count * A);
    csize.resize((size_t)count * A);
    cells.assign((size_t)count * e->cells_stride, 0u);
    for (int k = 0; k < count; ++k) {
      if ((int64_t)inst[k].cells.size() > e->cells_stride)
        return fail(PGM_ERR_INVALID, "internal: component table larger than the map");
      memcpy(&rng[(size_t)k * A], inst[k].rng.data(), (size_t)A * sizeof(Pcg64));
      memcpy(&cstart[(size_t)k * A], inst[k].comp_start.data(), (size_t)A * 4);
      memcpy(&csize[(size_t)k * A], inst[k].comp_size.data(), (size_t)A * 4);
      memcpy(&cells[(size_t)k * e->cells_stride], inst[k].cells.data(), inst[k].cells.size() * 4);
    }
    CUDA_TRY(cudaMemcpyAsync(e->d_rng + (size_t)first * A, rng.data(), rng.size() * sizeof(Pcg64),
                             cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->d_rng0 + (size_t)first * A, rng.data(), rng.size() * sizeof(Pcg64),
                             cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->d_cstart + (size_t)first * A, cstart.data(), cstart.size() * 4,
                             cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->d_csize + (size_t)first * A, csize.data(), csize.size() * 4,
                             cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->d_cells + (size_t)first * e->cells_stride, cells.data(), cells.size() * 4,
                             cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));  // host vectors go out of scope
  e->tasks_ready = true;
  return PGM_OK;
}

GenParams gen_params(const pgm_engine* e, double density, const uint8_t* map) {
  GenParams p;
  p.H = e->cfg.height;
  p.W = e->cfg.width;
  p.A = e->cfg.num_agents;
  p.r = e->cfg.obs_radius;
  p.density = density;
  p.lifelong = e->lifelong;
  p.map = map;
  return p;
}

template <typename T>
int dev_alloc(T** p, size_t n) {
  CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  CUDA_TRY(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  return PGM_OK;
}

constexpr int64_t kSmallBlock = 256 * 1024;

int ensure_host_scratch(pgm_engine* e, int itemsize) {
  const size_t NA = (size_t)e->cfg.num_envs * e->cfg.num_agents;
  if (!e->d_obs_h) {
    auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
    e->off_rew = up(e->obs_bytes);
    e->off_term = e->off_rew + up((int64_t)NA * 4);
    e->off_trunc = e->off_term + up((int64_t)NA);
    e->out_block_bytes = e->off_trunc + up((int64_t)NA);
    CUDA_TRY(cudaMalloc((void**)&e->d_obs_h, (size_t)e->out_block_bytes));
    e->d_rew_h = (float*)(e->d_obs_h + e->off_rew);
    e->d_term_h = e->d_obs_h + e->off_term;
    e->d_trunc_h = e->d_obs_h + e->off_trunc;
    if (e->out_block_bytes <= kSmallBlock) {
      CUDA_TRY(cudaHostAlloc((void**)&e->h_small, (size_t)e->out_block_bytes + (NA * 9 + 15) / 16 * 16 + NA * 8, cudaHostAllocMapped));
      CUDA_TRY(cudaHostGetDevicePointer((void**)&e->h_small_dev, e->h_small, 0));
      if (e->out_block_bytes > 64 * 1024) e->h_small_dev = nullptr;  // beyond a few instances the copy engine is the better mover
      if (const char* v = getenv("PGM_ZERO_COPY")) {  // tuning knob: 0 = copy engine instead of direct stores
        if (v[0] == '0') e->h_small_dev = nullptr;
      }
    }
  }
  if (e->act_h_itemsize < itemsize) {
    if (e->d_act_h) cudaFree(e->d_act_h);
    e->d_act_h = nullptr;
    CUDA_TRY(cudaMalloc((void**)&e->d_act_h, NA * itemsize));
    e->act_h_itemsize = itemsize;
  }
  return PGM_OK;
}

int ensure_gen_buffers(pgm_engine* e, int slots_wanted) {
  const int N = e->cfg.num_envs, A = e->cfg.num_agents, HW = e->cfg.height * e->cfg.width;
  if (!e->d_gen_seeds) {
    CUDA_TRY(cudaMalloc((void**)&e->d_gen_seeds, (size_t)N * 8));
    CUDA_TRY(cudaMalloc((void**)&e->d_gen_fail, (size_t)N * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_gen_index, (size_t)N * 4));
    CUDA_TRY(cudaMalloc((void**)&e->d_gen_map, (size_t)HW));
  }
  const long long per_inst = devgen_scratch_bytes(HW, A, 1);
  const int slots = (int)std::max<long long>(1, std::min<long long>(slots_wanted, (256LL << 20) / per_inst));
  if (e->gen_scratch_bytes < per_inst * slots) {
    if (e->d_gen_scratch) cudaFree(e->d_gen_scratch);
    e->d_gen_scratch = nullptr;
    CUDA_TRY(cudaMalloc((void**)&e->d_gen_scratch, (size_t)(per_inst * slots)));
    e->gen_scratch_bytes = per_inst * slots;
  }
  return slots;
}

// remember how the tasks were generated (rebuilds with new seeds reuse it) and the current seeds
int remember_generation(pgm_engine* e, int first, int count, const uint64_t* seeds, double density,
                        const uint8_t* map_host, bool explicit_tasks, cudaStream_t s) {
  e->gen_density = density;
  e->gen_has_map = map_host != nullptr;
  e->gen_explicit = explicit_tasks;
  if (seeds) CUDA_TRY(cudaMemcpyAsync(e->d_cur_seeds + first, seeds, (size_t)count * 8, cudaMemcpyHostToDevice, s));
  if (map_host) {
    int rc = ensure_gen_buffers(e, 1);
    if (rc < 0) return rc;
    CUDA_TRY(cudaMemcpyAsync(e->d_gen_map, map_host, (size_t)e->cfg.height * e->cfg.width, cudaMemcpyHostToDevice, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return PGM_OK;
}

DevGenArgs devgen_args(pgm_engine* e, double density, bool has_map) {
  DevGenArgs a{};
  a.H = e->cfg.height;
  a.W = e->cfg.width;
  a.A = e->cfg.num_agents;
  a.r = e->cfg.obs_radius;
  a.lifelong = e->lifelong ? 1 : 0;
  a.map = has_map ? e->d_gen_map : nullptr;
  binomial1_constants(density, &a.binom_zero, &a.binom_flip, &a.binom_qn, &a.binom_px1);
  a.scratch = e->d_gen_scratch;
  a.err_flag = e->d_err;
  a.obst = e->d_obst;
  a.obst_stride = e->obst_stride;
  a.state = e->d_state;
  a.state0 = e->d_state0;
  a.elapsed = e->d_elapsed;
  a.episode_done = e->d_done;
  a.was_on_goal = e->d_was;
  a.metric_acc = e->d_macc;
  a.metric_last = e->d_mlast;
  a.rng = e->d_rng;
  a.rng0 = e->d_rng0;
  a.comp_start = e->d_cstart;
  a.comp_size = e->d_csize;
  a.cells = e->d_cells;
  a.cells_stride = e->cells_stride;
  return a;
}

// ---- packed host transport ------------------------------------------------------------------------
// The step kernel writes each instance's observation bit stream (obs_format 3), the copy engine moves it
// to pinned staging in chunks, and host threads widen chunk c while chunk c+1 is still on the bus.
constexpr int kMaxStreamChunks = 64;
inline int64_t us_since(const pgm_engine* e) {
  return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - e->t_call).count();
}

bool use_packed(const pgm_engine* e) {
  if (e->cfg.obs_format == PGM_OBS_BITS) return false;
  if (e->host_transport >= 0) return e->host_transport == 1;
  if (const char* v = getenv("PGM_HOST_TRANSPORT")) return v[0] == '1' || v[0] == 'p';
  return e->obs_bytes >= (4 << 20);  // auto: below a few MB the DMA of the final tensor is latency-, not PCIe-bound
}

int ensure_stream(pgm_engine* e) {
  if (!e->d_stream) {
    // geometry of obs_format 3 (pgm_kernels.cuh): batch b of an instance starts at the word its first agent would
    // have in the bits format (ceil(bits_per_agent / 32) words per agent); inside a batch the agents are bit-contiguous
    const int64_t A = e->cfg.num_agents, g = e->batch_agents;
    const int64_t wpa = (e->bits_per_agent + 31) / 32;
    e->stream_batch_bytes = g * wpa * 4;
    e->stream_unit_bytes = A * wpa * 4;
    e->stream_bytes = e->stream_unit_bytes * e->cfg.num_envs;
    CUDA_TRY(cudaMalloc((void**)&e->d_stream, (size_t)e->stream_bytes + 64));
    CUDA_TRY(cudaHostAlloc((void**)&e->h_stream, (size_t)e->stream_bytes + 64, cudaHostAllocDefault));
    memset(e->h_stream, 0, (size_t)e->stream_bytes + 64);
    if (const char* v = getenv("PGM_STREAM_CHUNKS")) e->stream_chunks = std::max(1, std::min(kMaxStreamChunks, atoi(v)));
    e->stream_chunks = (int)std::min<int64_t>(e->stream_chunks, e->cfg.num_envs);
    CUDA_TRY(cudaMalloc((void**)&e->d_flags, (size_t)kMaxStreamChunks * 4));
    CUDA_TRY(cudaHostAlloc((void**)&e->h_flags, (size_t)kMaxStreamChunks * 4, cudaHostAllocDefault));
    memset(e->h_flags, 0, (size_t)kMaxStreamChunks * 4);
  }
  if (!e->pool) {
    int t = e->host_threads;
    if (t <= 0) {
      if (const char* v = getenv("PGM_HOST_THREADS")) t = atoi(v);
    }
    if (t <= 0) t = std::min(32, std::max(1, (int)std::thread::hardware_concurrency()));  // the caller is one of them
    e->pool = new pgm::ExpandPool(t);
  }
  return PGM_OK;
}

// Polled by the widening loop when a chunk flag is overdue: has the stream feeding the staging buffer failed?
bool stream_failed(void* ctx) {
  pgm_engine* e = (pgm_engine*)ctx;
  const cudaError_t q = cudaStreamQuery(e->expand_stream);
  return q != cudaSuccess && q != cudaErrorNotReady;
}

void begin_expand(pgm_engine* e, void* obs_host, cudaStream_t s) {
  pgm::ExpandJob j;
  e->expand_stream = s;
  j.producer_failed = stream_failed;
  j.producer_ctx = e;
  const int64_t A = e->cfg.num_agents, g = e->batch_agents;
  j.src = e->h_stream;
  j.dst = (uint8_t*)obs_host;
  j.units = e->cfg.num_envs;
  j.src_unit_stride = e->stream_unit_bytes;
  j.dst_unit_stride = e->obs_inst_stride;
  j.batches = (A + g - 1) / g;
  j.src_batch_stride = e->stream_batch_bytes;
  j.batch_elems = g * e->bits_per_agent;
  j.unit_elems = A * e->bits_per_agent;
  j.elem_size = obs_elem_size(e->cfg.obs_format);
  e->epoch = e->epoch % 255u + 1u;  // 1..255, never the value the flags hold from the previous call
  j.flags = e->h_flags;
  j.flag_value = e->epoch * 0x01010101u;
  j.chunks = e->stream_chunks;
  e->pool->begin(j);
}

void abort_expand(pgm_engine* e) {
  // a failed call: the threads stop waiting for flags and run over whatever the staging buffer holds
  // (the caller ignores the output of a failed call), so that the pool is idle again
  e->pool->abort();
  e->pool->work();
  e->pool->finish();
}

// Once begin_expand() has woken the pool, every exit path must leave it idle again.
struct ExpandGuard {
  pgm_engine* e;
  bool armed;
  ~ExpandGuard() {
    if (armed) abort_expand(e);
  }
};

// Chunked copy of the device stream; the flag copy behind chunk c is stream-ordered after it, so a host
// thread that reads flags[c] == epoch also sees the chunk.
int enqueue_stream_copies(pgm_engine* e, cudaStream_t s) {
  const int64_t N = e->cfg.num_envs;
  const int C = e->stream_chunks;
  cudaError_t err = cudaMemsetAsync(e->d_flags, (int)e->epoch, (size_t)C * 4, s);
  for (int c = 0; c < C && err == cudaSuccess; ++c) {
    const int64_t u0 = N * c / C, u1 = N * (c + 1) / C;
    err = cudaMemcpyAsync(e->h_stream + u0 * e->stream_unit_bytes, e->d_stream + u0 * e->stream_unit_bytes,
                          (size_t)((u1 - u0) * e->stream_unit_bytes), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(e->h_flags + c, e->d_flags + c, 4, cudaMemcpyDeviceToHost, s);
  }
  if (err != cudaSuccess) return fail(PGM_ERR_CUDA, "stream copy failed: %s", cudaGetErrorString(err));
  return PGM_OK;
}

int drain_expand(pgm_engine* e) {
  e->last_us[0] = us_since(e);
  e->pool->work();  // the calling thread widens too
  e->pool->finish();
  e->last_us[1] = e->pool->first_chunk_us();
  e->last_us[2] = e->pool->last_chunk_us();
  e->last_us[3] = us_since(e);
  if (e->pool->aborted()) {
    const cudaError_t q = cudaStreamQuery(e->expand_stream);
    return fail(PGM_ERR_CUDA, "the stream feeding the packed host transport failed: %s", cudaGetErrorString(q));
  }
  return PGM_OK;
}

// auto_reset == 2: after a step, rebuild every instance whose episode ended from its next seed
// (compact the flags -> device generator over the list -> masked observe pass)
int enqueue_rebuilds(pgm_engine* e, void* obs_dev, cudaStream_t s) {
  if (e->gen_density < 0.0 || e->gen_explicit)
    return fail(PGM_ERR_STATE, "auto_reset=2 needs tasks built by pgm_generate / pgm_generate_device");
  if (e->regen_slots == 0) {
    int slots = ensure_gen_buffers(e, e->cfg.num_envs);
    if (slots < 0) return slots;
    e->regen_slots = slots;
  }
  const uint64_t stride = e->cfg.reserved[0] > 0 ? (uint64_t)e->cfg.reserved[0] : (uint64_t)e->cfg.num_envs;
  CUDA_TRY(cudaMemsetAsync(e->d_regen_count, 0, sizeof(int), s));
  int err = launch_regen_compact(e->cfg.num_envs, e->d_regen_flag, e->d_cur_seeds, stride, e->d_gen_index,
                                 e->d_regen_count, s);
  if (err != 0) return fail(PGM_ERR_CUDA, "regen compact launch failed: %s", cudaGetErrorString((cudaError_t)err));
  DevGenArgs a = devgen_args(e, e->gen_density, e->gen_has_map);
  a.first = 0;
  a.count = 0;
  a.index = e->d_gen_index;
  a.count_ptr = e->d_regen_count;
  a.seeds = nullptr;
  a.cur_seeds = e->d_cur_seeds;
  a.fail = nullptr;
  a.slots = e->regen_slots;
  err = launch_devgen(a, s);
  if (err != 0) return fail(PGM_ERR_CUDA, "device generator launch failed: %s", cudaGetErrorString((cudaError_t)err));
  e->launches += 2;
  e->h_obst_valid = false;
  // The step kernel's prologue copies the obstacle bitmap BEFORE griddepcontrol.wait; with programmatic
  // serialization it could read d_obst while the generator grid is still writing it.  The launch that follows a
  // rebuild is therefore an ordinary stream-ordered one.
  e->serialize_next = true;
  if (obs_dev) {
    StepArgs o = make_args(e);
    o.obs = (uint8_t*)obs_dev;
    o.mask = e->d_regen_flag;
    return launch(e, o, OP_OBSERVE, s);
  }
  CUDA_TRY(cudaMemsetAsync(e->d_regen_flag, 0, (size_t)e->cfg.num_envs, s));
  return PGM_OK;
}

}  // namespace

extern "C" {

const char* pgm_last_error(void) { return g_last_error.c_str(); }
int pgm_abi_version(void) { return PGM_ABI_VERSION; }

int pgm_create(const pgm_config* cfg, pgm_engine** out) {
  if (!cfg || !out) return fail(PGM_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != PGM_ABI_VERSION)
    return fail(PGM_ERR_INVALID, "abi_version %d != %d", cfg->abi_version, PGM_ABI_VERSION);
  if (cfg->num_envs < 1) return fail(PGM_ERR_INVALID, "num_envs must be >= 1");
  if (cfg->num_agents < 1 || cfg->num_agents > 65534)
    return fail(PGM_ERR_UNSUPPORTED, "num_agents must be in [1, 65534] per instance, got %d", cfg->num_agents);
  if (cfg->height < 1 || cfg->width < 1 || cfg->height > 1024 || cfg->width > 1024)
    return fail(PGM_ERR_INVALID, "map size must be in [1,1024], got %dx%d", cfg->height, cfg->width);
  if (cfg->obs_radius < 1 || cfg->obs_radius > 128) return fail(PGM_ERR_INVALID, "obs_radius must be in [1,128]");
  if (cfg->max_episode_steps < 1) return fail(PGM_ERR_INVALID, "max_episode_steps must be >= 1");
  if (cfg->collision_system < 0 || cfg->collision_system > 2) return fail(PGM_ERR_INVALID, "bad collision_system");
  if (cfg->on_target < 0 || cfg->on_target > 2) return fail(PGM_ERR_INVALID, "bad on_target");
  if (cfg->auto_reset < 0 || cfg->auto_reset > 2) return fail(PGM_ERR_INVALID, "auto_reset must be 0, 1 or 2");
  if (cfg->obs_format < 0 || cfg->obs_format > 3) return fail(PGM_ERR_INVALID, "bad obs_format");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(PGM_ERR_INVALID, "device %d out of range (%d visible)", cfg->device, ndev);
  DeviceGuard guard(cfg->device);
  pgm_engine* e = new pgm_engine();
  e->cfg = *cfg;
  if (const char* v = getenv("PGM_NO_PDL")) e->use_pdl = !(v[0] == '1');
  {
    int sms = 0;
    const cudaError_t pe = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    if (pe != cudaSuccess) {
      delete e;
      return fail(PGM_ERR_CUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(pe));
    }
    e->sm_count = sms;
  }
  const int r = cfg->obs_radius;
  e->D = 2 * r + 1;
  e->PH = cfg->height + 2 * r;
  e->PW = cfg->width + 2 * r;
  e->WPR = (e->PW + 31) / 32;
  e->obst_stride = round_up(e->PH * e->WPR + 1, 4);
  e->bits_per_agent = 3 * e->D * e->D;
  e->lifelong = cfg->on_target == PGM_ON_TARGET_RESTART;
  const int64_t A = cfg->num_agents, N = cfg->num_envs;
  if (cfg->obs_format == PGM_OBS_BITS) {
    e->stage_bpa = round_up(e->bits_per_agent, 32);
    e->obs_inst_stride = A * (e->stage_bpa / 8);
  } else {
    e->stage_bpa = e->bits_per_agent;
    e->obs_inst_stride = A * e->bits_per_agent * obs_elem_size(cfg->obs_format);
  }
  e->obs_bytes = N * e->obs_inst_stride;
  e->cells_stride = (int64_t)cfg->height * cfg->width;
  int rc = compute_plan(e);
  if (rc != PGM_OK) {
    delete e;
    return rc;
  }
#define TRY_ALLOC(x)   \
  if ((rc = (x)) != PGM_OK) { \
    pgm_destroy(e);    \
    return rc;         \
  }
  TRY_ALLOC(dev_alloc(&e->d_obst, (size_t)N * e->obst_stride));
  TRY_ALLOC(dev_alloc(&e->d_state, (size_t)N * A));
  TRY_ALLOC(dev_alloc(&e->d_state0, (size_t)N * A));
  TRY_ALLOC(dev_alloc(&e->d_was, (size_t)N * A));
  TRY_ALLOC(dev_alloc(&e->d_done, (size_t)N));
  TRY_ALLOC(dev_alloc(&e->d_elapsed, (size_t)N));
  TRY_ALLOC(dev_alloc(&e->d_macc, (size_t)N * 4));
  TRY_ALLOC(dev_alloc(&e->d_mlast, (size_t)N * 4));
  TRY_ALLOC(dev_alloc(&e->d_err, 1));
  TRY_ALLOC(dev_alloc(&e->d_cur_seeds, (size_t)N));
  TRY_ALLOC(dev_alloc(&e->d_regen_flag, (size_t)N));
  TRY_ALLOC(dev_alloc(&e->d_regen_count, 1));
  if (e->lifelong) {
    TRY_ALLOC(dev_alloc(&e->d_rng, (size_t)N * A));
    TRY_ALLOC(dev_alloc(&e->d_rng0, (size_t)N * A));
    TRY_ALLOC(dev_alloc(&e->d_cstart, (size_t)N * A));
    TRY_ALLOC(dev_alloc(&e->d_csize, (size_t)N * A));
    TRY_ALLOC(dev_alloc(&e->d_cells, (size_t)N * e->cells_stride));
  }
#undef TRY_ALLOC
  if (e->fast) {
    // [zeros: one agent bitmap | 0xFF: the uint16 cell grid (priority / soft)], copied by cp.async.bulk
    const int bitmap_bytes = round_up((e->PH * e->WPR + 1) * 4, 16);
    const int grid_bytes = cfg->collision_system == PGM_COLLISION_BLOCK_BOTH ? 0 : round_up(e->PH * e->PW * 2, 16);
    e->fast_fill_bytes = bitmap_bytes + grid_bytes;
    std::vector<uint8_t> tmpl((size_t)e->fast_fill_bytes, 0xFF);
    memset(tmpl.data(), 0, (size_t)bitmap_bytes);
    if (cudaMalloc((void**)&e->d_fast_fill, tmpl.size()) != cudaSuccess ||
        cudaMemcpy(e->d_fast_fill, tmpl.data(), tmpl.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      pgm_destroy(e);
      return fail(PGM_ERR_CUDA, "allocating the fast kernel's fill template failed");
    }
    if (const char* v = getenv("PGM_STAGGER_NS")) e->stagger_ns = std::max(0, atoi(v));
  }
  e->h_obst.assign((size_t)N * e->obst_stride, 0u);
  *out = e;
  return PGM_OK;
}

int pgm_destroy(pgm_engine* e) {
  if (!e) return PGM_OK;
  DeviceGuard guard(e->cfg.device);
  void* ptrs[] = {e->d_obst,  e->d_state, e->d_state0, e->d_was,
                  e->d_done,  e->d_elapsed, e->d_macc, e->d_mlast,  e->d_rng,   e->d_rng0,   e->d_cstart,
                  e->d_csize, e->d_cells, e->d_err,   e->d_act_h,  e->d_obs_h /* base of the result block */,
                  e->d_gen_seeds, e->d_gen_fail, e->d_gen_index, e->d_gen_map, e->d_gen_scratch, e->d_cur_seeds, e->d_regen_flag, e->d_regen_count,
                  e->d_fast_fill};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete e->pool;
  if (e->d_stream) cudaFree(e->d_stream);
  if (e->h_stream) cudaFreeHost(e->h_stream);
  if (e->h_small) cudaFreeHost(e->h_small);
  if (e->d_flags) cudaFree(e->d_flags);
  if (e->h_flags) cudaFreeHost(e->h_flags);
  delete e;
  return PGM_OK;
}

int64_t pgm_obs_bytes(const pgm_engine* e) { return e ? e->obs_bytes : 0; }
int64_t pgm_obs_instance_stride(const pgm_engine* e) { return e ? e->obs_inst_stride : 0; }
int64_t pgm_launch_count(const pgm_engine* e) { return e ? e->launches : 0; }

int pgm_plan(const pgm_engine* e, int32_t* out, int32_t n) {
  if (!e || !out) return fail(PGM_ERR_INVALID, "null argument");
  // [0..6] the generic kernel's plan; [7..12] the fast step kernel's (pgm_fast.cuh): 1 if step launches use it,
  // team threads, agents per thread, teams per CTA, shared memory per CTA, grid
  const int32_t v[13] = {e->team, e->tpc, e->cta_threads, e->smem_cta, e->grid, e->batch_agents, e->occ_mode,
                         e->fast ? 1 : 0, e->f_team, e->f_apt, e->f_tpc, e->f_smem_cta, e->f_grid};
  for (int i = 0; i < n && i < 13; ++i) out[i] = v[i];
  return PGM_OK;
}

int pgm_generate(pgm_engine* e, int32_t first, int32_t count, const uint64_t* seeds, double density,
                 const uint8_t* map_host, int32_t num_threads, int32_t* failed_index, void* stream) {
  if (!e || !seeds) return fail(PGM_ERR_INVALID, "null argument");
  if (first < 0 || count < 0 || first + count > e->cfg.num_envs) return fail(PGM_ERR_INVALID, "bad instance range");
  if (!(density >= 0.0 && density <= 1.0)) return fail(PGM_ERR_INVALID, "density must be in [0,1]");
  DeviceGuard guard(e->cfg.device);
  GenParams gp = gen_params(e, density, map_host);
  std::vector<GenInstance> inst(count);
  std::atomic<int> next(0), bad(-1);
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, count));
  auto work = [&]() {
    for (;;) {
      int k = next.fetch_add(1);
      if (k >= count) break;
      if (generate_instance(gp, seeds[k], inst[k]) != 0) {
        int expect = -1;
        bad.compare_exchange_strong(expect, k);
      }
    }
  };
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  if (bad.load() >= 0) {
    if (failed_index) *failed_index = first + bad.load();
    return fail(PGM_ERR_OVERFLOW,
                "Can't create task. Please check grid grid_config, especially density, num_agent and map. "
                "(instance %d, seed %llu)",
                first + bad.load(), (unsigned long long)seeds[bad.load()]);
  }
  {
    int rc = remember_generation(e, first, count, seeds, density, map_host, false, (cudaStream_t)stream);
    if (rc != PGM_OK) return rc;
  }
  return upload_instances(e, first, count, inst, (cudaStream_t)stream);
}

int pgm_generate_host(int32_t height, int32_t width, int32_t num_agents, int32_t obs_radius, double density,
                      int32_t lifelong, const uint8_t* map_host, uint64_t seed, uint8_t* obstacles_out,
                      int32_t* agents_xy_out, int32_t* targets_xy_out, uint64_t* rng_out,
                      int32_t* comp_size_out) {
  if (height < 1 || width < 1 || num_agents < 1 || obs_radius < 1) return fail(PGM_ERR_INVALID, "bad geometry");
  if (!obstacles_out || !agents_xy_out || !targets_xy_out) return fail(PGM_ERR_INVALID, "null argument");
  GenParams gp;
  gp.H = height;
  gp.W = width;
  gp.A = num_agents;
  gp.r = obs_radius;
  gp.density = density;
  gp.lifelong = lifelong != 0;
  gp.map = map_host;
  GenInstance inst;
  if (generate_instance(gp, seed, inst) != 0)
    return fail(PGM_ERR_OVERFLOW,
                "Can't create task. Please check grid grid_config, especially density, num_agent and map.");
  const int r = obs_radius;
  for (int x = 0; x < height; ++x)
    for (int y = 0; y < width; ++y)
      obstacles_out[(size_t)x * width + y] =
          (inst.obst_bits[(size_t)(x + r) * inst.WPR + ((y + r) >> 5)] >> ((y + r) & 31)) & 1u;
  for (int a = 0; a < num_agents; ++a) {
    agents_xy_out[2 * a] = (int)(inst.pos[a] & 0xFFFF) - r;
    agents_xy_out[2 * a + 1] = (int)(inst.pos[a] >> 16) - r;
    targets_xy_out[2 * a] = (int)(inst.tgt[a] & 0xFFFF) - r;
    targets_xy_out[2 * a + 1] = (int)(inst.tgt[a] >> 16) - r;
    if (gp.lifelong && rng_out) {
      rng_out[4 * a] = inst.rng[a].state_hi;
      rng_out[4 * a + 1] = inst.rng[a].state_lo;
      rng_out[4 * a + 2] = inst.rng[a].inc_hi;
      rng_out[4 * a + 3] = inst.rng[a].inc_lo;
    }
    if (gp.lifelong && comp_size_out) comp_size_out[a] = inst.comp_size[a];
  }
  return PGM_OK;
}

int pgm_generate_device(pgm_engine* e, int32_t first, int32_t count, const uint64_t* seeds, double density,
                        const uint8_t* map_host, int32_t* num_host_fallbacks, void* stream) {
  if (!e || !seeds) return fail(PGM_ERR_INVALID, "null argument");
  if (first < 0 || count < 0 || first + count > e->cfg.num_envs) return fail(PGM_ERR_INVALID, "bad instance range");
  if (!(density >= 0.0 && density <= 1.0)) return fail(PGM_ERR_INVALID, "density must be in [0,1]");
  if (num_host_fallbacks) *num_host_fallbacks = 0;
  if (count == 0) return PGM_OK;
  DeviceGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const int chunk = ensure_gen_buffers(e, count);
  if (chunk < 0) return chunk;
  int rc = remember_generation(e, first, count, seeds, density, map_host, false, s);
  if (rc != PGM_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->d_gen_seeds, seeds, (size_t)count * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(e->d_gen_fail, 0, (size_t)count * 4, s));
  DevGenArgs a = devgen_args(e, density, map_host != nullptr);
  a.index = nullptr;
  a.count_ptr = nullptr;
  a.cur_seeds = nullptr;
  for (int off = 0; off < count; off += chunk) {
    a.first = first + off;
    a.count = std::min(chunk, count - off);
    a.slots = a.count;
    a.seeds = e->d_gen_seeds + off;
    a.fail = e->d_gen_fail + off;
    int err = launch_devgen(a, s);
    if (err != 0) return fail(PGM_ERR_CUDA, "device generator launch failed: %s", cudaGetErrorString((cudaError_t)err));
    e->launches++;
    e->serialize_next = true;
  }
  std::vector<int> failed(count);
  CUDA_TRY(cudaMemcpyAsync(failed.data(), e->d_gen_fail, (size_t)count * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  e->h_obst_valid = false;
  e->tasks_ready = true;
  // the rare instances that need upstream's retry loop (or raise OverflowError) go through the host generator
  GenParams gp = gen_params(e, density, map_host);
  int nfb = 0;
  for (int k = 0; k < count; ++k) {
    if (!failed[k]) continue;
    std::vector<GenInstance> one(1);
    if (generate_instance(gp, seeds[k], one[0]) != 0)
      return fail(PGM_ERR_OVERFLOW,
                  "Can't create task. Please check grid grid_config, especially density, num_agent and map. "
                  "(instance %d, seed %llu)",
                  first + k, (unsigned long long)seeds[k]);
    rc = upload_instances(e, first + k, 1, one, s);
    if (rc != PGM_OK) return rc;
    nfb++;
  }
  if (num_host_fallbacks) *num_host_fallbacks = nfb;
  return PGM_OK;
}

int pgm_set_tasks(pgm_engine* e, int32_t first, int32_t count, const uint8_t* obstacles,
                  const int32_t* agents_xy, const int32_t* targets_xy, const uint64_t* seeds, void* stream) {
  if (!e || !obstacles || !agents_xy || !targets_xy) return fail(PGM_ERR_INVALID, "null argument");
  if (first < 0 || count < 0 || first + count > e->cfg.num_envs) return fail(PGM_ERR_INVALID, "bad instance range");
  DeviceGuard guard(e->cfg.device);
  GenParams gp = gen_params(e, 0.0, nullptr);
  std::vector<GenInstance> inst(count);
  const size_t hw = (size_t)e->cfg.height * e->cfg.width;
  const size_t a2 = (size_t)e->cfg.num_agents * 2;
  for (int k = 0; k < count; ++k) {
    int rc = explicit_instance(gp, seeds ? seeds[k] : 0ull, obstacles + k * hw, agents_xy + k * a2,
                               targets_xy + k * a2, inst[k]);
    if (rc != 0) return fail(PGM_ERR_INVALID, "Position is out of bounds! (instance %d)", first + k);
  }
  e->gen_explicit = true;
  return upload_instances(e, first, count, inst, (cudaStream_t)stream);
}

int pgm_reset(pgm_engine* e, void* obs_dev, void* stream) {
  if (!e) return fail(PGM_ERR_INVALID, "null engine");
  if (!e->tasks_ready) return fail(PGM_ERR_STATE, "pgm_reset before pgm_generate / pgm_set_tasks");
  DeviceGuard guard(e->cfg.device);
  StepArgs a = make_args(e);
  a.obs = (uint8_t*)obs_dev;
  CUDA_TRY(cudaMemsetAsync(e->d_regen_flag, 0, (size_t)e->cfg.num_envs, (cudaStream_t)stream));
  return launch(e, a, OP_RESET, (cudaStream_t)stream);
}

int pgm_observe(pgm_engine* e, void* obs_dev, void* stream) {
  if (!e || !obs_dev) return fail(PGM_ERR_INVALID, "null argument");
  if (!e->tasks_ready) return fail(PGM_ERR_STATE, "pgm_observe before pgm_generate / pgm_set_tasks");
  DeviceGuard guard(e->cfg.device);
  StepArgs a = make_args(e);
  a.obs = (uint8_t*)obs_dev;
  return launch(e, a, OP_OBSERVE, (cudaStream_t)stream);
}

int pgm_step(pgm_engine* e, const void* actions_dev, int32_t action_itemsize, void* obs_dev, float* rewards_dev,
             uint8_t* terminated_dev, uint8_t* truncated_dev, void* stream) {
  if (!e || !actions_dev || !rewards_dev || !terminated_dev || !truncated_dev)
    return fail(PGM_ERR_INVALID, "null argument");
  if (action_itemsize != 1 && action_itemsize != 2 && action_itemsize != 4 && action_itemsize != 8)
    return fail(PGM_ERR_INVALID, "action_itemsize must be 1, 2, 4 or 8");
  if (!e->tasks_ready) return fail(PGM_ERR_STATE, "pgm_step before pgm_generate / pgm_set_tasks");
  DeviceGuard guard(e->cfg.device);
  StepArgs a = make_args(e);
  a.actions = (const uint8_t*)actions_dev;
  a.act_itemsize = action_itemsize;
  a.obs = (uint8_t*)obs_dev;
  a.rewards = rewards_dev;
  a.terminated = terminated_dev;
  a.truncated = truncated_dev;
  if (!e->ovr_stream) a.batch_agents = e->batch_single;  // (the packed stream's geometry follows batch_agents)
  int rc = launch(e, a, OP_STEP, (cudaStream_t)stream);
  if (rc != PGM_OK || e->cfg.auto_reset != 2) return rc;
  return enqueue_rebuilds(e, obs_dev, (cudaStream_t)stream);
}

int pgm_step_many(pgm_engine* e, int32_t num_steps, const void* actions_dev, int32_t action_itemsize, void* obs_dev,
                  int32_t obs_ring, float* rewards_dev, uint8_t* terminated_dev, uint8_t* truncated_dev,
                  void* stream) {
  if (!e || !actions_dev || !rewards_dev || !terminated_dev || !truncated_dev)
    return fail(PGM_ERR_INVALID, "null argument");
  if (num_steps < 1) return fail(PGM_ERR_INVALID, "num_steps must be >= 1");
  if (action_itemsize != 1 && action_itemsize != 2 && action_itemsize != 4 && action_itemsize != 8)
    return fail(PGM_ERR_INVALID, "action_itemsize must be 1, 2, 4 or 8");
  if (obs_dev && obs_ring < 1) return fail(PGM_ERR_INVALID, "obs_ring must be >= 1");
  if (!e->tasks_ready) return fail(PGM_ERR_STATE, "pgm_step_many before pgm_generate / pgm_set_tasks");
  if (e->cfg.auto_reset == 2 && num_steps > 1)
    return fail(PGM_ERR_UNSUPPORTED, "auto_reset=2 rebuilds tasks between launches: use one step per launch");
  DeviceGuard guard(e->cfg.device);
  const long long NA = (long long)e->cfg.num_envs * e->cfg.num_agents;
  StepArgs a = make_args(e);
  a.actions = (const uint8_t*)actions_dev;
  a.act_itemsize = action_itemsize;
  a.num_steps = num_steps;
  a.act_step_stride = NA * action_itemsize;
  a.out_step_stride = NA;
  a.obs = (uint8_t*)obs_dev;
  a.obs_ring = obs_dev ? obs_ring : 1;
  a.obs_slot_stride = e->obs_bytes;
  a.rewards = rewards_dev;
  a.terminated = terminated_dev;
  a.truncated = truncated_dev;
  return launch(e, a, OP_STEP, (cudaStream_t)stream);
}

int pgm_set_host_transport(pgm_engine* e, int32_t mode, int32_t num_threads) {
  if (!e) return fail(PGM_ERR_INVALID, "null engine");
  if (mode < -1 || mode > 1) return fail(PGM_ERR_INVALID, "host transport mode must be -1 (auto), 0 (plain) or 1 (packed)");
  if (num_threads < 0) return fail(PGM_ERR_INVALID, "num_threads must be >= 0");
  if (mode == 1 && e->cfg.obs_format == PGM_OBS_BITS)
    return fail(PGM_ERR_INVALID, "obs_format=bits is already packed: nothing to expand on the host");
  e->host_transport = mode;
  if (num_threads != e->host_threads) {
    delete e->pool;
    e->pool = nullptr;
    e->host_threads = num_threads;
  }
  return PGM_OK;
}

int pgm_host_transport_info(const pgm_engine* e, int64_t* out, int32_t n) {
  if (!e || !out) return fail(PGM_ERR_INVALID, "null argument");
  const int64_t v[10] = {use_packed(e) ? 1 : 0, e->pool ? e->pool->threads() : 0, e->last_h2d_bytes, e->last_d2h_bytes,
                         pgm::expand_isa()[0] == 'a' ? (pgm::expand_isa()[3] == '5' ? 2 : 1) : 0,
                         e->last_us[0], e->last_us[1], e->last_us[2], e->last_us[3], e->last_us[4]};
  for (int i = 0; i < n && i < 10; ++i) out[i] = v[i];
  return PGM_OK;
}

int pgm_expand_bits_host(const uint32_t* src_host, int64_t nbits, void* dst_host, int32_t elem_size) {
  if (!src_host || !dst_host || nbits < 0) return fail(PGM_ERR_INVALID, "bad argument");
  if (elem_size != 1 && elem_size != 2 && elem_size != 4)
    return fail(PGM_ERR_INVALID, "elem_size must be 1 (uint8), 2 (float16) or 4 (float32)");
  // the vector paths may read up to 16 bytes past the last stream word: go through a padded copy
  std::vector<uint32_t> tmp((size_t)((nbits + 31) / 32) + 8, 0u);
  memcpy(tmp.data(), src_host, (size_t)((nbits + 31) / 32) * 4);
  pgm::expand_bits(tmp.data(), (size_t)nbits, dst_host, elem_size);
  return PGM_OK;
}

double pgm_host_fill_gbps(void* dst_host, int64_t bytes, int32_t num_threads, int32_t reps) {
  if (!dst_host || bytes < 4096 || num_threads < 1 || reps < 1) {
    fail(PGM_ERR_INVALID, "pgm_host_fill_gbps: bad argument");
    return -1.0;
  }
  return pgm::host_fill_gbps(dst_host, (size_t)bytes, num_threads, reps);
}

int pgm_step_host(pgm_engine* e, const void* actions_host, int32_t action_itemsize, void* obs_host,
                  float* rewards_host, uint8_t* terminated_host, uint8_t* truncated_host, void* stream) {
  return pgm_step_host_ex(e, actions_host, action_itemsize, obs_host, rewards_host, terminated_host, truncated_host,
                          nullptr, nullptr, stream);
}

int pgm_step_host_ex(pgm_engine* e, const void* actions_host, int32_t action_itemsize, void* obs_host,
                     float* rewards_host, uint8_t* terminated_host, uint8_t* truncated_host, uint8_t* active_host,
                     uint8_t* was_on_goal_host, void* stream) {
  if (!e || !actions_host || !rewards_host || !terminated_host || !truncated_host)
    return fail(PGM_ERR_INVALID, "null argument");
  if (action_itemsize != 1 && action_itemsize != 2 && action_itemsize != 4 && action_itemsize != 8)
    return fail(PGM_ERR_INVALID, "action_itemsize must be 1, 2, 4 or 8");
  DeviceGuard guard(e->cfg.device);
  int rc = ensure_host_scratch(e, action_itemsize);
  if (rc != PGM_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t NA = (size_t)e->cfg.num_envs * e->cfg.num_agents;
  const bool packed = obs_host && use_packed(e);
  const bool small = !packed && e->h_small != nullptr;
  e->t_call = std::chrono::steady_clock::now();
  if (packed && (rc = ensure_stream(e)) != PGM_OK) return rc;
  if (packed) begin_expand(e, obs_host, s);  // wake the host threads under the upload + kernel
  ExpandGuard guard_pool{e, packed};
  const bool zero_copy = small && e->h_small_dev != nullptr;
  if (zero_copy) {
    // a tiny engine (the list API's single instance): the kernel reads the actions from and writes its results to
    // pinned host memory itself - no copy engine round trips, only the launch and one wait
    const size_t act_off = (size_t)e->out_block_bytes + (NA * 9 + 15) / 16 * 16;  // 16-byte aligned: wide actions are read in full
    uint8_t* acts = e->h_small + act_off;
    memcpy(acts, actions_host, NA * action_itemsize);
    uint8_t* dv = e->h_small_dev;
    e->ovr_stream = false;
    rc = pgm_step(e, dv + act_off, action_itemsize, obs_host ? dv : nullptr, (float*)(dv + e->off_rew),
                  dv + e->off_term, dv + e->off_trunc, stream);
  } else {
    CUDA_TRY(cudaMemcpyAsync(e->d_act_h, actions_host, NA * action_itemsize, cudaMemcpyHostToDevice, s));
    e->ovr_stream = packed;
    rc = pgm_step(e, e->d_act_h, action_itemsize, obs_host ? (packed ? e->d_stream : e->d_obs_h) : nullptr, e->d_rew_h,
                  e->d_term_h, e->d_trunc_h, stream);
  }
  e->ovr_stream = false;
  if (rc != PGM_OK) return rc;
  e->last_h2d_bytes = (int64_t)(NA * action_itemsize);
  e->last_d2h_bytes = (int64_t)(NA * 6) + (obs_host ? (packed ? e->stream_bytes : e->obs_bytes) : 0) +
                      (active_host ? (int64_t)NA * 8 : 0) + (was_on_goal_host ? (int64_t)NA : 0);
  const uint2* state_host = nullptr;
  if (small) {
    // a single instance behind the list API: everything in three async copies into pinned staging, one wait
    // (five separate copies into pageable buffers cost ~12 us each, more than the step itself)
    uint8_t* st = e->h_small + e->out_block_bytes;
    if (!zero_copy) CUDA_TRY(cudaMemcpyAsync(e->h_small, e->d_obs_h, (size_t)e->out_block_bytes, cudaMemcpyDeviceToHost, s));
    if (active_host) CUDA_TRY(cudaMemcpyAsync(st, e->d_state, NA * 8, cudaMemcpyDeviceToHost, s));
    if (was_on_goal_host) CUDA_TRY(cudaMemcpyAsync(st + NA * 8, e->d_was, NA, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (obs_host) memcpy(obs_host, e->h_small, (size_t)e->obs_bytes);
    memcpy(rewards_host, e->h_small + e->off_rew, NA * 4);
    memcpy(terminated_host, e->h_small + e->off_term, NA);
    memcpy(truncated_host, e->h_small + e->off_trunc, NA);
    if (was_on_goal_host) memcpy(was_on_goal_host, st + NA * 8, NA);
    state_host = reinterpret_cast<const uint2*>(st);
  } else {
    if (packed) {
      if ((rc = enqueue_stream_copies(e, s)) != PGM_OK) return rc;
    } else if (obs_host) {
      CUDA_TRY(cudaMemcpyAsync(obs_host, e->d_obs_h, (size_t)e->obs_bytes, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaMemcpyAsync(rewards_host, e->d_rew_h, NA * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(terminated_host, e->d_term_h, NA, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(truncated_host, e->d_trunc_h, NA, cudaMemcpyDeviceToHost, s));
    if (active_host) {
      e->h_state_tmp.resize(NA);
      CUDA_TRY(cudaMemcpyAsync(e->h_state_tmp.data(), e->d_state, NA * 8, cudaMemcpyDeviceToHost, s));
      state_host = e->h_state_tmp.data();
    }
    if (was_on_goal_host) CUDA_TRY(cudaMemcpyAsync(was_on_goal_host, e->d_was, NA, cudaMemcpyDeviceToHost, s));
    if (packed) {
      guard_pool.armed = false;  // drain_expand runs the job to its end itself
      if ((rc = drain_expand(e)) != PGM_OK) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  if (active_host)
    for (size_t i = 0; i < NA; ++i) active_host[i] = (uint8_t)((state_host[i].x >> 15) & 1u);
  e->last_us[4] = us_since(e);
  return PGM_OK;
}

int pgm_observe_host(pgm_engine* e, void* obs_host, void* stream) {
  if (!e || !obs_host) return fail(PGM_ERR_INVALID, "null argument");
  DeviceGuard guard(e->cfg.device);
  int rc = ensure_host_scratch(e, 1);
  if (rc != PGM_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool packed = use_packed(e);
  if (packed && (rc = ensure_stream(e)) != PGM_OK) return rc;
  if (packed) begin_expand(e, obs_host, s);
  ExpandGuard guard_pool{e, packed};
  e->ovr_stream = packed;
  rc = pgm_observe(e, packed ? e->d_stream : e->d_obs_h, stream);
  e->ovr_stream = false;
  if (rc != PGM_OK) return rc;
  if (packed) {
    if ((rc = enqueue_stream_copies(e, s)) != PGM_OK) return rc;
    guard_pool.armed = false;
    if ((rc = drain_expand(e)) != PGM_OK) return rc;
  } else {
    CUDA_TRY(cudaMemcpyAsync(obs_host, e->d_obs_h, (size_t)e->obs_bytes, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return PGM_OK;
}

int pgm_get_state(pgm_engine* e, int32_t what, void* dst, int64_t dst_bytes, void* stream) {
  if (!e || !dst) return fail(PGM_ERR_INVALID, "null argument");
  DeviceGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t N = e->cfg.num_envs, A = e->cfg.num_agents, r = e->cfg.obs_radius;
  auto need = [&](int64_t n) { return dst_bytes >= n ? 0 : fail(PGM_ERR_INVALID, "destination too small: %lld < %lld", (long long)dst_bytes, (long long)n); };
  switch (what) {
    case PGM_STATE_POSITIONS:
    case PGM_STATE_TARGETS:
    case PGM_STATE_ACTIVE: {
      if (need(what == PGM_STATE_ACTIVE ? N * A : N * A * 8)) return PGM_ERR_INVALID;
      std::vector<uint2> tmp((size_t)(N * A));
      CUDA_TRY(cudaMemcpyAsync(tmp.data(), e->d_state, tmp.size() * 8, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      if (what == PGM_STATE_ACTIVE) {
        uint8_t* o = (uint8_t*)dst;
        for (size_t i = 0; i < tmp.size(); ++i) o[i] = (tmp[i].x >> 15) & 1u;
      } else {
        int32_t* o = (int32_t*)dst;
        for (size_t i = 0; i < tmp.size(); ++i) {
          const uint32_t w = what == PGM_STATE_POSITIONS ? tmp[i].x : tmp[i].y;
          o[2 * i] = (int32_t)(w & 0x7FFF) - (int32_t)r;
          o[2 * i + 1] = (int32_t)(w >> 16) - (int32_t)r;
        }
      }
      return PGM_OK;
    }
    case PGM_STATE_WAS_ON_GOAL: {
      if (need(N * A)) return PGM_ERR_INVALID;
      CUDA_TRY(cudaMemcpyAsync(dst, e->d_was, (size_t)(N * A), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      return PGM_OK;
    }
    case PGM_STATE_ELAPSED: {
      if (need(N * 4)) return PGM_ERR_INVALID;
      CUDA_TRY(cudaMemcpyAsync(dst, e->d_elapsed, (size_t)N * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      return PGM_OK;
    }
    case PGM_STATE_EPISODE_DONE: {
      if (need(N)) return PGM_ERR_INVALID;
      CUDA_TRY(cudaMemcpyAsync(dst, e->d_done, (size_t)N, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      return PGM_OK;
    }
    case PGM_STATE_METRICS: {
      if (need(N * 16)) return PGM_ERR_INVALID;
      CUDA_TRY(cudaMemcpyAsync(dst, e->d_mlast, (size_t)N * 16, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      return PGM_OK;
    }
    case PGM_STATE_SEEDS: {
      if (need(N * 8)) return PGM_ERR_INVALID;
      CUDA_TRY(cudaMemcpyAsync(dst, e->d_cur_seeds, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      return PGM_OK;
    }
    case PGM_STATE_OBSTACLES: {
      const int64_t H = e->cfg.height, W = e->cfg.width;
      if (need(N * H * W)) return PGM_ERR_INVALID;
      if (!e->h_obst_valid) {
        CUDA_TRY(cudaMemcpyAsync(e->h_obst.data(), e->d_obst, e->h_obst.size() * 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        e->h_obst_valid = true;
      }
      uint8_t* o = (uint8_t*)dst;
      for (int64_t n = 0; n < N; ++n) {
        const uint32_t* bits = &e->h_obst[(size_t)n * e->obst_stride];
        for (int64_t x = 0; x < H; ++x)
          for (int64_t y = 0; y < W; ++y) {
            int64_t px = x + r, py = y + r;
            o[(n * H + x) * W + y] = (bits[px * e->WPR + (py >> 5)] >> (py & 31)) & 1u;
          }
      }
      return PGM_OK;
    }
    default: return fail(PGM_ERR_INVALID, "unknown state selector %d", what);
  }
}

void* pgm_state_ptr(pgm_engine* e, int32_t what) {
  if (!e) return nullptr;
  switch (what) {
    case PGM_STATE_POSITIONS: return e->d_state;  // packed agent state words, see pgm_b200.h
    case PGM_STATE_ELAPSED: return e->d_elapsed;
    case PGM_STATE_WAS_ON_GOAL: return e->d_was;
    case PGM_STATE_EPISODE_DONE: return e->d_done;
    case PGM_STATE_METRICS: return e->d_mlast;
    default: return nullptr;
  }
}

namespace {
// Checkpoint blob: a 64-byte header (magic, ABI, shape and modes - validated on load) followed by the raw arrays.
// With auto_reset == 2 the TASKS change every episode (new seed -> new map, starts, goals, lifelong tables), so
// they are part of the mutable state and are saved too; otherwise the tasks are the ones the engine was built
// with and only the per-step state is saved.
struct CkptHeader {
  uint32_t magic;  // 'PGMC'
  int32_t abi, num_envs, num_agents, height, width, obs_radius, collision_system, on_target, auto_reset, lifelong;
  int32_t reserved[5];
};
static_assert(sizeof(CkptHeader) == 64, "checkpoint header is 64 bytes");
constexpr uint32_t kCkptMagic = 0x434D4750u;

CkptHeader ckpt_header(const pgm_engine* e) {
  CkptHeader h{};
  h.magic = kCkptMagic;
  h.abi = PGM_ABI_VERSION;
  h.num_envs = e->cfg.num_envs;
  h.num_agents = e->cfg.num_agents;
  h.height = e->cfg.height;
  h.width = e->cfg.width;
  h.obs_radius = e->cfg.obs_radius;
  h.collision_system = e->cfg.collision_system;
  h.on_target = e->cfg.on_target;
  h.auto_reset = e->cfg.auto_reset;
  h.lifelong = e->lifelong ? 1 : 0;
  return h;
}

struct CkptPart {
  void* dev;
  size_t bytes;
};
std::vector<CkptPart> ckpt_parts(const pgm_engine* e) {
  const size_t N = e->cfg.num_envs, A = e->cfg.num_agents;
  std::vector<CkptPart> v = {{e->d_state, N * A * 8}, {e->d_was, N * A},   {e->d_elapsed, N * 4},  {e->d_done, N},
                             {e->d_macc, N * 16},     {e->d_mlast, N * 16}, {e->d_cur_seeds, N * 8}};
  if (e->lifelong) v.push_back({e->d_rng, N * A * sizeof(Pcg64)});
  if (e->cfg.auto_reset == 2) {
    v.push_back({e->d_obst, N * (size_t)e->obst_stride * 4});
    v.push_back({e->d_state0, N * A * 8});
    v.push_back({e->d_regen_flag, N});
    if (e->lifelong) {
      v.push_back({e->d_rng0, N * A * sizeof(Pcg64)});
      v.push_back({e->d_cstart, N * A * 4});
      v.push_back({e->d_csize, N * A * 4});
      v.push_back({e->d_cells, N * (size_t)e->cells_stride * 4});
    }
  }
  return v;
}
}  // namespace

int64_t pgm_checkpoint_bytes(const pgm_engine* e) {
  if (!e) return 0;
  int64_t b = (int64_t)sizeof(CkptHeader);
  for (auto& p : ckpt_parts(e)) b += (int64_t)p.bytes;
  return b;
}

int pgm_checkpoint_save(pgm_engine* e, void* dst, int64_t dst_bytes, void* stream) {
  if (!e || !dst) return fail(PGM_ERR_INVALID, "null argument");
  if (dst_bytes < pgm_checkpoint_bytes(e)) return fail(PGM_ERR_INVALID, "checkpoint buffer too small");
  DeviceGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* o = (uint8_t*)dst;
  const CkptHeader h = ckpt_header(e);
  memcpy(o, &h, sizeof(h));
  o += sizeof(h);
  for (auto& p : ckpt_parts(e)) {
    CUDA_TRY(cudaMemcpyAsync(o, p.dev, p.bytes, cudaMemcpyDeviceToHost, s));
    o += p.bytes;
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return PGM_OK;
}

int pgm_checkpoint_load(pgm_engine* e, const void* src, int64_t src_bytes, void* stream) {
  if (!e || !src) return fail(PGM_ERR_INVALID, "null argument");
  if (src_bytes < pgm_checkpoint_bytes(e)) return fail(PGM_ERR_INVALID, "checkpoint buffer too small");
  CkptHeader h;
  memcpy(&h, src, sizeof(h));
  const CkptHeader want = ckpt_header(e);
  if (h.magic != kCkptMagic) return fail(PGM_ERR_INVALID, "not a pgm checkpoint (bad magic)");
  if (memcmp(&h, &want, sizeof(h)) != 0)
    return fail(PGM_ERR_INVALID,
                "checkpoint belongs to another engine: abi %d, %d envs x %d agents, map %dx%d, r=%d, collision %d, "
                "on_target %d, auto_reset %d (this engine: abi %d, %d x %d, %dx%d, r=%d, %d, %d, %d)",
                h.abi, h.num_envs, h.num_agents, h.height, h.width, h.obs_radius, h.collision_system, h.on_target,
                h.auto_reset, want.abi, want.num_envs, want.num_agents, want.height, want.width, want.obs_radius,
                want.collision_system, want.on_target, want.auto_reset);
  DeviceGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const uint8_t* o = (const uint8_t*)src + sizeof(h);
  if (e->cfg.auto_reset != 2) {
    // the tasks are not part of the blob: it must have been taken from an engine built from the same seeds
    std::vector<uint64_t> cur((size_t)e->cfg.num_envs);
    CUDA_TRY(cudaMemcpyAsync(cur.data(), e->d_cur_seeds, cur.size() * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint8_t* q = o;
    for (auto& p : ckpt_parts(e)) {
      if (p.dev == (void*)e->d_cur_seeds) break;
      q += p.bytes;
    }
    if (memcmp(q, cur.data(), cur.size() * 8) != 0)
      return fail(PGM_ERR_INVALID, "checkpoint was taken from an engine with different task seeds");
  }
  for (auto& p : ckpt_parts(e)) {
    CUDA_TRY(cudaMemcpyAsync(p.dev, o, p.bytes, cudaMemcpyHostToDevice, s));
    o += p.bytes;
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  if (e->cfg.auto_reset == 2) e->h_obst_valid = false;  // the maps came with the checkpoint
  return PGM_OK;
}

int pgm_set_debug_buffer(pgm_engine* e, void* dev_ptr) {
  if (!e) return fail(PGM_ERR_INVALID, "null engine");
  e->d_debug = (long long*)dev_ptr;
  return PGM_OK;
}

int pgm_check_errors(pgm_engine* e, void* stream) {
  if (!e) return fail(PGM_ERR_INVALID, "null engine");
  DeviceGuard guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(&flag, e->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  if (flag) CUDA_TRY(cudaMemsetAsync(e->d_err, 0, sizeof(int), s));
  if (flag & 1) return fail(PGM_ERR_ACTION, "an action outside [0,5) was passed to pgm_step (treated as 0 = stay)");
  if (flag & 4)
    return fail(PGM_ERR_OVERFLOW,
                "Can't create task for a new seed during auto_reset=2 (upstream would retry or raise OverflowError); "
                "the instance was reset to its previous task");
  return PGM_OK;
}

}  // extern "C"
