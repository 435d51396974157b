// pgm_plan.cu - planner and launch of the step kernels: shared-memory layouts, team / CTA geometry, eligibility
// of the register-resident kernel (pgm_fast.cuh), kernel arguments, dispatch to the instantiated variants.
#include "pgm_engine.h"

using namespace pgm;

namespace pgm_impl {
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
  // One agent per thread when a team of up to 256 threads can hold the instance: measured on a B200 with both launch
  // forms, 64 agents run 8-9 % faster on 64 threads than on 32 (16 steps per launch 17.0 -> 15.6 us, one launch per
  // step 19.4 -> 18.8 us), 256 agents 4-9 % faster on 256 threads than on 128 (16.8 -> 16.1 / 21.5 -> 19.6 us): twice
  // the warps keep twice the stores in flight, and a step has one observation batch instead of two.
  // Jobs that put more than 32 instances on an SM keep two agents per thread: more instances stay resident
  // (16384 x 64 agents, r=3: 28.0 us per step on 32 threads, 29.7 on 64; one launch per step 33.7 / 40.3).
  const int per_sm = (c.num_envs + e->sm_count - 1) / e->sm_count;
  if (c.team_threads != 0) team = std::max(32, std::min(team, 256));  // the caller's choice
  else team = std::max(32, std::min(per_sm > 32 ? pow2_ceil((A + 1) / 2) : pow2_ceil(A), 256));
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
  auto build = [&](StepArgs* L) {
    int off = 0;
    L->off_obst = off;
    off += e->obst_stride * 4;
    L->off_abits = off;
    off += bitmap_bytes;
    L->off_pbits = off;  // block_both: the second agent bitmap
    if (bb) off += bitmap_bytes;
    L->off_occ = off;
    L->off_stage = off;  // block_both: the observation stream lies over the claim planes (zeroed at the start of a step)
    if (bb) {
      off += std::max(2 * bitmap_bytes, stage_one);
    } else {
      off += round_up(e->PH * e->PW * 2, 16);
      L->off_stage = off;
      off += stage_one;
    }
    L->off_link = off;
    if (!bb) off += round_up(apt * team * 4, 16);
    L->off_npos = off;
    if (!bb) off += round_up(apt * team * 4, 16);
    L->off_misc = off;
    off += 16;
    L->team_smem = round_up(off, 16);
    L->plane_words = bitmap_bytes / 4;
    L->narrow = (e->WPR == 2 && c.width <= 32) ? 1 : 0;
    return L->team_smem;
  };
  auto fit = [](int team_smem) { return (228 * 1024) / (team_smem + 1024); };
  StepArgs L{};
  const int need = std::min(want, std::max(1, 1024 / team));
  const int sm = build(&L);
  if (sm > smem_max) return false;
  if (fit(sm) < std::min(need, 2) && want > 1) return false;  // the generic kernel's leaner layouts keep more instances resident
  e->f_layout = L;
  e->f_team = team;
  e->f_apt = apt;
  return true;
}

}  // namespace

int compute_plan(pgm_engine* e) {
  const pgm_config& c = e->cfg;
  const int A = c.num_agents;
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
    f.plane_words = L.plane_words;
    f.narrow = L.narrow;
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
  a.solve = e->d_solve;
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

}  // namespace pgm_impl
