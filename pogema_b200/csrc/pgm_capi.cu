// pgm_capi.cu - the C-ABI declared in include/pgm_b200.h: engine lifetime, task generation, device-pointer
// step / reset / observe, state access, checkpoints, errors.  (Planner + launch: pgm_plan.cu; host-buffer calls
// and the packed transport: pgm_transport.cu; the engine object: pgm_engine.h.)
#include "pgm_engine.h"

using namespace pgm;
using namespace pgm_impl;

namespace {
thread_local std::string g_last_error;
}  // namespace

namespace pgm_impl {
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace pgm_impl

namespace {

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
  if (e->d_solve) CUDA_TRY(cudaMemsetAsync(e->d_solve + (size_t)first * A * 2, 0, (size_t)count * A * 8, s));
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
  a.solve = e->d_solve;
  a.rng = e->d_rng;
  a.rng0 = e->d_rng0;
  a.comp_start = e->d_cstart;
  a.comp_size = e->d_csize;
  a.cells = e->d_cells;
  a.cells_stride = e->cells_stride;
  return a;
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
  if (e->cfg.on_target == 1) TRY_ALLOC(dev_alloc(&e->d_solve, (size_t)N * A * 2));
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
                  e->d_gen_seeds, e->d_gen_fail, e->d_gen_index, e->d_gen_map, e->d_gen_scratch, e->d_cur_seeds, e->d_regen_flag, e->d_regen_count, e->d_solve};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  free_transport(e);
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
    case PGM_STATE_SOLVE_COSTS: {
      if (!e->d_solve) return fail(PGM_ERR_INVALID, "solve costs exist only with on_target = nothing");
      if (need(N * A * 4)) return PGM_ERR_INVALID;
      std::vector<int32_t> tmp((size_t)(N * A * 2));
      CUDA_TRY(cudaMemcpyAsync(tmp.data(), e->d_solve, tmp.size() * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      int32_t* o = (int32_t*)dst;
      for (int64_t i = 0; i < N * A; ++i) o[i] = tmp[2 * i + 1];
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
    case PGM_STATE_SOLVE_COSTS: return e->d_solve;
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
  if (e->d_solve) v.push_back({e->d_solve, N * A * 8});
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
