// pgm_transport.cu - host-buffer calls of the C-ABI: pgm_step_host(_ex) / pgm_observe_host and their two transports
// (plain DMA of the final tensor; packed = GPU-written bit stream + host threads that widen it, pgm_hostexpand.cpp).
#include "pgm_engine.h"

using namespace pgm;
using namespace pgm_impl;

namespace pgm_impl {

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

void free_transport(pgm_engine* e) {
  delete e->pool;
  e->pool = nullptr;
  if (e->d_stream) cudaFree(e->d_stream);
  if (e->h_stream) cudaFreeHost(e->h_stream);
  if (e->h_small) cudaFreeHost(e->h_small);
  if (e->d_flags) cudaFree(e->d_flags);
  if (e->h_flags) cudaFreeHost(e->h_flags);
}

}  // namespace pgm_impl

namespace {

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
    // the stream of an instance is bit-contiguous and ends before its word-per-agent slot does: the gap is copied
    // to the host with the rest (never read there); zero it once so that it is not uninitialised memory
    CUDA_TRY(cudaMemset(e->d_stream, 0, (size_t)e->stream_bytes + 64));
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

}  // namespace

extern "C" {

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

}  // extern "C"
