// pgm_hostexpand.cpp - bit stream -> uint8 / float32 expansion on the host (see pgm_hostexpand.h).
//
// This is the receiving end of a transport encoding, not a compute path: every bit was produced by the
// step kernel on the GPU (pgm_kernels.cuh, obs_format 3); the host only widens bit k to element k.
#include "pgm_hostexpand.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace pgm {
namespace {

inline uint64_t get64(const uint8_t* src, size_t bitpos) {
  unsigned __int128 v;
  memcpy(&v, src + (bitpos >> 3), 16);
  return (uint64_t)(v >> (bitpos & 7));
}
inline uint64_t lowmask(size_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

// ----------------------------------------------------------------------------- scalar
struct Lut {
  uint64_t t[256];
  Lut() {
    for (int b = 0; b < 256; ++b) {
      uint64_t v = 0;
      for (int i = 0; i < 8; ++i) v |= (uint64_t)((b >> i) & 1) << (8 * i);
      t[b] = v;
    }
  }
};
const Lut g_lut;

void expand_u8_scalar(const uint8_t* src, size_t nbits, uint8_t* dst) {
  size_t pos = 0;
  for (; pos + 8 <= nbits; pos += 8) memcpy(dst + pos, &g_lut.t[src[pos >> 3]], 8);
  for (; pos < nbits; ++pos) dst[pos] = (src[pos >> 3] >> (pos & 7)) & 1;
}
void expand_f32_scalar(const uint8_t* src, size_t nbits, float* dst) {
  for (size_t pos = 0; pos < nbits; ++pos) dst[pos] = ((src[pos >> 3] >> (pos & 7)) & 1) ? 1.0f : 0.0f;
}

void expand_f16_scalar(const uint8_t* src, size_t nbits, uint16_t* dst) {
  for (size_t pos = 0; pos < nbits; ++pos) dst[pos] = ((src[pos >> 3] >> (pos & 7)) & 1) ? 0x3C00 : 0;
}

// ----------------------------------------------------------------------------- AVX-512
__attribute__((target("avx512f,avx512bw"))) void expand_u8_avx512(const uint8_t* src, size_t nbits, uint8_t* dst) {
  const __m512i one = _mm512_set1_epi8(1);
  size_t pos = std::min(nbits, (size_t)((64 - ((uintptr_t)dst & 63)) & 63));
  if (pos) {
    const __mmask64 k = lowmask(pos);
    _mm512_mask_storeu_epi8(dst, k, _mm512_maskz_mov_epi8(get64(src, 0) & k, one));
  }
  for (; pos + 64 <= nbits; pos += 64)
    _mm512_stream_si512((__m512i*)(dst + pos), _mm512_maskz_mov_epi8(get64(src, pos), one));
  if (pos < nbits) {
    const __mmask64 k = lowmask(nbits - pos);
    _mm512_mask_storeu_epi8(dst + pos, k, _mm512_maskz_mov_epi8(get64(src, pos) & k, one));
  }
}
__attribute__((target("avx512f,avx512bw"))) void expand_f32_avx512(const uint8_t* src, size_t nbits, float* dst) {
  const __m512 one = _mm512_set1_ps(1.0f);
  size_t pos = std::min(nbits, (size_t)(((64 - ((uintptr_t)dst & 63)) & 63) >> 2));
  if (pos) {
    const __mmask16 k = (__mmask16)lowmask(pos);
    _mm512_mask_storeu_ps(dst, k, _mm512_maskz_mov_ps((__mmask16)get64(src, 0) & k, one));
  }
  for (; pos + 64 <= nbits; pos += 64) {
    const uint64_t m = get64(src, pos);
    _mm512_stream_ps(dst + pos, _mm512_maskz_mov_ps((__mmask16)m, one));
    _mm512_stream_ps(dst + pos + 16, _mm512_maskz_mov_ps((__mmask16)(m >> 16), one));
    _mm512_stream_ps(dst + pos + 32, _mm512_maskz_mov_ps((__mmask16)(m >> 32), one));
    _mm512_stream_ps(dst + pos + 48, _mm512_maskz_mov_ps((__mmask16)(m >> 48), one));
  }
  for (; pos < nbits; pos += 16) {
    const __mmask16 k = (__mmask16)lowmask(std::min<size_t>(16, nbits - pos));
    _mm512_mask_storeu_ps(dst + pos, k, _mm512_maskz_mov_ps((__mmask16)get64(src, pos) & k, one));
  }
}

__attribute__((target("avx512f,avx512bw"))) void expand_f16_avx512(const uint8_t* src, size_t nbits, uint16_t* dst) {
  const __m512i one = _mm512_set1_epi16(0x3C00);  // 1.0 in binary16
  size_t pos = std::min(nbits, (size_t)(((64 - ((uintptr_t)dst & 63)) & 63) >> 1));
  if (pos) {
    const __mmask32 k = (__mmask32)lowmask(pos);
    _mm512_mask_storeu_epi16(dst, k, _mm512_maskz_mov_epi16((__mmask32)get64(src, 0) & k, one));
  }
  for (; pos + 64 <= nbits; pos += 64) {
    const uint64_t m = get64(src, pos);
    _mm512_stream_si512((__m512i*)(dst + pos), _mm512_maskz_mov_epi16((__mmask32)m, one));
    _mm512_stream_si512((__m512i*)(dst + pos + 32), _mm512_maskz_mov_epi16((__mmask32)(m >> 32), one));
  }
  for (; pos < nbits; pos += 32) {
    const __mmask32 k = (__mmask32)lowmask(std::min<size_t>(32, nbits - pos));
    _mm512_mask_storeu_epi16(dst + pos, k, _mm512_maskz_mov_epi16((__mmask32)get64(src, pos) & k, one));
  }
}

// ----------------------------------------------------------------------------- AVX2
__attribute__((target("avx2"))) void expand_f16_avx2(const uint8_t* src, size_t nbits, uint16_t* dst) {
  size_t pos = std::min(nbits, (size_t)(((32 - ((uintptr_t)dst & 31)) & 31) >> 1));
  for (size_t i = 0; i < pos; ++i) dst[i] = ((src[i >> 3] >> (i & 7)) & 1) ? 0x3C00 : 0;
  const __m256i bit = _mm256_setr_epi16(1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, (short)0x8000);
  const __m256i one = _mm256_set1_epi16(0x3C00);
  for (; pos + 16 <= nbits; pos += 16) {
    const __m256i v = _mm256_set1_epi16((short)(get64(src, pos) & 0xFFFF));
    const __m256i hit = _mm256_cmpeq_epi16(_mm256_and_si256(v, bit), bit);
    _mm256_stream_si256((__m256i*)(dst + pos), _mm256_and_si256(hit, one));
  }
  for (; pos < nbits; ++pos) dst[pos] = ((src[pos >> 3] >> (pos & 7)) & 1) ? 0x3C00 : 0;
}
__attribute__((target("avx2"))) void expand_u8_avx2(const uint8_t* src, size_t nbits, uint8_t* dst) {
  size_t pos = std::min(nbits, (size_t)((32 - ((uintptr_t)dst & 31)) & 31));
  for (size_t i = 0; i < pos; ++i) dst[i] = (src[i >> 3] >> (i & 7)) & 1;
  const __m256i shuf = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3,
                                        3, 3, 3, 3, 3);
  const __m256i bit = _mm256_set1_epi64x((long long)0x8040201008040201ull);
  const __m256i one = _mm256_set1_epi8(1);
  for (; pos + 32 <= nbits; pos += 32) {
    __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)(uint32_t)get64(src, pos)), shuf);
    v = _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_and_si256(v, bit), bit), one);
    _mm256_stream_si256((__m256i*)(dst + pos), v);
  }
  for (; pos < nbits; ++pos) dst[pos] = (src[pos >> 3] >> (pos & 7)) & 1;
}
__attribute__((target("avx2"))) void expand_f32_avx2(const uint8_t* src, size_t nbits, float* dst) {
  size_t pos = std::min(nbits, (size_t)(((32 - ((uintptr_t)dst & 31)) & 31) >> 2));
  for (size_t i = 0; i < pos; ++i) dst[i] = ((src[i >> 3] >> (i & 7)) & 1) ? 1.0f : 0.0f;
  const __m256i bit = _mm256_setr_epi32(1, 2, 4, 8, 16, 32, 64, 128);
  const __m256 one = _mm256_set1_ps(1.0f);
  for (; pos + 8 <= nbits; pos += 8) {
    const __m256i v = _mm256_set1_epi32((int)(get64(src, pos) & 0xFF));
    const __m256i hit = _mm256_cmpeq_epi32(_mm256_and_si256(v, bit), bit);
    _mm256_stream_ps(dst + pos, _mm256_and_ps(_mm256_castsi256_ps(hit), one));
  }
  for (; pos < nbits; ++pos) dst[pos] = ((src[pos >> 3] >> (pos & 7)) & 1) ? 1.0f : 0.0f;
}

int detect_isa() {
  if (const char* v = getenv("PGM_HOST_ISA")) {  // testing knob: scalar | avx2 | avx512bw
    if (!strcmp(v, "scalar")) return 0;
    if (!strcmp(v, "avx2") && __builtin_cpu_supports("avx2")) return 1;
  }
  if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) return 2;
  if (__builtin_cpu_supports("avx2")) return 1;
  return 0;
}
int isa() {
  static const int v = detect_isa();
  return v;
}

}  // namespace

const char* expand_isa() { return isa() == 2 ? "avx512bw" : (isa() == 1 ? "avx2" : "scalar"); }

void expand_bits(const uint32_t* src32, size_t nbits, void* dst, int elem_size) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(src32);
  const int k = isa();
  if (elem_size == 1) {
    if (k == 2) expand_u8_avx512(src, nbits, (uint8_t*)dst);
    else if (k == 1) expand_u8_avx2(src, nbits, (uint8_t*)dst);
    else expand_u8_scalar(src, nbits, (uint8_t*)dst);
  } else if (elem_size == 2) {
    if (((uintptr_t)dst & 1) != 0) expand_f16_scalar(src, nbits, (uint16_t*)dst);
    else if (k == 2) expand_f16_avx512(src, nbits, (uint16_t*)dst);
    else if (k == 1) expand_f16_avx2(src, nbits, (uint16_t*)dst);
    else expand_f16_scalar(src, nbits, (uint16_t*)dst);
  } else {
    if (((uintptr_t)dst & 3) != 0) expand_f32_scalar(src, nbits, (float*)dst);
    else if (k == 2) expand_f32_avx512(src, nbits, (float*)dst);
    else if (k == 1) expand_f32_avx2(src, nbits, (float*)dst);
    else expand_f32_scalar(src, nbits, (float*)dst);
  }
}

// ----------------------------------------------------------------------------- thread pool
struct ExpandPool::Impl {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  std::atomic<uint64_t> generation{0};
  std::atomic<bool> stop{false};
  std::atomic<int> running{0};
  std::atomic<bool> aborted{false};
  std::chrono::steady_clock::time_point t_begin;
  std::atomic<int64_t> first_flag_us{-1}, last_flag_us{-1};  // when a thread first saw chunk 0 / the last chunk complete
  int spin_us = 0;  // a worker polls this long for the next job before it sleeps (the wake-up hides under the
                    // upload + kernel + first chunk anyway; polling only pays when cores are to spare)
  ExpandJob job;
  int64_t grain = 1;
  alignas(64) std::atomic<int64_t> next{0};

  void run_job(bool caller = false) {
    const ExpandJob& j = job;
    int c = 0;  // chunk of the last unit claimed (claims only move forward)
    for (;;) {
      const int64_t u0 = next.fetch_add(grain, std::memory_order_relaxed);
      if (u0 >= j.units) break;
      const int64_t u1 = std::min(j.units, u0 + grain);
      if (j.flags != nullptr) {
        while (j.units * (c + 1) / j.chunks < u1) ++c;
        // chunks land in order; x86 does not reorder the data loads before this flag load
        for (int spins = 0; j.flags[c] != j.flag_value && !aborted.load(std::memory_order_relaxed); ++spins) {
          if (spins < 4000) {
            _mm_pause();
          } else {
            // a chunk that takes this long is unusual: the calling thread asks the producer whether it is still alive
            if (caller && j.producer_failed != nullptr && (spins & 255) == 0 && j.producer_failed(j.producer_ctx))
              aborted.store(true, std::memory_order_relaxed);
            std::this_thread::yield();
          }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        if ((c == 0 && first_flag_us.load(std::memory_order_relaxed) < 0) ||
            (c == j.chunks - 1 && last_flag_us.load(std::memory_order_relaxed) < 0)) {
          const int64_t us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_begin).count();
          int64_t unset = -1;
          if (c == 0) first_flag_us.compare_exchange_strong(unset, us, std::memory_order_relaxed);
          unset = -1;
          if (c == j.chunks - 1) last_flag_us.compare_exchange_strong(unset, us, std::memory_order_relaxed);
        }
      }
      for (int64_t u = u0; u < u1; ++u) {
        const uint8_t* s = j.src + u * j.src_unit_stride;
        uint8_t* d = j.dst + u * j.dst_unit_stride;
        int64_t left = j.unit_elems;
        for (int64_t b = 0; b < j.batches && left > 0; ++b) {
          const int64_t n = std::min(left, j.batch_elems);
          expand_bits(reinterpret_cast<const uint32_t*>(s + b * j.src_batch_stride), (size_t)n,
                      d + b * j.batch_elems * j.elem_size, j.elem_size);
          left -= n;
        }
      }
    }
    _mm_sfence();  // non-temporal stores are globally visible before the job is reported done
  }

  void worker() {
    uint64_t seen = 0;
    for (;;) {
      const auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; generation.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed); ++i) {
        _mm_pause();
        if ((i & 63) == 63 &&
            std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us)
          break;
      }
      if (generation.load(std::memory_order_acquire) == seen) {
        std::unique_lock<std::mutex> lk(mu);
        cv_start.wait(lk, [&] { return stop.load() || generation.load() != seen; });
      }
      if (stop.load()) return;
      seen = generation.load(std::memory_order_acquire);
      run_job();
      if (running.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(mu);
        cv_done.notify_all();
      }
    }
  }
};

ExpandPool::ExpandPool(int threads) : impl_(new Impl()), nthreads_(std::max(1, threads)) {
  if (const char* v = getenv("PGM_HOST_SPIN_US")) impl_->spin_us = std::max(0, atoi(v));
  for (int i = 0; i + 1 < nthreads_; ++i) impl_->workers.emplace_back([this] { impl_->worker(); });
}

ExpandPool::~ExpandPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv_start.notify_all();
  for (auto& t : impl_->workers) t.join();
  delete impl_;
}

void ExpandPool::begin(const ExpandJob& job) {
  std::lock_guard<std::mutex> lk(impl_->mu);
  impl_->job = job;
  // ~64 KB of output per claim: coarse enough to amortise the atomic, fine enough to balance the tail
  const int64_t unit_bytes = std::max<int64_t>(1, job.unit_elems * job.elem_size);
  impl_->grain = std::max<int64_t>(1, std::min<int64_t>((64 * 1024) / unit_bytes,
                                                          (job.units + 4 * nthreads_ - 1) / (4 * nthreads_)));
  impl_->next.store(0, std::memory_order_relaxed);
  impl_->aborted.store(false, std::memory_order_relaxed);
  impl_->t_begin = std::chrono::steady_clock::now();
  impl_->first_flag_us.store(-1, std::memory_order_relaxed);
  impl_->last_flag_us.store(-1, std::memory_order_relaxed);
  impl_->running.store(nthreads_ - 1, std::memory_order_relaxed);
  impl_->generation.fetch_add(1, std::memory_order_release);
  impl_->cv_start.notify_all();
}

void ExpandPool::work() { impl_->run_job(true); }
bool ExpandPool::aborted() const { return impl_->aborted.load(std::memory_order_relaxed); }
int64_t ExpandPool::first_chunk_us() const { return impl_->first_flag_us.load(); }
int64_t ExpandPool::last_chunk_us() const { return impl_->last_flag_us.load(); }
void ExpandPool::abort() { impl_->aborted.store(true, std::memory_order_relaxed); }

// Plain multi-threaded non-temporal fill: the host's DRAM write ceiling for the widening loop (bench.py).
namespace {
__attribute__((target("avx512f"))) void fill_avx512(uint8_t* d, size_t n) {
  const __m512i v = _mm512_set1_epi8(1);
  for (size_t i = 0; i + 64 <= n; i += 64) _mm512_stream_si512((__m512i*)(d + i), v);
}
__attribute__((target("avx2"))) void fill_avx2(uint8_t* d, size_t n) {
  const __m256i v = _mm256_set1_epi8(1);
  for (size_t i = 0; i + 32 <= n; i += 32) _mm256_stream_si256((__m256i*)(d + i), v);
}
}  // namespace

double host_fill_gbps(void* dst, size_t bytes, int threads, int reps) {
  threads = std::max(1, threads);
  reps = std::max(1, reps);
  uint8_t* base = (uint8_t*)(((uintptr_t)dst + 63) & ~(uintptr_t)63);
  const size_t usable = (bytes - (size_t)(base - (uint8_t*)dst)) & ~(size_t)63;
  const size_t grain = 64 * 1024;  // claimed dynamically, like the widening loop: a descheduled thread does not hold the others up
  const size_t ngrains = usable / grain;
  if (ngrains == 0) return 0.0;
  const int k = isa();
  double best = 0.0;
  for (int attempt = 0; attempt < 3; ++attempt) {
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::atomic<size_t> next{0};
    auto body = [&]() {
      ready.fetch_add(1);
      while (!go.load(std::memory_order_acquire)) _mm_pause();
      for (;;) {
        const size_t g = next.fetch_add(1, std::memory_order_relaxed);
        if (g >= ngrains * (size_t)reps) break;
        uint8_t* d = base + (g % ngrains) * grain;
        if (k == 2) fill_avx512(d, grain);
        else if (k == 1) fill_avx2(d, grain);
        else memset(d, 1, grain);
      }
      _mm_sfence();
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(body);
    while (ready.load() < threads - 1) _mm_pause();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    body();
    for (auto& t : th) t.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::max(best, (double)grain * ngrains * reps / s / 1e9);
  }
  return best;
}

void ExpandPool::finish() {
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; impl_->running.load(std::memory_order_acquire) != 0; ++i) {
    _mm_pause();
    if ((i & 63) == 63 &&
        std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= 2000)
      break;
  }
  if (impl_->running.load(std::memory_order_acquire) == 0) return;
  std::unique_lock<std::mutex> lk(impl_->mu);
  impl_->cv_done.wait(lk, [&] { return impl_->running.load() == 0; });
}

}  // namespace pgm
