// pgm_hostexpand.h - host half of the packed observation transport of pgm_step_host / pgm_observe_host.
//
// The device writes an instance's observations as a contiguous BIT stream (1 bit per element of upstream
// envs.py :: _get_agents_obs, 8x / 32x smaller than the uint8 / float32 tensor), the copy engine moves that
// stream over PCIe, and a pool of host threads turns bits into the caller's uint8 / float32 buffer with
// non-temporal 64-byte stores while later chunks are still in flight.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace pgm {

// One contiguous run of the stream: bit k of `src` (little endian 32-bit words) -> element k of `dst`.
// elem_size 1: uint8 0/1; 2: float16, 4: float32 0.0/1.0.  `src` may be over-read by up to 16 bytes.
void expand_bits(const uint32_t* src, size_t nbits, void* dst, int elem_size);
const char* expand_isa();  // "avx512bw" | "avx2" | "scalar"

// GB/s of a plain non-temporal fill of `bytes` at dst by `threads` threads, `reps` passes: the DRAM write
// ceiling of this host for the widening loop above (reported by bench.py beside the e2e number).
double host_fill_gbps(void* dst, size_t bytes, int threads, int reps);

// Geometry of the packed stream of a whole observation tensor.
struct ExpandJob {
  const uint8_t* src = nullptr;  // pinned staging copy of the device stream
  uint8_t* dst = nullptr;        // caller's observation buffer
  int64_t units = 0;             // instances
  int64_t src_unit_stride = 0;   // bytes between instances in the stream
  int64_t dst_unit_stride = 0;   // bytes between instances in dst
  int64_t batches = 1;           // observation batches per instance (each starts on a 16-byte boundary)
  int64_t src_batch_stride = 0;  // bytes
  int64_t batch_elems = 0;       // elements (= bits) per full batch
  int64_t unit_elems = 0;        // elements per instance
  int elem_size = 1;
  // Arrival of the stream: chunk c = units [units*c/chunks, units*(c+1)/chunks) is complete once
  // flags[c] == flag_value (the copy engine writes the flag right after the chunk, in stream order).
  // flags == nullptr: everything is already there.
  const volatile uint32_t* flags = nullptr;
  uint32_t flag_value = 0;
  int chunks = 1;
  // Health check of the producer, polled by the CALLING thread (ExpandPool::work) while it waits for a flag:
  // returns true if the stream that feeds the staging buffer has failed (sticky CUDA error), in which case the
  // job is aborted instead of spinning forever on a flag that will never arrive.
  bool (*producer_failed)(void* ctx) = nullptr;
  void* producer_ctx = nullptr;
};

class ExpandPool {
 public:
  explicit ExpandPool(int threads);
  ~ExpandPool();
  int threads() const { return nthreads_; }
  // `threads` counts the calling thread: threads - 1 workers are started, the caller joins in through work().
  void begin(const ExpandJob& job);  // wake the workers; they widen chunks as their flags arrive
  void work();                       // the calling thread widens units too, until none is left to claim
  void abort();                      // stop waiting for flags (a failed launch): units are processed as they are
  void finish();                     // returns when every unit has been widened (sfence'd)
  bool aborted() const;              // the job was cut short (abort() or a failed producer): the output is garbage
  int64_t first_chunk_us() const;    // microseconds after begin() at which chunk 0 / the last chunk was seen complete
  int64_t last_chunk_us() const;
 private:
  struct Impl;
  Impl* impl_;
  int nthreads_;
};

}  // namespace pgm
