// Host-side throughput of the packed transport's widening loop (no GPU): config-2 geometry.
// g++ -O3 -std=c++17 -Ipogema_b200/csrc tools/prototypes/hostexpand_bench.cpp build/pgm_hostexpand.o -lpthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "pgm_hostexpand.h"
int main(int argc, char** argv) {
  int threads = argc > 1 ? atoi(argv[1]) : 8;
  int es = argc > 2 ? atoi(argv[2]) : 1;
  const int64_t N = 4096, bits = 64 * 363, sstride = ((bits + 31) / 32 + 3) / 4 * 16;
  std::vector<uint8_t> src(N * sstride + 64);
  for (auto& b : src) b = rand();
  uint8_t* dst = (uint8_t*)aligned_alloc(64, N * bits * es);
  memset(dst, 0, N * bits * es);
  pgm::ExpandPool pool(threads);
  pgm::ExpandJob j;
  j.src = src.data(); j.dst = dst; j.units = N; j.src_unit_stride = sstride; j.dst_unit_stride = bits * es;
  j.batches = 1; j.src_batch_stride = sstride; j.batch_elems = bits; j.unit_elems = bits; j.elem_size = es;
  for (int it = 0; it < 3; ++it) { pool.begin(j); pool.work(); pool.finish(); }
  auto t0 = std::chrono::steady_clock::now();
  const int iters = 20;
  for (int it = 0; it < iters; ++it) { pool.begin(j); pool.work(); pool.finish(); }
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
  printf("isa %s threads %d elem %d: %.3f ms per tensor (%.1f MB) = %.1f GB/s written, %.1f M agent-steps/s\n",
         pgm::expand_isa(), threads, es, dt * 1e3, N * bits * es / 1e6, N * bits * es / dt / 1e9, N * 64 / dt / 1e6);
  return 0;
}
