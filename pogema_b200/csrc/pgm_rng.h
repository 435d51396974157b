// pgm_rng.h - bit-exact restatement of the numpy Generator pieces the POGEMA
// path consumes (numpy is a third-party dependency of upstream pogema; pinned
// here against numpy 2.3.5, see tests/test_native_rng.py):
//   SeedSequence(entropy).generate_state(4, uint64)      numpy/random/bit_generator.pyx
//   PCG64 (setseq 128, XSL-RR 64) seeding / next64 / buffered next32 / next_double
//                                                         numpy/random/src/pcg64/pcg64.h
//   bounded Lemire uint32 (Generator.integers / Generator.choice with replace)
//                                                         src/distributions/distributions.c
//   random_interval (masked rejection, Generator.shuffle) src/distributions/distributions.c
// Usable from host and device code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PGM_HD __host__ __device__ __forceinline__
#else
#define PGM_HD inline
#endif

namespace pgm {

struct Pcg64 {
  uint64_t state_lo, state_hi;
  uint64_t inc_lo, inc_hi;
  uint32_t has_uint32;
  uint32_t uinteger;
};

PGM_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// PCG_DEFAULT_MULTIPLIER_128 = 0x2360ed051fc65da4_4385df649fccf645
PGM_HD void pcg64_step(Pcg64& g) {
  const uint64_t m_hi = 0x2360ed051fc65da4ULL, m_lo = 0x4385df649fccf645ULL;
  uint64_t lo = g.state_lo * m_lo;
  uint64_t hi = mulhi64(g.state_lo, m_lo) + g.state_lo * m_hi + g.state_hi * m_lo;
  uint64_t nlo = lo + g.inc_lo;
  uint64_t carry = nlo < lo ? 1u : 0u;
  g.state_lo = nlo;
  g.state_hi = hi + g.inc_hi + carry;
}

PGM_HD uint64_t pcg64_next64(Pcg64& g) {
  pcg64_step(g);
  uint64_t x = g.state_hi ^ g.state_lo;
  unsigned rot = (unsigned)(g.state_hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

PGM_HD uint32_t pcg64_next32(Pcg64& g) {
  if (g.has_uint32) {
    g.has_uint32 = 0;
    return g.uinteger;
  }
  uint64_t n = pcg64_next64(g);
  g.has_uint32 = 1;
  g.uinteger = (uint32_t)(n >> 32);
  return (uint32_t)n;
}

PGM_HD double pcg64_next_double(Pcg64& g) {
  return (double)(pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0);
}

// buffered_bounded_lemire_uint32: uniform integer in [0, rng] (rng inclusive).
PGM_HD uint32_t pcg64_bounded32(Pcg64& g, uint32_t rng) {
  if (rng == 0) return 0;
  if (rng == 0xFFFFFFFFu) return pcg64_next32(g);
  const uint32_t rng_excl = rng + 1u;
  uint64_t m = (uint64_t)pcg64_next32(g) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)pcg64_next32(g) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return (uint32_t)(m >> 32);
}

// random_interval: uniform integer in [0, max] by masked rejection.
PGM_HD uint64_t pcg64_interval(Pcg64& g, uint64_t max) {
  if (max == 0) return 0;
  uint64_t mask = max, value;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  mask |= mask >> 32;
  if (max <= 0xffffffffULL) {
    while ((value = ((uint64_t)pcg64_next32(g) & mask)) > max) {
    }
  } else {
    while ((value = (pcg64_next64(g) & mask)) > max) {
    }
  }
  return value;
}

// SeedSequence(entropy = one non-negative integer < 2^64).generate_state(4, uint64)
// followed by pcg64_set_seed (pcg_setseq_128_srandom_r).
PGM_HD void pcg64_seed(Pcg64& g, uint64_t entropy) {
  const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u;
  const uint32_t INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
  const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
  uint32_t words[2] = {(uint32_t)entropy, (uint32_t)(entropy >> 32)};
  int nwords = words[1] ? 2 : 1;  // int -> uint32 words, little endian, no trailing zeros
  uint32_t pool[4];
  uint32_t hc = INIT_A;
#define PGM_HASHMIX(dst, v)  \
  {                          \
    uint32_t _v = (v) ^ hc;  \
    hc *= MULT_A;            \
    _v *= hc;                \
    _v ^= _v >> 16;          \
    (dst) = _v;              \
  }
  for (int i = 0; i < 4; ++i) PGM_HASHMIX(pool[i], i < nwords ? words[i] : 0u);
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) {
        uint32_t h;
        PGM_HASHMIX(h, pool[s]);
        uint32_t r = MIX_L * pool[d] - MIX_R * h;
        r ^= r >> 16;
        pool[d] = r;
      }
#undef PGM_HASHMIX
  uint32_t hb = INIT_B;
  uint32_t out[8];
  for (int i = 0; i < 8; ++i) {
    uint32_t d = pool[i & 3];
    d ^= hb;
    hb *= MULT_B;
    d *= hb;
    d ^= d >> 16;
    out[i] = d;
  }
  uint64_t s0 = (uint64_t)out[0] | ((uint64_t)out[1] << 32);
  uint64_t s1 = (uint64_t)out[2] | ((uint64_t)out[3] << 32);
  uint64_t s2 = (uint64_t)out[4] | ((uint64_t)out[5] << 32);
  uint64_t s3 = (uint64_t)out[6] | ((uint64_t)out[7] << 32);
  // initstate = (s0 << 64) | s1 ; initseq = (s2 << 64) | s3
  g.state_lo = 0;
  g.state_hi = 0;
  g.inc_hi = (s2 << 1) | (s3 >> 63);
  g.inc_lo = (s3 << 1) | 1u;
  pcg64_step(g);
  uint64_t lo = g.state_lo + s1;
  uint64_t carry = lo < g.state_lo ? 1u : 0u;
  g.state_lo = lo;
  g.state_hi = g.state_hi + s0 + carry;
  pcg64_step(g);
  g.has_uint32 = 0;
  g.uinteger = 0;
}

}  // namespace pgm
