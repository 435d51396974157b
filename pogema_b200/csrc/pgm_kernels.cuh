// pgm_kernels.cuh - the generic fused POGEMA step kernel for sm_100a: every shape GridConfig allows, the reset and
// observe launches of every engine, and the step launches of the shapes the register-resident kernel of
// pgm_fast.cuh does not cover (that kernel shares the helpers below and gives identical results).
//
// One TEAM of threads (a warp, or 64..1024 threads on a named barrier) owns one
// instance for the whole launch; everything an instance needs lives in that
// team's slice of shared memory:
//
//   obst   bit-packed PADDED obstacle map  (cp.async.bulk global->shared, mbarrier;
//          read from global memory instead when two bitmaps do not fit: OG = 1)
//   occ    pre-move occupancy: dense uint16 cell -> agent grid (OCC = 0) or tile
//          buckets behind a pre-move bitmap (OCC = 1), for the conflict lookups
//   abits  bit-packed post-move agent occupancy
//   stage  the observation bit stream of a batch of agents (aliases occ)
//
// A launch advances its instances by num_steps steps (pgm_step: 1, pgm_step_many: K;
// every team then runs its own timeline without any grid-wide barrier).  Phases of
// a step: fills + actions -> occupancy -> move resolution (closed forms of the three
// upstream collision systems, pointer jumping for the dependent chains) ->
// on_target bookkeeping / time limit / auto reset -> observation bits ->
// bit->byte expansion with 16-byte streaming stores.  OP_OBSERVE / OP_RESET reuse
// the state load and the observation phases.
//
// Upstream symbols restated (pure Python upstream; /root/reference holds only
// README.md:1-5, so symbols are cited by name - SURVEY.md section 8a):
//   envs.py :: Pogema.move_agents ('priority' | 'block_both' | 'soft'), _revert_action
//   grid.py :: Grid.move, move_without_checks, on_goal, hide_agent
//   envs.py :: Pogema.step, PogemaLifeLong.step, PogemaCoopFinish.step, update_was_on_goal
//   generator.py :: generate_new_target (numpy Generator.choice on PCG64)
//   wrappers/multi_time_limit.py :: MultiTimeLimit.step
//   wrappers/metrics.py (raw counters only)
//   envs.py :: _get_agents_obs; grid.py :: get_obstacles_for_agent, get_positions, get_square_target
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pgm_rng.h"

namespace pgm {

enum { OP_STEP = 0, OP_OBSERVE = 1, OP_RESET = 2 };
enum { ST_FAIL = 0u, ST_OK = 1u, ST_PEND = 2u };
constexpr uint32_t OCC_NONE = 0xFFFFu;

struct StepArgs {
  // geometry
  int N, A, PH, PW, WPR, r, D;
  int obst_stride;      // uint32 words per instance (multiple of 4, >= PH*WPR + 1)
  int bits_per_agent;   // 3*D*D
  int stage_bpa;        // stage bits per agent: bits_per_agent (U8) or rounded up to 32 (BITS)
  int obs_format;       // 0 u8, 1 bits, 2 float32, 4 float16, 3 raw stream (packed host transport: the stage bit stream of
                        // every observation batch as it is; a batch starts where its first agent's words
                        // would start in format 1)
  int max_steps, auto_reset;
  int on_target;        // 0 finish, 1 nothing, 2 restart
  int batch_agents;     // agents per observation batch (stage capacity)
  int max_rounds;       // pointer jumping round limit
  // state
  const uint32_t* obst;
  uint2* state;         // [N][A] .x = x | active << 15 | y << 16, .y = target x | y << 16  (padded coords)
  const uint2* state0;  // initial task (auto reset / pgm_reset)
  int32_t* elapsed;
  Pcg64* rng;
  const Pcg64* rng0;
  const int32_t* comp_start;
  const int32_t* comp_size;
  const uint32_t* cells;
  long long cells_stride;
  uint8_t* was_on_goal;
  uint8_t* episode_done;
  int32_t* metric_acc;   // [N][4] running: solved, time_sum, cur_step, unused
  int32_t* metric_last;  // [N][4] latched at episode end: solved, time_sum, steps, on_goal_now
  int32_t* solve;        // on_target == nothing only, [N][A][2]: step at which the agent's current stay on its goal
                         // began | its cost in the last finished episode (SumOfCostsAndMakespanMetric); touched
                         // only when an agent arrives and when an episode ends
  // io
  const uint8_t* actions;
  int act_itemsize;
  int num_steps;               // steps advanced by this launch (pgm_step_many)
  long long act_step_stride;   // bytes between the action tensors of consecutive steps
  long long out_step_stride;   // elements between rewards/terminated/truncated of consecutive steps
  int obs_ring;                // step k writes observation slot k % obs_ring
  long long obs_slot_stride;   // bytes between observation slots
  uint8_t* obs;
  long long obs_inst_stride;  // bytes
  float* rewards;
  uint8_t* terminated;
  uint8_t* truncated;
  int* err_flag;
  uint8_t* regen_flag;  // auto_reset == 2: set to 1 when the episode of instance n ended (task is rebuilt after the step)
  const uint8_t* mask;  // OP_OBSERVE only: if not NULL, only instances with mask[n] != 0 are observed (and regen_flag[n] cleared)
  long long* debug;     // optional [N][16] clock64 stamps at phase boundaries (tools/phase_timeline.py)
  // shared memory layout, byte offsets inside a team slice
  int off_pbits;  // OCC == 1: pre-move occupancy bitmap
  int off_obst, off_abits, off_occ, off_pos, off_tgt, off_npos, off_link, off_act, off_flag, off_misc;
  int team_smem;
  int teams_per_cta;
  int occ_tiles, occ_tiles_w, occ_tshift;  // OCC == 1: number of tiles (padded to 4), tiles per row, log2(tile side)
  // pgm_fast_step_kernel (pgm_fast.cuh) only
  int off_stage;    // observation stream of one batch: TEAM agents, every warp owns its word-aligned piece
  int plane_words;  // block_both: words of one claim plane
  int narrow;       // map at most 32 cells wide and two bitmap words per row: one 64-bit load per observation row
};

// ------------------------------------------------------------------------- //
// small PTX helpers
// ------------------------------------------------------------------------- //
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

template <int TEAM>
__device__ __forceinline__ void team_sync(int bar_id) {
  if (TEAM == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TEAM) : "memory");
  }
}
// barrier + OR reduction over the team
template <int TEAM>
__device__ __forceinline__ bool team_any(int bar_id, bool p) {
  if (TEAM == 32) {
    __syncwarp();
    return __any_sync(0xffffffffu, p);
  } else {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred q, s;\n"
        "setp.ne.u32 q, %2, 0;\n"
        "barrier.red.or.pred s, %1, %3, q;\n"
        "selp.u32 %0, 1, 0, s;\n"
        "}"
        : "=r"(r)
        : "r"(bar_id), "r"((uint32_t)p), "n"(TEAM)
        : "memory");
    return r != 0;
  }
}

// Action element idx of a tensor of 1/2/4/8-byte integers, read in full (a wide value such as 256 or -256 must
// not alias to a valid move through its low byte); anything above 4 is reported by the caller's range check.
__device__ __forceinline__ uint32_t load_action(const uint8_t* base, long long idx, int itemsize) {
  if (itemsize == 1) return base[idx];
  if (itemsize == 2) return reinterpret_cast<const uint16_t*>(base)[idx];
  if (itemsize == 4) return reinterpret_cast<const uint32_t*>(base)[idx];
  const unsigned long long v = reinterpret_cast<const unsigned long long*>(base)[idx];
  return v > 4ull ? 5u : (uint32_t)v;
}

__device__ __forceinline__ int opposite(int a) { return a == 0 ? 0 : (((a - 1) ^ 1) + 1); }  // 1<->2, 3<->4
__device__ __forceinline__ int move_dx(int a) { return a == 1 ? -1 : (a == 2 ? 1 : 0); }

// upstream wrappers/metrics.py :: SumOfCostsAndMakespanMetric (on_target == nothing), per agent and step:
//   solve_time is None and (on_goal or finished) -> solve_time = step;   not on_goal and not finished -> None;
//   finished -> cost = solve_time.
// sv[0] holds the step at which the current stay on the goal began (0 after a reset: an agent standing on its goal
// without having arrived has been there since the reset); global memory is touched only on an arrival and at the end
// of an episode.  An agent that steps off its goal on the finishing step keeps the stay's start (upstream's rule).
__device__ __forceinline__ void solve_time_update(int32_t* sv, bool was, bool moved, bool was_before, bool done, int step) {
  const bool arrived = was && moved;
  if (arrived) sv[0] = step;
  if (done) {
    sv[1] = arrived ? step : ((was || was_before) ? sv[0] : step);
    sv[0] = 0;
  }
}
__device__ __forceinline__ int move_dy(int a) { return a == 3 ? -1 : (a == 4 ? 1 : 0); }

__device__ __forceinline__ uint32_t bit_at(const uint32_t* bits, int WPR, int x, int y) {
  return (bits[x * WPR + (y >> 5)] >> (y & 31)) & 1u;
}

// Pre-move occupancy lookup "which active agent stands on cell (x,y)":
//   OCC == 0: dense uint16 grid over the padded map (one LDS per lookup; small maps)
//   OCC == 1: tile buckets - the padded map is cut into 2^t x 2^t tiles, every tile heads a linked list
//             of the agents standing in it (ATOMS.EXCH insert, no probing loops).  Memory ~ cells / 4^t
//             + agents: large maps keep several instances per SM (256x256 with 1024 agents: 4 per SM).
template <int OCC>
struct OccMap {
  uint16_t* dense;       // OCC 0
  uint32_t* heads;       // OCC 1: [tiles] first agent of the tile or 0xFFFFFFFF
  uint16_t* next;        // OCC 1: [A] next agent in the same tile or OCC_NONE
  const uint32_t* pos;   // OCC 1: packed pre-move positions of the agents
  uint32_t* pbits;       // OCC 1: pre-move occupancy bitmap (a miss - the common case - costs one LDS)
  int tshift, tiles_w;
  int PW, WPR;
  __device__ __forceinline__ void insert(int x, int y, uint32_t a) const {
    if (OCC == 0) {
      dense[x * PW + y] = (uint16_t)a;
    } else {
      const uint32_t prev = atomicExch(&heads[(x >> tshift) * tiles_w + (y >> tshift)], a);
      next[a] = (uint16_t)prev;  // 0xFFFFFFFF -> OCC_NONE
      atomicOr(&pbits[x * WPR + (y >> 5)], 1u << (y & 31));
    }
  }
  __device__ __forceinline__ uint32_t lookup(int x, int y) const {
    if (OCC == 0) {
      return dense[x * PW + y];
    } else {
      if (((pbits[x * WPR + (y >> 5)] >> (y & 31)) & 1u) == 0u) return OCC_NONE;
      const uint32_t key = (uint32_t)x | ((uint32_t)y << 16);
      uint32_t k = heads[(x >> tshift) * tiles_w + (y >> tshift)] & 0xFFFFu;
      while (k != OCC_NONE) {
        if (pos[k] == key) return k;
        k = next[k];
      }
      return OCC_NONE;
    }
  }
};

// Is some agent other than the one standing on (sx,sy) heading into (tx,ty)?
// kind 0: any claimant;  kind 1: a claimant with lo < index < hi.
template <int KIND, int OCC>
__device__ __forceinline__ bool other_claimant(const OccMap<OCC>& occ, const uint8_t* act, int tx, int ty, int sx, int sy,
                                               int lo, int hi) {
  bool found = false;
#pragma unroll
  for (int m = 1; m <= 4; ++m) {
    int nx = tx + move_dx(m), ny = ty + move_dy(m);
    if (nx == sx && ny == sy) continue;
    uint32_t k = occ.lookup(nx, ny);
    if (k != OCC_NONE && act[k] == opposite(m)) {
      if (KIND == 0) found = true;
      else if ((int)k > lo && (int)k < hi) found = true;
    }
  }
  return found;
}

// 4 bits -> 4 bytes (bit i -> byte i, value 0/1)
__device__ __forceinline__ uint32_t expand4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// ------------------------------------------------------------------------- //
// observation bits of ONE agent
// ------------------------------------------------------------------------- //
// Generic radius (runtime D): stream the row windows of channel 0 (obstacles)
// and 1 (agents) into the stage bit stream at an arbitrary bit offset.
__device__ __forceinline__ void agent_bits_generic(const StepArgs& p, const uint32_t* s_obst, const uint32_t* s_abits,
                                                   uint32_t* stage, int x, int y, uint32_t bitpos) {
  const int r = p.r, D = p.D, WPR = p.WPR;
  uint32_t widx = bitpos >> 5;
  uint32_t fill = bitpos & 31u;
  unsigned long long acc = 0ull;
  bool first = true;
  const int y0 = y - r;
#pragma unroll 1
  for (int ch = 0; ch < 2; ++ch) {
    const uint32_t* rows = (ch == 0 ? s_obst : s_abits) + (x - r) * WPR;
#pragma unroll 1
    for (int k = 0; k < D; ++k, rows += WPR) {
      for (int c0 = 0; c0 < D; c0 += 32) {
        const int nb = min(32, D - c0);
        const int yy = y0 + c0;
        const int w = yy >> 5, sh = yy & 31;
        uint32_t v = __funnelshift_r(rows[w], rows[w + 1], sh);
        if (nb < 32) v &= (1u << nb) - 1u;
        acc |= (unsigned long long)v << fill;
        fill += nb;
        if (fill >= 32u) {
          if (first) {
            atomicOr(&stage[widx], (uint32_t)acc);
            first = false;
          } else {
            stage[widx] = (uint32_t)acc;
          }
          widx++;
          acc >>= 32;
          fill -= 32u;
        }
      }
    }
  }
  if (fill > 0u) atomicOr(&stage[widx], (uint32_t)acc);
}

// Compile-time radius (D <= 32): the two occupancy channels are assembled in
// registers at static bit offsets (fully unrolled), then stored to the stage
// stream with one funnel shift per word.
template <int D, int OFF>
__device__ __forceinline__ void insert_bits(uint32_t (&acc)[(3 * D * D + 31) / 32], uint32_t v) {
  constexpr int W = OFF >> 5, C = OFF & 31;
  acc[W] |= v << C;
  if (C + D > 32) acc[W + 1] |= v >> ((32 - C) & 31);
}

template <int D, int K>
struct RowUnroll {
  static __device__ __forceinline__ void run(uint32_t (&acc)[(3 * D * D + 31) / 32], const uint32_t* ro,
                                             const uint32_t* ra, int WPR, int sh) {
    constexpr uint32_t MASK = (D == 32) ? 0xFFFFFFFFu : ((1u << D) - 1u);
    const uint32_t vo = __funnelshift_r(ro[0], ro[1], sh) & MASK;
    const uint32_t va = __funnelshift_r(ra[0], ra[1], sh) & MASK;
    insert_bits<D, K * D>(acc, vo);
    insert_bits<D, D * D + K * D>(acc, va);
    RowUnroll<D, K + 1>::run(acc, ro + WPR, ra + WPR, WPR, sh);
  }
};
template <int D>
struct RowUnroll<D, D> {
  static __device__ __forceinline__ void run(uint32_t (&)[(3 * D * D + 31) / 32], const uint32_t*, const uint32_t*,
                                             int, int) {}
};

template <int D>
__device__ __forceinline__ void agent_bits_static(const uint32_t* s_obst, const uint32_t* s_abits, uint32_t* stage,
                                                  int WPR, int x, int y, uint32_t bitpos, bool word_aligned) {
  constexpr int R = D / 2;
  constexpr int NW = (3 * D * D + 31) / 32;
  uint32_t acc[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) acc[i] = 0u;
  const int y0 = y - R;
  const int w = y0 >> 5, sh = y0 & 31;
  const int rowoff = (x - R) * WPR + w;
  RowUnroll<D, 0>::run(acc, s_obst + rowoff, s_abits + rowoff, WPR, sh);
  const uint32_t wbase = bitpos >> 5;
  if (word_aligned) {
#pragma unroll
    for (int i = 0; i < NW; ++i) stage[wbase + i] = acc[i];
  } else {
    const uint32_t s = bitpos & 31u;
    uint32_t prev = 0u;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const uint32_t out = __funnelshift_l(prev, acc[i], s);
      if (i == 0 || i >= NW - 1) atomicOr(&stage[wbase + i], out);
      else stage[wbase + i] = out;
      prev = acc[i];
    }
    const uint32_t tail = __funnelshift_l(prev, 0u, s);
    if (tail) atomicOr(&stage[wbase + NW], tail);
  }
}

// ------------------------------------------------------------------------- //
// observation generation: batches of agents -> stage bit stream -> HBM
// ------------------------------------------------------------------------- //
// Phase stamps for tools/phase_timeline.py (pgm_set_debug_buffer).  They cost ~35 warp instructions per instance and
// step even when no buffer is set, so they are compiled only into the timeline build of the library
// (`make timeline` -> pogema_b200/_lib/libpgm_b200_timeline.so, -DPGM_TIMELINE; load it with PGM_B200_LIB).
#ifdef PGM_TIMELINE
#define PGM_STAMP(k)                        \
  do {                                      \
    if (dbg != nullptr) dbg[(k)] = clock64(); \
  } while (0)
// wall-clock (globaltimer, ns) stamps: comparable across SMs
#define PGM_STAMP_NS(k)                                              \
  do {                                                               \
    if (dbg != nullptr) {                                            \
      unsigned long long _t;                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));       \
      dbg[(k)] = (long long)_t;                                      \
    }                                                                \
  } while (0)
#else
#define PGM_STAMP(k) ((void)dbg)
#define PGM_STAMP_NS(k) ((void)dbg)
#endif

template <int TEAM, int RT>
__device__ __forceinline__ void emit_observations(const StepArgs& p, long long* dbg, uint8_t* obs, int n, int tid,
                                                  int bar_id, const uint32_t* s_obst,
                                                  const uint32_t* s_abits, uint32_t* stage, const uint32_t* s_npos,
                                                  const uint32_t* s_tgt) {
  const int r = (RT > 0) ? RT : p.r;
  const int D = 2 * r + 1;
  const int bpa = 3 * D * D, sbpa = p.stage_bpa;
  const bool word_aligned = (sbpa & 31) == 0;
  for (int g0 = 0; g0 < p.A; g0 += p.batch_agents) {
    const int gcount = min(p.batch_agents, p.A - g0);
    const int nbits = gcount * sbpa;
    const int nvec = ((nbits + 31) / 32 + 1 + 3) >> 2;  // stage is padded to a multiple of 16 bytes
    uint4* stage4 = reinterpret_cast<uint4*>(stage);
    for (int w = tid; w < nvec; w += TEAM) stage4[w] = make_uint4(0u, 0u, 0u, 0u);
    team_sync<TEAM>(bar_id);
    PGM_STAMP(6);
    for (int s = tid; s < gcount; s += TEAM) {
      const int a = g0 + s;
      const uint32_t pp = s_npos[a];
      const int x = pp & 0xFFFF, y = pp >> 16;
      const uint32_t bitpos = (uint32_t)s * (uint32_t)sbpa;
      if (RT > 0) agent_bits_static<2 * (RT > 0 ? RT : 1) + 1>(s_obst, s_abits, stage, p.WPR, x, y, bitpos, word_aligned);
      else agent_bits_generic(p, s_obst, s_abits, stage, x, y, bitpos);
      // channel 2: upstream grid.py :: get_square_target (clamped projection of the goal)
      const uint32_t tt = s_tgt[a];
      int dx = x - (int)(tt & 0xFFFF), dy = y - (int)(tt >> 16);
      dx = max(-r, min(r, dx));
      dy = max(-r, min(r, dy));
      const uint32_t tb = bitpos + 2u * D * D + (uint32_t)(r - dx) * D + (uint32_t)(r - dy);
      atomicOr(&stage[tb >> 5], 1u << (tb & 31u));
    }
    team_sync<TEAM>(bar_id);
    PGM_STAMP(7);
    // ---- write out
    if (p.obs_format & 1) {
      // 1: bits, 32-bit words per agent.  3 (packed host transport, pgm_step_host): the stage bit stream of the
      // batch as it is (agents bit-contiguous); batch b starts at the word offset agent b*batch would have in
      // format 1, so both formats share this loop.
      const int wpa = (bpa + 31) >> 5;
      uint32_t* out = reinterpret_cast<uint32_t*>(obs + (long long)n * p.obs_inst_stride) + (long long)g0 * wpa;
      const int nw = (nbits + 31) >> 5;
      for (int w = tid; w < nw; w += TEAM) __stcs(out + w, stage[w]);
    } else if (p.obs_format == 2) {
      // float32 0.0 / 1.0 (the reference's observation dtype): one stream bit -> one float, 16-byte stores
      float* out = reinterpret_cast<float*>(obs + (long long)n * p.obs_inst_stride) + (long long)g0 * bpa;
      const int nfl = gcount * bpa;
      int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u)) & 15u) >> 2);
      head = min(head, nfl);
      const int chunks = (nfl - head) >> 2;
      uint4* out16 = reinterpret_cast<uint4*>(out + head);
      for (int c = tid; c < chunks; c += TEAM) {
        const uint32_t bit = (uint32_t)head + ((uint32_t)c << 2);
        const uint32_t w = bit >> 5, sh = bit & 31u;
        const uint32_t v = __funnelshift_r(stage[w], stage[w + 1], sh);
        uint4 o;
        o.x = (v & 1u) ? 0x3F800000u : 0u;
        o.y = (v & 2u) ? 0x3F800000u : 0u;
        o.z = (v & 4u) ? 0x3F800000u : 0u;
        o.w = (v & 8u) ? 0x3F800000u : 0u;
        __stcs(out16 + c, o);
      }
      const int tail0 = head + (chunks << 2);
      for (int b = tid; b < head + (nfl - tail0); b += TEAM) {
        const int bb = b < head ? b : tail0 + (b - head);
        out[bb] = ((stage[bb >> 5] >> (bb & 31)) & 1u) ? 1.0f : 0.0f;
      }
    } else if (p.obs_format != 4) {
      uint8_t* out = obs + (long long)n * p.obs_inst_stride + (long long)g0 * bpa;
      const int nbytes = gcount * bpa;
      int head = (int)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u)) & 15u);
      head = min(head, nbytes);
      const int chunks = (nbytes - head) >> 4;
      uint4* out16 = reinterpret_cast<uint4*>(out + head);
      if (head == 0) {
        // aligned fast path: chunk c <- 16 stage bits at halfword c
        const uint16_t* st16 = reinterpret_cast<const uint16_t*>(stage);
#pragma unroll 4
        for (int c = tid; c < chunks; c += TEAM) {
          const uint32_t v = st16[c];
          uint4 o;
          o.x = expand4(v & 15u);
          o.y = expand4((v >> 4) & 15u);
          o.z = expand4((v >> 8) & 15u);
          o.w = expand4(v >> 12);
          __stcs(out16 + c, o);
        }
      } else {
        for (int c = tid; c < chunks; c += TEAM) {
          const uint32_t bit = (uint32_t)head + ((uint32_t)c << 4);
          const uint32_t w = bit >> 5, sh = bit & 31u;
          const uint32_t v = __funnelshift_r(stage[w], stage[w + 1], sh);
          uint4 o;
          o.x = expand4(v & 15u);
          o.y = expand4((v >> 4) & 15u);
          o.z = expand4((v >> 8) & 15u);
          o.w = expand4((v >> 12) & 15u);
          __stcs(out16 + c, o);
        }
      }
      const int tail0 = head + (chunks << 4);
      for (int b = tid; b < head + (nbytes - tail0); b += TEAM) {
        const int bb = b < head ? b : tail0 + (b - head);
        out[bb] = (uint8_t)((stage[bb >> 5] >> (bb & 31)) & 1u);
      }
    } else {
      // float16 0.0 / 1.0 for half-precision policies (no cast pass over the tensor): 8 stream bits -> 8 halves
      uint16_t* out = reinterpret_cast<uint16_t*>(obs + (long long)n * p.obs_inst_stride) + (long long)g0 * bpa;
      const int nel = gcount * bpa;
      int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u)) & 15u) >> 1);
      head = min(head, nel);
      const int chunks = (nel - head) >> 3;
      uint4* out16 = reinterpret_cast<uint4*>(out + head);
      for (int c = tid; c < chunks; c += TEAM) {
        const uint32_t bit = (uint32_t)head + ((uint32_t)c << 3);
        const uint32_t w = bit >> 5, sh = bit & 31u;
        const uint32_t v = __funnelshift_r(stage[w], stage[w + 1], sh);
        uint4 o;  // two bits -> bits 0 and 16, times 0x3C00 (= 1.0 in binary16) in both halves
        o.x = ((v & 1u) | ((v & 2u) << 15)) * 0x3C00u;
        o.y = (((v >> 2) & 1u) | ((v & 8u) << 13)) * 0x3C00u;
        o.z = (((v >> 4) & 1u) | ((v & 32u) << 11)) * 0x3C00u;
        o.w = (((v >> 6) & 1u) | ((v & 128u) << 9)) * 0x3C00u;
        __stcs(out16 + c, o);
      }
      const int tail0 = head + (chunks << 3);
      for (int b = tid; b < head + (nel - tail0); b += TEAM) {
        const int bb = b < head ? b : tail0 + (b - head);
        out[bb] = ((stage[bb >> 5] >> (bb & 31)) & 1u) ? (uint16_t)0x3C00u : (uint16_t)0u;
      }
    }
    if (g0 + p.batch_agents < p.A) team_sync<TEAM>(bar_id);
  }
  PGM_STAMP(8);
  PGM_STAMP_NS(11);
}

// ------------------------------------------------------------------------- //
// the fused step kernel
// ------------------------------------------------------------------------- //
// Agent state word (uint2 per agent): .x = x | active << 15 | y << 16 (padded
// coordinates), .y = target x | y << 16.
__device__ __forceinline__ uint32_t st_pos(uint32_t w) { return w & 0xFFFF7FFFu; }
__device__ __forceinline__ uint32_t st_active(uint32_t w) { return (w >> 15) & 1u; }

// Programmatic dependent launch (PTX griddepcontrol): the launch latency and the
// prologue of step t+1 (shared memory fills, obstacle bulk copy) hide under step t.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// OG == 1 (maps whose two bitmaps do not fit in shared memory, e.g. 1024x1024): the obstacle bitmap is
// read from global memory / L2 instead of being staged, and the pre-move bitmap shares storage with the
// post-move one.
template <int TEAM, int COLL, int OP, int RT, int OCC, int OG>
__global__ void __launch_bounds__(1024, 1)
    pgm_step_kernel(const StepArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int team = threadIdx.x / TEAM;
  const int tid = threadIdx.x % TEAM;
  const int n = blockIdx.x * p.teams_per_cta + team;
  if (n >= p.N) return;  // whole team leaves together
  if (OP == OP_OBSERVE && p.mask != nullptr && p.mask[n] == 0) return;
  const int bar_id = 1 + team;
  unsigned char* base = smem_raw + (size_t)team * p.team_smem;
  const uint32_t* s_obst = OG ? (p.obst + (long long)n * p.obst_stride)
                              : reinterpret_cast<const uint32_t*>(base + p.off_obst);
  uint32_t* s_abits = reinterpret_cast<uint32_t*>(base + p.off_abits);
  OccMap<OCC> occ;
  occ.dense = reinterpret_cast<uint16_t*>(base + p.off_occ);
  occ.heads = reinterpret_cast<uint32_t*>(base + p.off_occ);
  occ.next = reinterpret_cast<uint16_t*>(base + p.off_occ + 4 * p.occ_tiles);
  occ.pbits = reinterpret_cast<uint32_t*>(base + p.off_pbits);
  occ.tshift = p.occ_tshift;
  occ.tiles_w = p.occ_tiles_w;
  occ.PW = p.PW;
  occ.WPR = p.WPR;
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(base + p.off_occ);  // aliases occ
  uint32_t* s_pos = reinterpret_cast<uint32_t*>(base + p.off_pos);
  occ.pos = s_pos;
  uint32_t* s_tgt = reinterpret_cast<uint32_t*>(base + p.off_tgt);
  uint32_t* s_npos = reinterpret_cast<uint32_t*>(base + p.off_npos);
  uint32_t* s_link = reinterpret_cast<uint32_t*>(base + p.off_link);
  uint8_t* s_act = base + p.off_act;
  uint8_t* s_flag = base + p.off_flag;  // bit0 active, bit1 on_goal(after move), bit2 was_on_goal
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(base + p.off_misc);
  int* s_cnt = reinterpret_cast<int*>(base + p.off_misc + 8);  // [0] on_goal count [1] was_on_goal count

  const int A = p.A, PW = p.PW, WPR = p.WPR;
  const int ONTGT = p.on_target;
  const long long ia = (long long)n * A;
  long long* dbg = (p.debug != nullptr && tid == 0) ? p.debug + (long long)n * 16 : nullptr;  // phase stamps

  // Let the next launch in the stream be scheduled as soon as SM resources free up: its
  // prologue (phase 0a) overlaps this grid's tail; its griddepcontrol.wait still waits for
  // this grid to complete and flush, so there is no cross-launch race on any buffer.
  pdl_trigger();
  PGM_STAMP(0);
  PGM_STAMP_NS(9);
  // ---- phase 0a (independent of the previous launch): obstacle map by bulk copy
  if (!OG && tid == 0) {
    mbar_init(s_bar, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)p.obst_stride * 4u;
    mbar_expect_tx(s_bar, bytes);
    bulk_g2s(base + p.off_obst, p.obst + (long long)n * p.obst_stride, bytes, s_bar);
  }
  // per-step fills: bitmaps to zero, occupancy structure to "empty" (both regions are padded to multiples of
  // 16 bytes by the host-side layout).  The first step's fills do not depend on the previous launch.
  auto step_fills = [&]() {
    const int abits_vec = (p.PH * WPR + 1 + 3) >> 2;
    uint4* a4 = reinterpret_cast<uint4*>(s_abits);
    for (int w = tid; w < abits_vec; w += TEAM) a4[w] = make_uint4(0u, 0u, 0u, 0u);
    if (OP == OP_STEP) {
      // dense: every cell = OCC_NONE; buckets: every tile head = empty (both are all-ones fills)
      const int occ_vec = (OCC == 0) ? ((p.PH * PW * 2 + 4 + 15) >> 4) : ((p.occ_tiles + 3) >> 2);
      uint4* o4 = reinterpret_cast<uint4*>(occ.dense);
      for (int w = tid; w < occ_vec; w += TEAM) o4[w] = make_uint4(~0u, ~0u, ~0u, ~0u);
      if (OCC == 1 && p.off_pbits != p.off_abits) {
        uint4* p4 = reinterpret_cast<uint4*>(occ.pbits);
        for (int w = tid; w < abits_vec; w += TEAM) p4[w] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (tid == 0) {
        s_cnt[0] = 0;
        s_cnt[1] = 0;
      }
    }
  };
  step_fills();
  // ---- phase 0b: mutable state of this instance (two agents per thread in flight)
  PGM_STAMP(1);
  pdl_wait();
  PGM_STAMP(2);
  PGM_STAMP_NS(10);
  int step_idx = 0;
  int m_acc0 = 0, m_acc1 = 0, m_acc2 = 0;
  if (OP == OP_STEP) {
    step_idx = p.elapsed[n];
    const int4 m = *reinterpret_cast<const int4*>(p.metric_acc + 4 * (long long)n);
    m_acc0 = m.x;
    m_acc1 = m.y;
    m_acc2 = m.z;
  }
  const uint2* src = (OP == OP_RESET) ? p.state0 : p.state;
  for (int a0 = tid; a0 < A; a0 += 2 * TEAM) {
    const int a1 = a0 + TEAM;
    const bool has1 = a1 < A;
    uint2 w0 = src[ia + a0];
    uint2 w1 = has1 ? src[ia + a1] : make_uint2(0u, 0u);
    if (OP == OP_RESET) {
      w0.x |= 0x8000u;
      w1.x |= 0x8000u;
    }
    s_pos[a0] = st_pos(w0.x);
    s_npos[a0] = st_pos(w0.x);
    s_tgt[a0] = w0.y;
    s_flag[a0] = (uint8_t)st_active(w0.x);
    if (has1) {
      s_pos[a1] = st_pos(w1.x);
      s_npos[a1] = st_pos(w1.x);
      s_tgt[a1] = w1.y;
      s_flag[a1] = (uint8_t)st_active(w1.x);
    }
  }
  team_sync<TEAM>(bar_id);  // the mbarrier was initialised by thread 0 of the team
  if (!OG) mbar_wait(s_bar, 0);

  // One launch advances this instance by p.num_steps steps (1 for pgm_step; >1 for pgm_step_many,
  // where every team runs its own timeline: no grid-wide barrier between steps, the observation
  // stores of one instance overlap the move phases of the others).
  const int num_steps = (OP == OP_STEP) ? p.num_steps : 1;
  int obs_slot = 0;  // k % obs_ring without a division per step
#pragma unroll 1
  for (int k = 0; k < num_steps; ++k) {
    uint8_t* obs_k = p.obs;
    if (OP == OP_STEP) {
      // actions of step k
      const uint8_t* act_k = p.actions + (long long)k * p.act_step_stride;
      for (int a0 = tid; a0 < A; a0 += 2 * TEAM) {
        const int a1 = a0 + TEAM;
        const bool has1 = a1 < A;
        uint32_t act0 = load_action(act_k, ia + a0, p.act_itemsize);
        uint32_t act1 = has1 ? load_action(act_k, ia + a1, p.act_itemsize) : 0u;
        if (act0 > 4u || act1 > 4u) {
          atomicOr(p.err_flag, 1);
          if (act0 > 4u) act0 = 0u;
          if (act1 > 4u) act1 = 0u;
        }
        s_act[a0] = (uint8_t)act0;
        if (has1) s_act[a1] = (uint8_t)act1;
      }
      if (p.obs != nullptr) obs_k = p.obs + (long long)obs_slot * p.obs_slot_stride;
      if (++obs_slot == p.obs_ring) obs_slot = 0;
    }
    team_sync<TEAM>(bar_id);

    if (OP == OP_STEP) {
      // ---- phase 1: pre-move occupancy grid (active agents only) ------------
      for (int a = tid; a < A; a += TEAM) {
        if (s_flag[a] & 1u) {
          const uint32_t pp = s_pos[a];
          occ.insert(pp & 0xFFFF, pp >> 16, a);
        }
      }
      team_sync<TEAM>(bar_id);
      PGM_STAMP(3);

      // ---- phase 2: move resolution -----------------------------------------
      if (COLL == 2) {
        // soft, pass A: moves into obstacles and edge swaps become 'stay'
        for (int a = tid; a < A; a += TEAM) {
          uint32_t eff = 0u;
          const uint32_t act = s_act[a];
          if ((s_flag[a] & 1u) && act != 0u) {
            const uint32_t pp = s_pos[a];
            const int tx = (int)(pp & 0xFFFF) + move_dx(act), ty = (int)(pp >> 16) + move_dy(act);
            eff = act;
            if (bit_at(s_obst, WPR, tx, ty)) {
              eff = 0u;
            } else {
              const uint32_t j = occ.lookup(tx, ty);
              if (j != OCC_NONE && s_act[j] == opposite(act)) eff = 0u;
            }
          }
          s_link[a] = eff;  // temporarily: effective action
        }
        team_sync<TEAM>(bar_id);
        // the effective actions replace the raw ones (each thread rewrites its own entries)
        for (int a = tid; a < A; a += TEAM) s_act[a] = (uint8_t)s_link[a];
        team_sync<TEAM>(bar_id);
      }
      bool pend = false;
      for (int a = tid; a < A; a += TEAM) {
        uint32_t link = ST_FAIL;
        const uint32_t act = s_act[a];
        if ((s_flag[a] & 1u) && act != 0u) {
          const uint32_t pp = s_pos[a];
          const int sx = pp & 0xFFFF, sy = pp >> 16;
          const int tx = sx + move_dx(act), ty = sy + move_dy(act);
          if (!bit_at(s_obst, WPR, tx, ty)) {
            const uint32_t j = occ.lookup(tx, ty);
            if (COLL == 1) {
              // block_both: free cell, sole claimant
              if (j == OCC_NONE && !other_claimant<0, OCC>(occ, s_act, tx, ty, sx, sy, 0, 0)) link = ST_OK;
            } else if (COLL == 0) {
              // priority: occupant must have a lower index and leave; first claimant above it wins
              bool ok = true;
              int lo = -1;
              if (j != OCC_NONE) {
                if ((int)j > a || s_act[j] == 0) ok = false;
                else lo = (int)j;
              }
              if (ok && other_claimant<1, OCC>(occ, s_act, tx, ty, sx, sy, lo, a)) ok = false;
              if (ok) link = (j == OCC_NONE) ? ST_OK : (ST_PEND | (j << 2));
            } else {
              // soft: no stayer on the cell, lowest-index claimant, occupant must leave
              bool ok = !(j != OCC_NONE && s_act[j] == 0);
              if (ok && other_claimant<1, OCC>(occ, s_act, tx, ty, sx, sy, -1, a)) ok = false;
              if (ok) link = (j == OCC_NONE) ? ST_OK : (ST_PEND | (j << 2));
            }
          }
        }
        s_link[a] = link;
        pend |= ((link & 3u) == ST_PEND);
      }
      // pointer jumping along occupant chains, double buffered (s_link <-> s_npos, which is free until
      // phase 3) so that a round only reads what the previous round wrote; what is still pending after
      // max_rounds is a rotation cycle (soft only) and succeeds
      uint32_t* link_cur = s_link;
      if (COLL != 1) {
        uint32_t* link_nxt = s_npos;
        int rounds = 0;
        while (team_any<TEAM>(bar_id, pend)) {
          if (++rounds > p.max_rounds) break;
          pend = false;
          for (int a = tid; a < A; a += TEAM) {
            uint32_t l = link_cur[a];
            if ((l & 3u) == ST_PEND) {
              const uint32_t lp = link_cur[l >> 2];
              l = ((lp & 3u) == ST_PEND) ? (ST_PEND | (lp & ~3u)) : (lp & 3u);
              pend |= ((l & 3u) == ST_PEND);
            }
            link_nxt[a] = l;
          }
          uint32_t* t = link_cur;
          link_cur = link_nxt;
          link_nxt = t;
        }
      }
      team_sync<TEAM>(bar_id);
      if (OCC == 1 && p.off_pbits == p.off_abits) {
        // the pre-move bitmap shared the storage of the post-move one: clear it again
        const int abits_vec = (p.PH * WPR + 1 + 3) >> 2;
        uint4* a4 = reinterpret_cast<uint4*>(s_abits);
        for (int w = tid; w < abits_vec; w += TEAM) a4[w] = make_uint4(0u, 0u, 0u, 0u);
        team_sync<TEAM>(bar_id);
      }
      PGM_STAMP(4);

      // ---- phase 3: apply moves, on_target bookkeeping, time limit -----------
      int c_on = 0, c_was = 0;
      for (int a = tid; a < A; a += TEAM) {
        const uint32_t act = s_act[a];
        uint32_t pp = s_pos[a];
        if ((link_cur[a] & 3u) != ST_FAIL) {
          const int tx = (int)(pp & 0xFFFF) + move_dx(act), ty = (int)(pp >> 16) + move_dy(act);
          pp = (uint32_t)tx | ((uint32_t)ty << 16);
        }
        s_npos[a] = pp;
        const uint32_t on = (pp == s_tgt[a]) ? 1u : 0u;
        const uint32_t fl = s_flag[a] & 1u;
        const uint32_t was = on & fl;
        s_flag[a] = (uint8_t)(fl | (on << 1) | (was << 2));
        c_on += on;
        c_was += was;
      }
      c_on = __reduce_add_sync(0xffffffffu, c_on);
      c_was = __reduce_add_sync(0xffffffffu, c_was);
      if (TEAM > 32) {
        if ((threadIdx.x & 31) == 0) {
          atomicAdd(&s_cnt[0], c_on);
          atomicAdd(&s_cnt[1], c_was);
        }
        team_sync<TEAM>(bar_id);
        c_on = s_cnt[0];
        c_was = s_cnt[1];
      }
      // (all reads of occ are done: the barrier after pointer jumping; stage may reuse it)
      const bool trunc = (step_idx + 1 >= p.max_steps);
      const bool solved = (c_was == A);
      const bool all_term = (ONTGT == 2) ? false : (c_on == A);
      const bool done = trunc || all_term;
      const bool do_reset = done && p.auto_reset == 1;
      const bool reseed = done && p.auto_reset == 2;  // new task from a new seed: built after this launch
      const long long oa = ia + (long long)k * p.out_step_stride;  // outputs of step k

      for (int a = tid; a < A; a += TEAM) {
        const uint32_t f = s_flag[a];
        const uint32_t fl = f & 1u, on = (f >> 1) & 1u, was = (f >> 2) & 1u;
        float rew;
        uint8_t term;
        uint32_t nfl = fl;
        uint32_t tt = s_tgt[a];
        if (ONTGT == 0) {  // finish: reward once, agent disappears
          rew = was ? 1.0f : 0.0f;
          term = (uint8_t)on;
          nfl = fl & (on ^ 1u);
        } else if (ONTGT == 1) {  // nothing (cooperative finish)
          rew = solved ? 1.0f : 0.0f;
          term = solved ? 1 : 0;
        } else {  // restart (lifelong): new target from the agent's own generator
          rew = was ? 1.0f : 0.0f;
          term = 0;
          if (on && !do_reset) {
            Pcg64 g = p.rng[ia + a];
            const uint32_t kk = pcg64_bounded32(g, (uint32_t)(p.comp_size[ia + a] - 1));
            tt = p.cells[(long long)n * p.cells_stride + p.comp_start[ia + a] + kk];
            p.rng[ia + a] = g;
          }
        }
        // evict-first like the observations: written once, read by the caller - as ordinary lines the rewards and
        // flags of a rollout pile up in L2 at the expense of the observation stream (pgm_fast.cuh has the numbers)
        __stcs(p.rewards + oa + a, rew);
        __stcs(p.terminated + oa + a, term);
        __stcs(p.truncated + oa + a, (uint8_t)(trunc ? 1 : 0));
        p.was_on_goal[ia + a] = (uint8_t)was;
        uint32_t pp = s_npos[a];
        if (ONTGT == 1) solve_time_update(p.solve + 2 * (ia + a), was != 0u, pp != s_pos[a], s_pos[a] == tt, done, m_acc2);
        if (do_reset) {
          const uint2 w = p.state0[ia + a];
          pp = st_pos(w.x);
          tt = w.y;
          nfl = 1u;
          if (ONTGT == 2) p.rng[ia + a] = p.rng0[ia + a];
        }
        s_npos[a] = pp;
        s_pos[a] = pp;  // next step of a multi-step launch starts here
        s_tgt[a] = tt;
        s_flag[a] = (uint8_t)nfl;
        p.state[ia + a] = make_uint2(pp | (nfl << 15), tt);
      }
      {
        // raw counters of upstream wrappers/metrics.py (kept in registers across the steps of a launch)
        const int mstep = m_acc2;
        const int solved_sum = m_acc0 + c_was;
        const int time_sum = m_acc1 + c_was * mstep;
        if (done) {
          if (tid == 0) {
            int4* last = reinterpret_cast<int4*>(p.metric_last + 4 * (long long)n);
            *last = make_int4(solved_sum, time_sum + (A - solved_sum) * mstep, mstep + 1, c_was);
          }
          m_acc0 = 0;
          m_acc1 = 0;
          m_acc2 = 0;
        } else {
          m_acc0 = solved_sum;
          m_acc1 = time_sum;
          m_acc2 = mstep + 1;
        }
        step_idx = do_reset ? 0 : step_idx + 1;
      if (reseed) obs_k = nullptr;  // the observation of the rebuilt task is written by the masked observe pass
        if (tid == 0) {
          p.elapsed[n] = step_idx;
          p.episode_done[n] = done ? 1 : 0;
          if (reseed) p.regen_flag[n] = 1;
          *reinterpret_cast<int4*>(p.metric_acc + 4 * (long long)n) = make_int4(m_acc0, m_acc1, m_acc2, 0);
        }
      }
    } else if (OP == OP_RESET) {
      for (int a = tid; a < A; a += TEAM) {
        p.state[ia + a] = make_uint2(s_pos[a] | 0x8000u, s_tgt[a]);
        p.was_on_goal[ia + a] = (s_pos[a] == s_tgt[a]) ? 1 : 0;
        if (ONTGT == 2) p.rng[ia + a] = p.rng0[ia + a];
        if (ONTGT == 1) p.solve[2 * (ia + a)] = 0;
      }
      if (tid == 0) {
        p.elapsed[n] = 0;
        p.episode_done[n] = 0;
        *reinterpret_cast<int4*>(p.metric_acc + 4 * (long long)n) = make_int4(0, 0, 0, 0);
      }
    }
    PGM_STAMP(5);
    if (obs_k != nullptr) {
      // ---- phase 4: post-move agent bitmap -------------------------------------
      for (int a = tid; a < A; a += TEAM) {
        if (s_flag[a] & 1u) {
          const uint32_t pp = s_npos[a];
          const int x = pp & 0xFFFF, y = pp >> 16;
          atomicOr(&s_abits[x * WPR + (y >> 5)], 1u << (y & 31));
        }
      }
      // (emit_observations starts with stage zeroing + team_sync, which also orders the atomics)
      // ---- phase 5/6: observation bits, expansion, stores -------------------------
      emit_observations<TEAM, RT>(p, dbg, obs_k, n, tid, bar_id, s_obst, s_abits, s_stage, s_npos, s_tgt);
      if (OP == OP_OBSERVE && p.mask != nullptr && tid == 0) p.regen_flag[n] = 0;
    }
    if (k + 1 < num_steps) {
      team_sync<TEAM>(bar_id);  // stage (aliasing occ) and s_act are rewritten next
      step_fills();
    }
  }
}

}  // namespace pgm
