// pgm_fast.cuh - the register-resident step kernel for the common shapes (sm_100a).
//
// Same step semantics and the same outputs, bit for bit, as pgm_step_kernel (pgm_kernels.cuh), for the shapes
// the planner marks `fast` (pgm_plan.cu :: plan_fast): compile-time radius 2..7, at most 4 agents per thread
// (APT), at most 8191 agents, uint8 / bit-packed observations, 16-byte aligned observation blocks.  What differs
// is how the work is laid out - the generic kernel is bound by issue slots, not by HBM, on single-step launches
// and on small radii (profiles/r01_single_step_instruction_mix.txt), so this one is written to issue less:
//
//   * agent state (position, target, flags, action) lives in REGISTERS for the whole launch: thread t of a team
//     owns agents t, t + TEAM, ... (APT of them); shared memory only holds what other threads must see;
//   * 'priority' / 'soft': one persistent uint16 cell grid whose entries carry the occupant AND its action
//     (index | action << 13).  A conflict probe is one LDS and one range compare; the grid is never refilled -
//     an agent clears the cell it leaves;
//   * 'block_both': no indices at all - the pre-move occupancy is last step's agent bitmap, claims are two bit
//     planes ("claimed", "claimed twice");
//   * observations: a thread assembles its agent's 3*D*D bits in registers, shifts them to the agent's offset in
//     the batch bit stream and stores whole words; the word two neighbouring agents share travels by warp shuffle.
//     No zero fill and no atomics on the stream (the generic kernel spends 10 ATOMS per warp and step there);
//   * an observation batch is one agent per thread; a warp's 32 agents are a word-aligned piece of the batch stream,
//     so each warp assembles, expands and stores its own piece (16-byte streaming stores, 512 contiguous bytes per
//     instruction) without a team barrier: the stores of one warp overlap the bit assembly of the others.
//
// Upstream symbols restated: see pgm_kernels.cuh (same list; /root/reference holds only README.md:1-5).
#pragma once
#include "pgm_kernels.cuh"

namespace pgm {

constexpr uint32_t G_NONE = 0xFFFFu;  // grid entry: no active agent on the cell
constexpr int G_ACT = 13;             // grid entry = agent index | action << 13   (index <= 8190)

// ---- observation rows ------------------------------------------------------------------------------------
// NARROW (maps at most 32 cells wide, two bitmap words per row): a row window never starts beyond bit 31, one
// 64-bit load brings both words.  Otherwise two 32-bit loads around the window's first word.
template <int D, int K, bool NARROW>
struct FastRows {
  static __device__ __forceinline__ void run(uint32_t (&acc)[(3 * D * D + 31) / 32], const uint32_t* ro,
                                             const uint32_t* ra, int WPR, int sh) {
    constexpr uint32_t MASK = (1u << D) - 1u;
    uint32_t vo, va;
    if (NARROW) {
      const uint2 o = *reinterpret_cast<const uint2*>(ro);
      const uint2 a = *reinterpret_cast<const uint2*>(ra);
      vo = __funnelshift_r(o.x, o.y, sh) & MASK;
      va = __funnelshift_r(a.x, a.y, sh) & MASK;
    } else {
      vo = __funnelshift_r(ro[0], ro[1], sh) & MASK;
      va = __funnelshift_r(ra[0], ra[1], sh) & MASK;
    }
    insert_bits<D, K * D>(acc, vo);
    insert_bits<D, D * D + K * D>(acc, va);
    FastRows<D, K + 1, NARROW>::run(acc, ro + WPR, ra + WPR, WPR, sh);
  }
};
template <int D, bool NARROW>
struct FastRows<D, D, NARROW> {
  static __device__ __forceinline__ void run(uint32_t (&)[(3 * D * D + 31) / 32], const uint32_t*, const uint32_t*,
                                             int, int) {}
};

// The three channels of one agent (upstream envs.py :: _get_agents_obs) as 3*D*D bits in registers.
template <int D, bool NARROW>
__device__ __forceinline__ void fast_agent_bits(uint32_t (&acc)[(3 * D * D + 31) / 32], const uint32_t* s_obst,
                                                const uint32_t* s_abits, int WPR, uint32_t pos, uint32_t tgt) {
  constexpr int R = D / 2;
  constexpr int NW = (3 * D * D + 31) / 32;
#pragma unroll
  for (int i = 0; i < NW; ++i) acc[i] = 0u;
  const int x = pos & 0xFFFF, y = pos >> 16;
  const int y0 = y - R;
  if (NARROW) {
    const int rowoff = (x - R) * 2;
    FastRows<D, 0, true>::run(acc, s_obst + rowoff, s_abits + rowoff, 2, y0);
  } else {
    const int rowoff = (x - R) * WPR + (y0 >> 5);
    FastRows<D, 0, false>::run(acc, s_obst + rowoff, s_abits + rowoff, WPR, y0 & 31);
  }
  // channel 2: upstream grid.py :: get_square_target (clamped projection of the goal)
  int dx = x - (int)(tgt & 0xFFFF), dy = y - (int)(tgt >> 16);
  dx = max(-R, min(R, dx));
  dy = max(-R, min(R, dy));
  const uint32_t tb = 2u * D * D + (uint32_t)(R - dx) * D + (uint32_t)(R - dy);
  const uint32_t tw = tb >> 5, tm = 1u << (tb & 31u);
#pragma unroll
  for (int i = (2 * D * D) >> 5; i < NW; ++i) acc[i] |= (tw == (uint32_t)i) ? tm : 0u;
}

// Store an agent's bits at bit offset `bitoff` of the batch stream (the agent's slot is `sbpa` bits wide; slots
// of consecutive lanes are adjacent, a warp's first slot starts on a word).  Every stream word is written exactly
// once: the word two lanes share is completed and stored by the upper lane, which gets the lower lane's part by
// shuffle.  Must be called by all 32 lanes (`present` = this lane has an agent, `next_present` = so has lane + 1).
template <int NW>
__device__ __forceinline__ void fast_store_stream(uint32_t* stage, const uint32_t (&acc)[NW], uint32_t bitoff,
                                                  uint32_t sbpa, bool present, bool next_present, int lane) {
  const uint32_t sh = bitoff & 31u, w0 = bitoff >> 5;
  const uint32_t e = sh + sbpa;            // bits from the start of word w0 to the end of this slot
  const uint32_t nwords = (e + 31u) >> 5;  // words touched: NW or NW + 1
  const bool partial = (e & 31u) != 0u;    // the last word is shared with the next lane
  uint32_t X[NW + 1];
  uint32_t prev = 0u;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    X[i] = __funnelshift_l(prev, acc[i], sh);
    prev = acc[i];
  }
  X[NW] = __funnelshift_l(prev, 0u, sh);
  uint32_t carry = (nwords == (uint32_t)(NW + 1)) ? X[NW] : X[NW - 1];
  if (!partial || !present) carry = 0u;
  uint32_t carry_in = __shfl_up_sync(0xffffffffu, carry, 1);
  if (lane == 0) carry_in = 0u;
  if (present) {
    X[0] |= carry_in;
    const uint32_t nstore = nwords - ((partial && next_present) ? 1u : 0u);
#pragma unroll
    for (int i = 0; i <= NW; ++i)
      if ((uint32_t)i < nstore) stage[w0 + i] = X[i];
  }
}

// Expansion of `nbytes` stream bits (a multiple of 16, 16-byte aligned destination) into uint8 0/1: a lane turns
// 16 stream bits into 16 bytes ((nibble * 0x00204081) & 0x01010101 per word) and issues one 16-byte streaming store;
// a warp writes 512 contiguous bytes per instruction.  (A 256-entry shared-memory table byte -> 8 bytes halves the
// instructions of this loop and was measured 10-30 % SLOWER: the two extra 64-bit loads per store bank-conflict.)
template <int TEAM>
__device__ __forceinline__ void fast_expand_u8(const uint32_t* stage, uint8_t* out, int nbytes, int tid) {
  const uint16_t* st16 = reinterpret_cast<const uint16_t*>(stage);
  uint4* out16 = reinterpret_cast<uint4*>(out);
  const int chunks = nbytes >> 4;
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
}

// ------------------------------------------------------------------------------------------------------------
template <int TEAM, int APT, int COLL, int RT>
__global__ void __launch_bounds__(1024, 1) pgm_fast_step_kernel(const StepArgs p) {
  constexpr int D = 2 * RT + 1;
  constexpr int BPA = 3 * D * D;
  constexpr int NW = (BPA + 31) / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int team = threadIdx.x / TEAM;
  const int tid = threadIdx.x % TEAM;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * p.teams_per_cta + team;
  if (n >= p.N) return;  // whole team leaves together
  const int bar_id = 1 + team;
  unsigned char* base = smem_raw + (size_t)team * p.team_smem;
  const uint32_t* s_obst = reinterpret_cast<const uint32_t*>(base + p.off_obst);
  uint32_t* const s_abits0 = reinterpret_cast<uint32_t*>(base + p.off_abits);
  uint32_t* const s_abits1 = reinterpret_cast<uint32_t*>(base + p.off_pbits);  // block_both only
  uint16_t* s_grid = reinterpret_cast<uint16_t*>(base + p.off_occ);    // COLL 0 / 2
  uint32_t* s_plane1 = reinterpret_cast<uint32_t*>(base + p.off_occ);  // COLL 1: "claimed"
  uint32_t* s_plane2 = s_plane1 + p.plane_words;                       //         "claimed twice"
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(base + p.off_stage);
  uint32_t* const s_link0 = reinterpret_cast<uint32_t*>(base + p.off_link);
  uint32_t* const s_link1 = reinterpret_cast<uint32_t*>(base + p.off_npos);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(base + p.off_misc);
  int* s_cnt = reinterpret_cast<int*>(base + p.off_misc + 8);  // [0] on_goal count [1] was_on_goal count

  const int A = p.A, PW = p.PW, WPR = p.WPR;
  const int ONTGT = p.on_target;
  const bool narrow = p.narrow != 0;
  const long long ia = (long long)n * A;
  long long* dbg = (p.debug != nullptr && tid == 0) ? p.debug + (long long)n * 16 : nullptr;

  pdl_trigger();
  PGM_STAMP(0);
  PGM_STAMP_NS(9);
#ifdef PGM_TIMELINE
  if (dbg != nullptr) {  // where this team runs: SM and hardware warp slot (tools/phase_timeline.py)
    uint32_t smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    dbg[12] = (long long)((smid << 16) | warpid);
  }
#endif
  // ---- prologue (independent of the previous launch): the obstacle bitmap comes by bulk copy (TMA engine), the
  // team zeroes its agent bitmap and sets every cell of the grid to "nobody" with its own 16-byte stores.  (Filling
  // them by a second bulk copy from a constant template saves ~70 instructions per warp and was measured SLOWER on
  // small single-step launches, 6.4 -> 7.0 us for 2048 instances at r=3: a bulk copy has ~1 us of latency, and a
  // launch that cannot become resident under its predecessor sees all of it.)
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)p.obst_stride * 4u;
    mbar_expect_tx(s_bar, bytes);
    bulk_g2s(base + p.off_obst, p.obst + (long long)n * p.obst_stride, bytes, s_bar);
  }
  const int bm_vec = (p.PH * WPR + 1 + 3) >> 2;  // 16-byte vectors of one bitmap
  auto zero_bitmap = [&](uint32_t* bm) {
    uint4* b4 = reinterpret_cast<uint4*>(bm);
    for (int w = tid; w < bm_vec; w += TEAM) b4[w] = make_uint4(0u, 0u, 0u, 0u);
  };
  zero_bitmap(s_abits0);
  if (COLL != 1) {
    const int gvec = (p.PH * PW * 2 + 15) >> 4;
    uint4* g4 = reinterpret_cast<uint4*>(s_grid);
    for (int w = tid; w < gvec; w += TEAM) g4[w] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  PGM_STAMP(1);
  pdl_wait();
  PGM_STAMP(2);
  PGM_STAMP_NS(10);
  // ---- mutable state of this instance into registers; the first step's actions travel at the same time
  const int isz = p.act_itemsize;
  const uint8_t* act_ptr = p.actions + (ia + tid) * isz;  // this thread's first agent, step 0
  auto issue_actions = [&](const uint8_t* ptr, uint32_t (&raw)[APT], const bool (&pres)[APT]) {
    if (isz == 1) {
#pragma unroll
      for (int q = 0; q < APT; ++q) raw[q] = pres[q] ? (uint32_t)ptr[q * TEAM] : 0u;
    } else {
#pragma unroll
      for (int q = 0; q < APT; ++q) raw[q] = pres[q] ? load_action(ptr, (long long)q * TEAM, isz) : 0u;
    }
  };
  uint32_t pos[APT], tgt[APT], act[APT], raw_act[APT];
  bool present[APT], active[APT];
  uint2 w_in[APT];
#pragma unroll
  for (int q = 0; q < APT; ++q) {
    const int a = q * TEAM + tid;
    present[q] = a < A;
    w_in[q] = make_uint2(0u, 0u);
    if (present[q]) w_in[q] = p.state[ia + a];
  }
  issue_actions(act_ptr, raw_act, present);
  int step_idx = p.elapsed[n];
  int m_acc0, m_acc1, m_acc2;
  {
    const int4 m = *reinterpret_cast<const int4*>(p.metric_acc + 4 * (long long)n);
    m_acc0 = m.x;
    m_acc1 = m.y;
    m_acc2 = m.z;
  }
#pragma unroll
  for (int q = 0; q < APT; ++q) {
    pos[q] = st_pos(w_in[q].x);
    tgt[q] = w_in[q].y;
    active[q] = present[q] && st_active(w_in[q].x) != 0u;
    act[q] = 0u;
  }
  int cur = 0;  // block_both: abits<cur> = occupancy before the move, abits<cur ^ 1> = after it
  team_sync<TEAM>(bar_id);  // the mbarrier was initialised by thread 0 of the team
  mbar_wait(s_bar, 0);
  if (COLL == 1) {
#pragma unroll
    for (int q = 0; q < APT; ++q)
      if (active[q]) {
        const int x = pos[q] & 0xFFFF, y = pos[q] >> 16;
        atomicOr(&s_abits0[x * WPR + (y >> 5)], 1u << (y & 31));
      }
  }

  const int num_steps = p.num_steps;
  int obs_slot = 0;
  uint8_t* obs_cur = p.obs;
  // per-thread output cursors: agent q of this thread sits q * TEAM elements further (a compile-time offset)
  float* rew_ptr = p.rewards + (ia + tid);
  uint8_t* term_ptr = p.terminated + (ia + tid);
  uint8_t* trunc_ptr = p.truncated + (ia + tid);
#pragma unroll 1
  for (int k = 0; k < num_steps; ++k) {
    // ---- actions of step k (loaded one step ahead), then the loads of step k + 1 go out
    {
      bool bad = false;
#pragma unroll
      for (int q = 0; q < APT; ++q) {
        uint32_t v = raw_act[q];
        if (v > 4u) {
          bad = true;
          v = 0u;
        }
        act[q] = v;
      }
      // (one vote and at most one atomic per warp: a plain `if (bad) atomicOr` costs ~17 instructions per warp and
      // step for its aggregation code even when no action is out of range)
      if (__ballot_sync(0xffffffffu, bad) != 0u && lane == 0) atomicOr(p.err_flag, 1);
      if (k + 1 < num_steps) {
        act_ptr += p.act_step_stride;
        issue_actions(act_ptr, raw_act, present);
      }
    }
    uint8_t* obs_k = obs_cur;  // this step's slot of the observation ring (nullptr: no observations wanted)
    if (p.obs != nullptr) {
      obs_cur += p.obs_slot_stride;
      if (++obs_slot == p.obs_ring) {
        obs_slot = 0;
        obs_cur = p.obs;
      }
    }
    if (tid == 0) {
      s_cnt[0] = 0;
      s_cnt[1] = 0;
    }

    bool moved[APT];
    if (COLL != 1) {
      // ---- phase 1: every active agent publishes (index, action) on its cell
      int cell[APT];
#pragma unroll
      for (int q = 0; q < APT; ++q) {
        cell[q] = (int)(pos[q] & 0xFFFF) * PW + (int)(pos[q] >> 16);
        if (active[q]) s_grid[cell[q]] = (uint16_t)((uint32_t)(q * TEAM + tid) | (act[q] << G_ACT));
      }
      team_sync<TEAM>(bar_id);
      // (every warp of the team is past its previous observation phase now: the agent bitmap may be cleared; the
      // barriers of the resolution phase order this before the atomicOr's of phase 3)
      if (k > 0) zero_bitmap(s_abits0);  // (first step: zeroed in the prologue)
      PGM_STAMP(3);
      if (COLL == 2) {
        // soft, pass A: moves into obstacles and edge swaps become 'stay' (judged on the raw actions)
        uint32_t eff[APT];
#pragma unroll
        for (int q = 0; q < APT; ++q) {
          eff[q] = act[q];
          if (active[q] && act[q] != 0u) {
            const int tx = (int)(pos[q] & 0xFFFF) + move_dx(act[q]), ty = (int)(pos[q] >> 16) + move_dy(act[q]);
            if (bit_at(s_obst, WPR, tx, ty)) {
              eff[q] = 0u;
            } else {
              const uint32_t e = s_grid[tx * PW + ty];
              if (e != G_NONE && (e >> G_ACT) == (uint32_t)opposite(act[q])) eff[q] = 0u;
            }
          }
        }
        team_sync<TEAM>(bar_id);
#pragma unroll
        for (int q = 0; q < APT; ++q) {
          if (active[q] && eff[q] != act[q]) s_grid[cell[q]] = (uint16_t)((uint32_t)(q * TEAM + tid) | (eff[q] << G_ACT));
          act[q] = eff[q];
        }
        team_sync<TEAM>(bar_id);
      }
      // ---- phase 2: move resolution (closed forms of upstream envs.py :: move_agents, see pgm_kernels.cuh)
      uint32_t link[APT];
      bool pend = false;
#pragma unroll
      for (int q = 0; q < APT; ++q) {
        const int a = q * TEAM + tid;
        uint32_t l = ST_FAIL;
        if (active[q] && act[q] != 0u) {
          const int dx = move_dx(act[q]), dy = move_dy(act[q]);
          const int tx = (int)(pos[q] & 0xFFFF) + dx, ty = (int)(pos[q] >> 16) + dy;
          if (!bit_at(s_obst, WPR, tx, ty)) {
            const int t = cell[q] + dx * PW + dy;
            const uint32_t e = s_grid[t];
            bool ok = true;
            int lo = -1;
            if (e != G_NONE) {
              const int j = (int)(e & ((1u << G_ACT) - 1u));
              if (COLL == 0) {
                // priority: the occupant must have a lower index and leave
                if (j > a || (e >> G_ACT) == 0u) ok = false;
                else lo = j;
              } else {
                // soft: a stayer rejects everybody
                if ((e >> G_ACT) == 0u) ok = false;
              }
            }
            // a claimant of the cell with an index in (lo, a) beats this agent: the occupants of the four
            // neighbours of t whose action points into t (this agent itself sits on one of them: index a, excluded)
            const uint32_t span = (uint32_t)(a - lo - 1);
            const uint32_t e1 = s_grid[t + PW], e2 = s_grid[t - PW], e3 = s_grid[t + 1], e4 = s_grid[t - 1];
            bool other = (e1 - ((1u << G_ACT) + (uint32_t)(lo + 1))) < span;   // from below, moving up    (action 1)
            other |= (e2 - ((2u << G_ACT) + (uint32_t)(lo + 1))) < span;       // from above, moving down  (action 2)
            other |= (e3 - ((3u << G_ACT) + (uint32_t)(lo + 1))) < span;       // from the right, moving left (3)
            other |= (e4 - ((4u << G_ACT) + (uint32_t)(lo + 1))) < span;       // from the left, moving right (4)
            if (ok && !other) l = (e == G_NONE) ? ST_OK : (ST_PEND | ((e & ((1u << G_ACT) - 1u)) << 2));
          }
        }
        link[q] = l;
        if (present[q]) s_link0[a] = l;
        pend |= ((l & 3u) == ST_PEND);
      }
      // pointer jumping along occupant chains (double buffered); what is still pending after max_rounds is a
      // rotation cycle (soft only) and succeeds
      int lc = 0, rounds = 0;
      while (team_any<TEAM>(bar_id, pend)) {
        if (++rounds > p.max_rounds) break;
        pend = false;
#pragma unroll
        for (int q = 0; q < APT; ++q) {
          uint32_t l = link[q];
          if ((l & 3u) == ST_PEND) {
            const uint32_t lp = (lc ? s_link1 : s_link0)[l >> 2];
            l = ((lp & 3u) == ST_PEND) ? (ST_PEND | (lp & ~3u)) : (lp & 3u);
            pend |= ((l & 3u) == ST_PEND);
          }
          link[q] = l;
          if (present[q]) (lc ? s_link0 : s_link1)[q * TEAM + tid] = l;
        }
        lc ^= 1;
      }
      PGM_STAMP(4);
      // (every thread is past its grid reads: the team_any barrier above)
#pragma unroll
      for (int q = 0; q < APT; ++q) moved[q] = (link[q] & 3u) != ST_FAIL;
    } else {
      // ---- block_both: move iff the target is free of obstacle, of any pre-move active agent, and claimed once
      const uint32_t* pre = cur ? s_abits1 : s_abits0;
      zero_bitmap(cur ? s_abits0 : s_abits1);
      zero_bitmap(s_plane1);
      zero_bitmap(s_plane2);
      team_sync<TEAM>(bar_id);
      PGM_STAMP(3);
      int tw[APT];
      uint32_t tbit[APT];
#pragma unroll
      for (int q = 0; q < APT; ++q) {
        moved[q] = false;
        tw[q] = 0;
        tbit[q] = 0u;
        if (active[q] && act[q] != 0u) {
          const int tx = (int)(pos[q] & 0xFFFF) + move_dx(act[q]), ty = (int)(pos[q] >> 16) + move_dy(act[q]);
          const int w = tx * WPR + (ty >> 5);
          const uint32_t b = 1u << (ty & 31);
          if (((s_obst[w] | pre[w]) & b) == 0u) {
            moved[q] = true;  // candidate
            tw[q] = w;
            tbit[q] = b;
            const uint32_t old = atomicOr(&s_plane1[w], b);
            if (old & b) atomicOr(&s_plane2[w], b);
          }
        }
      }
      team_sync<TEAM>(bar_id);
#pragma unroll
      for (int q = 0; q < APT; ++q)
        if (moved[q] && (s_plane2[tw[q]] & tbit[q]) != 0u) moved[q] = false;
      cur ^= 1;
      PGM_STAMP(4);
    }

    // ---- phase 3: apply moves, on_target bookkeeping, time limit (upstream Pogema.step and friends)
    int c_on = 0, c_was = 0;
    uint32_t npos[APT];
    bool on[APT], was[APT];
#pragma unroll
    for (int q = 0; q < APT; ++q) {
      uint32_t pp = pos[q];
      if (moved[q]) {
        const int tx = (int)(pp & 0xFFFF) + move_dx(act[q]), ty = (int)(pp >> 16) + move_dy(act[q]);
        pp = (uint32_t)tx | ((uint32_t)ty << 16);
      }
      npos[q] = pp;
      on[q] = present[q] && pp == tgt[q];
      was[q] = on[q] && active[q];
      c_on += on[q] ? 1 : 0;
      c_was += was[q] ? 1 : 0;
    }
    if (APT == 1) {
      c_on = __popc(__ballot_sync(0xffffffffu, on[0]));
      c_was = __popc(__ballot_sync(0xffffffffu, was[0]));
    } else {
      c_on = __reduce_add_sync(0xffffffffu, c_on);
      c_was = __reduce_add_sync(0xffffffffu, c_was);
    }
    if (TEAM > 32) {
      if (lane == 0) {
        atomicAdd(&s_cnt[0], c_on);
        atomicAdd(&s_cnt[1], c_was);
      }
      team_sync<TEAM>(bar_id);
      c_on = s_cnt[0];
      c_was = s_cnt[1];
    }
    const bool trunc = (step_idx + 1 >= p.max_steps);
    const bool solved = (c_was == A);
    const bool all_term = (ONTGT == 2) ? false : (c_on == A);
    const bool done = trunc || all_term;
    const bool do_reset = done && p.auto_reset == 1;
    const bool reseed = done && p.auto_reset == 2;  // new task from a new seed: built after this launch
    uint32_t* post = (COLL == 1 && cur) ? s_abits1 : s_abits0;
#pragma unroll
    for (int q = 0; q < APT; ++q) {
      const int a = q * TEAM + tid;
      if (!present[q]) continue;
      float rew;
      uint8_t term;
      bool nact = active[q];
      uint32_t tt = tgt[q];
      if (ONTGT == 0) {  // finish: reward once, agent disappears
        rew = was[q] ? 1.0f : 0.0f;
        term = on[q] ? 1 : 0;
        nact = active[q] && !on[q];
      } else if (ONTGT == 1) {  // nothing (cooperative finish)
        rew = solved ? 1.0f : 0.0f;
        term = solved ? 1 : 0;
      } else {  // restart (lifelong): new target from the agent's own generator
        rew = was[q] ? 1.0f : 0.0f;
        term = 0;
        if (on[q] && !do_reset) {
          Pcg64 g = p.rng[ia + a];
          const uint32_t kk = pcg64_bounded32(g, (uint32_t)(p.comp_size[ia + a] - 1));
          tt = p.cells[(long long)n * p.cells_stride + p.comp_start[ia + a] + kk];
          p.rng[ia + a] = g;
        }
      }
      // Evict-first stores (st.global.cs), like the observations.  As ordinary lines the rewards and flags of a
      // rollout stay in L2 behind the evict-first observation lines until they fill it: same box, 16 / 64 steps per
      // launch, configs[1] 15.2 / 15.9 -> 14.8 / 14.8 us per step, configs[2] 16.2 / 16.7 -> 16.1 / 15.9, configs[3]
      // 35.7 -> 34.0 (0.88 -> 0.925 of the roofline); one launch per step unchanged.
      __stcs(rew_ptr + q * TEAM, rew);
      __stcs(term_ptr + q * TEAM, term);
      __stcs(trunc_ptr + q * TEAM, (uint8_t)(trunc ? 1 : 0));
      p.was_on_goal[ia + a] = was[q] ? 1 : 0;
      uint32_t pp = npos[q];
      if (ONTGT == 1) solve_time_update(p.solve + 2 * (ia + a), was[q], pp != pos[q], pos[q] == tt, done, m_acc2);
      if (do_reset) {
        const uint2 w = p.state0[ia + a];
        pp = st_pos(w.x);
        tt = w.y;
        nact = true;
        if (ONTGT == 2) p.rng[ia + a] = p.rng0[ia + a];
      }
      if (COLL != 1) {
        // the grid keeps only what stays true: an agent that left its cell, disappeared or was reset clears it
        if (active[q] && (pp != pos[q] || !nact || do_reset))
          s_grid[(int)(pos[q] & 0xFFFF) * PW + (int)(pos[q] >> 16)] = (uint16_t)G_NONE;
      }
      pos[q] = pp;
      tgt[q] = tt;
      active[q] = nact;
      p.state[ia + a] = make_uint2(pp | ((nact ? 1u : 0u) << 15), tt);
      // post-move agent bitmap (observation channel 1; next step's pre-move occupancy for block_both)
      if (nact) atomicOr(&post[(int)(pp & 0xFFFF) * WPR + (int)(pp >> 21)], 1u << ((pp >> 16) & 31u));
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
    rew_ptr += p.out_step_stride;
    term_ptr += p.out_step_stride;
    trunc_ptr += p.out_step_stride;
    team_sync<TEAM>(bar_id);  // the agent bitmap is complete (and every thread is past the claim planes)
    PGM_STAMP(5);

    // ---- phase 4: observations.  A batch is one agent per thread; a WARP's 32 agents are a word-aligned piece of the
    // batch stream (32 * bits_per_agent bits = bits_per_agent words), so every warp assembles, expands and stores its
    // own piece without waiting for the other warps of the team: no team barrier in this phase
    if (obs_k != nullptr) {
      const uint32_t sbpa = (uint32_t)p.stage_bpa;
      uint8_t* obs_n = obs_k + (long long)n * p.obs_inst_stride;
      const int warp = tid >> 5;
      uint32_t* wstage = s_stage + (uint32_t)warp * sbpa;  // this warp's sbpa words
#pragma unroll
      for (int q = 0; q < APT; ++q) {
        const int wfirst = q * TEAM + (warp << 5);          // first agent of this warp's piece
        if (wfirst >= A) break;                             // warp-uniform
        const int cnt = min(32, A - wfirst);
        if (q > 0) __syncwarp();                            // the piece of the previous batch has been read
        {
          uint32_t acc[NW];
          if (present[q]) {
            if (narrow) fast_agent_bits<D, true>(acc, s_obst, post, WPR, pos[q], tgt[q]);
            else fast_agent_bits<D, false>(acc, s_obst, post, WPR, pos[q], tgt[q]);
          } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) acc[i] = 0u;
          }
          fast_store_stream<NW>(wstage, acc, (uint32_t)lane * sbpa, sbpa, present[q], lane + 1 < cnt, lane);
        }
        __syncwarp();
        if (q == 0) {
          PGM_STAMP(6);
          PGM_STAMP_NS(13);
        }
        if (p.obs_format & 1) {
          // 1: bits, 32-bit words per agent;  3: the raw stream of the batch (packed host transport): batch q starts at
          // the word its first agent has in format 1, this warp's piece `warp * bits_per_agent` words further
          constexpr int WPA = (BPA + 31) >> 5;
          uint32_t* out = reinterpret_cast<uint32_t*>(obs_n) + (long long)(q * TEAM) * WPA + (uint32_t)warp * sbpa;
          const int nw = (cnt * (int)sbpa + 31) >> 5;
          for (int w = lane; w < nw; w += 32) __stcs(out + w, wstage[w]);
        } else {
          fast_expand_u8<32>(wstage, obs_n + (long long)wfirst * BPA, cnt * BPA, lane);
        }
        if (q == 0) PGM_STAMP(7);
      }
    }
    PGM_STAMP(8);
    PGM_STAMP_NS(11);
    // block_both: the next step zeroes the claim planes, which lie under the stream pieces
    if (COLL == 1 && k + 1 < num_steps) team_sync<TEAM>(bar_id);
  }
}

}  // namespace pgm
