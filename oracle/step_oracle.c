/*
 * step_oracle.c - plain C restatement of the POGEMA step path on explicit state.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pogema_oracle.py for the rules: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the timed baseline).
 *
 * PARITY STATUS: parity unpinned against upstream pogema (the mounted reference
 * is README.md:1-5 only).  This file follows oracle/pogema_oracle.py function by
 * function - the same sequential per-agent loops as upstream, with the Python
 * dicts replaced by dense per-cell tables - and tests/test_oracle_c.py pins it
 * against that Python oracle on random scenarios for every mode combination.
 *
 *   move_agents           upstream envs.py :: Pogema.move_agents / _revert_action
 *   grid_move             upstream grid.py :: Grid.move
 *   step_*                upstream envs.py :: Pogema.step / PogemaLifeLong.step / PogemaCoopFinish.step
 *   time limit            upstream wrappers/multi_time_limit.py :: MultiTimeLimit.step
 *   write_obs             upstream envs.py :: _get_agents_obs, grid.py :: get_obstacles_for_agent /
 *                         get_positions / get_square_target
 *   pcg64 / lemire        numpy Generator.choice (third-party dependency of upstream)
 *
 * All coordinates are PADDED (upstream grid.py :: add_artificial_border); x = row.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int MOVES[5][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}};

typedef struct {
  uint64_t state_hi, state_lo, inc_hi, inc_lo;
  uint32_t has_uint32, uinteger;
} orc_pcg64;

typedef struct {
  int32_t PH, PW, A, r;
  int32_t collision; /* 0 priority 1 block_both 2 soft */
  int32_t on_target; /* 0 finish 1 nothing 2 restart */
  int32_t max_episode_steps;
} orc_cfg;

typedef struct {
  /* per instance */
  uint8_t* obstacles; /* [PH*PW] */
  uint8_t* positions; /* [PH*PW] occupancy of ACTIVE agents (upstream Grid.positions) */
  int32_t* pos;       /* [A][2] */
  int32_t* tgt;       /* [A][2] */
  uint8_t* active;    /* [A] */
  int32_t elapsed;
  /* lifelong */
  orc_pcg64* rng;      /* [A] */
  int32_t* comp_start; /* [A] */
  int32_t* comp_size;  /* [A] */
  int32_t* cells;      /* [ncells][2] grouped by component, row-major inside */
} orc_inst;

/* ---- numpy PCG64 (setseq 128 / XSL-RR 64), buffered uint32, bounded Lemire ---- */
static void pcg_step(orc_pcg64* g) {
  unsigned __int128 s = ((unsigned __int128)g->state_hi << 64) | g->state_lo;
  unsigned __int128 inc = ((unsigned __int128)g->inc_hi << 64) | g->inc_lo;
  unsigned __int128 m = ((unsigned __int128)0x2360ed051fc65da4ULL << 64) | 0x4385df649fccf645ULL;
  s = s * m + inc;
  g->state_hi = (uint64_t)(s >> 64);
  g->state_lo = (uint64_t)s;
}
static uint64_t pcg_next64(orc_pcg64* g) {
  pcg_step(g);
  uint64_t x = g->state_hi ^ g->state_lo;
  unsigned rot = (unsigned)(g->state_hi >> 58);
  return (x >> rot) | (x << ((-rot) & 63));
}
static uint32_t pcg_next32(orc_pcg64* g) {
  if (g->has_uint32) {
    g->has_uint32 = 0;
    return g->uinteger;
  }
  uint64_t n = pcg_next64(g);
  g->has_uint32 = 1;
  g->uinteger = (uint32_t)(n >> 32);
  return (uint32_t)n;
}
static uint32_t lemire32(orc_pcg64* g, uint32_t rng) {
  if (rng == 0) return 0;
  uint32_t rng_excl = rng + 1;
  uint64_t m = (uint64_t)pcg_next32(g) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    uint32_t threshold = (UINT32_MAX - rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)pcg_next32(g) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return (uint32_t)(m >> 32);
}

/* ---- upstream grid.py :: Grid.move ---- */
static void grid_move(const orc_cfg* c, orc_inst* s, int i, int action) {
  int x = s->pos[2 * i], y = s->pos[2 * i + 1];
  int dx = MOVES[action][0], dy = MOVES[action][1];
  s->positions[x * c->PW + y] = 0;
  if (s->obstacles[(x + dx) * c->PW + (y + dy)] == 0 && s->positions[(x + dx) * c->PW + (y + dy)] == 0) {
    x += dx;
    y += dy;
  }
  s->pos[2 * i] = x;
  s->pos[2 * i + 1] = y;
  s->positions[x * c->PW + y] = 1;
}

/* per-cell claim lists for the 'soft' system (used_cells of the Python code) */
typedef struct {
  int32_t* head; /* [cells] first claim index or -1 */
  int32_t* next; /* [A] */
  int32_t* cnt;  /* [cells] */
} claims_t;

static void claim_add(claims_t* cl, int cell, int agent) {
  /* append at the tail to keep Python list order */
  cl->next[agent] = -1;
  if (cl->head[cell] < 0) {
    cl->head[cell] = agent;
  } else {
    int k = cl->head[cell];
    while (cl->next[k] >= 0) k = cl->next[k];
    cl->next[k] = agent;
  }
  cl->cnt[cell]++;
}
static void claim_remove(claims_t* cl, int cell, int agent) {
  int k = cl->head[cell], prev = -1;
  while (k >= 0 && k != agent) {
    prev = k;
    k = cl->next[k];
  }
  if (k < 0) return;
  if (prev < 0) cl->head[cell] = cl->next[k];
  else cl->next[prev] = cl->next[k];
  cl->cnt[cell]--;
}

/* upstream envs.py :: Pogema._revert_action */
static void revert_action(const orc_cfg* c, orc_inst* s, claims_t* cl, int agent, int cell, uint8_t* actions) {
  actions[agent] = 0;
  claim_remove(cl, cell, agent);
  int new_cell = s->pos[2 * agent] * c->PW + s->pos[2 * agent + 1];
  /* snapshot of the agents currently claiming new_cell */
  int n_others = cl->cnt[new_cell];
  int32_t* others = (int32_t*)malloc(sizeof(int32_t) * (n_others > 0 ? n_others : 1));
  int m = 0;
  for (int k = cl->head[new_cell]; k >= 0; k = cl->next[k]) others[m++] = k;
  claim_add(cl, new_cell, agent);
  for (int q = 0; q < m; ++q) {
    int other = others[q];
    if (other != agent && actions[other] != 0) revert_action(c, s, cl, other, new_cell, actions);
  }
  free(others);
}

/* upstream envs.py :: Pogema.move_agents */
static void move_agents(const orc_cfg* c, orc_inst* s, const uint8_t* actions_in) {
  const int A = c->A, PW = c->PW, cells = c->PH * c->PW;
  uint8_t* actions = (uint8_t*)malloc(A);
  memcpy(actions, actions_in, A);
  if (c->collision == 0) {
    for (int i = 0; i < A; ++i)
      if (s->active[i]) grid_move(c, s, i, actions[i]);
  } else if (c->collision == 1) {
    /* used_cells: 0 absent, 1 'visited', 2 'blocked' */
    uint8_t* used = (uint8_t*)calloc(cells, 1);
    for (int i = 0; i < A; ++i) {
      if (!s->active[i]) continue;
      int x = s->pos[2 * i], y = s->pos[2 * i + 1];
      int t = (x + MOVES[actions[i]][0]) * PW + (y + MOVES[actions[i]][1]);
      used[t] = used[t] ? 2 : 1;
      used[x * PW + y] = 2;
    }
    /* agents_xy is a snapshot taken before any move */
    int32_t* snap = (int32_t*)malloc(sizeof(int32_t) * 2 * A);
    memcpy(snap, s->pos, sizeof(int32_t) * 2 * A);
    for (int i = 0; i < A; ++i) {
      if (!s->active[i]) continue;
      int x = snap[2 * i], y = snap[2 * i + 1];
      int t = (x + MOVES[actions[i]][0]) * PW + (y + MOVES[actions[i]][1]);
      if (used[t] != 2) grid_move(c, s, i, actions[i]);
    }
    free(snap);
    free(used);
  } else {
    claims_t cl;
    cl.head = (int32_t*)malloc(sizeof(int32_t) * cells);
    cl.cnt = (int32_t*)calloc(cells, sizeof(int32_t));
    cl.next = (int32_t*)malloc(sizeof(int32_t) * A);
    for (int k = 0; k < cells; ++k) cl.head[k] = -1;
    /* (1) obstacles cancel; register claims.  Edge users: an undirected swap is the only way two
       agents can share an edge, so used_edges is restated as "the occupant of my target heads to my cell" */
    for (int i = 0; i < A; ++i) {
      if (!s->active[i]) continue;
      int x = s->pos[2 * i], y = s->pos[2 * i + 1];
      int dx = MOVES[actions[i]][0], dy = MOVES[actions[i]][1];
      if (s->obstacles[(x + dx) * PW + (y + dy)]) {
        actions[i] = 0;
        dx = dy = 0;
      }
      claim_add(&cl, (x + dx) * PW + (y + dy), i);
    }
    /* (2) edge conflicts: both agents of a swap stay (decided on the actions after (1)) */
    {
      int32_t* who = (int32_t*)malloc(sizeof(int32_t) * cells);
      for (int k = 0; k < cells; ++k) who[k] = -1;
      for (int i = 0; i < A; ++i)
        if (s->active[i]) who[s->pos[2 * i] * PW + s->pos[2 * i + 1]] = i;
      uint8_t* swap = (uint8_t*)calloc(A, 1);
      for (int i = 0; i < A; ++i) {
        if (!s->active[i] || actions[i] == 0) continue;
        int x = s->pos[2 * i], y = s->pos[2 * i + 1];
        int tx = x + MOVES[actions[i]][0], ty = y + MOVES[actions[i]][1];
        int j = who[tx * PW + ty];
        if (j >= 0 && actions[j] != 0 && tx + MOVES[actions[j]][0] == x && ty + MOVES[actions[j]][1] == y) swap[i] = 1;
      }
      for (int i = 0; i < A; ++i) {
        if (!swap[i]) continue;
        int x = s->pos[2 * i], y = s->pos[2 * i + 1];
        int tx = x + MOVES[actions[i]][0], ty = y + MOVES[actions[i]][1];
        claim_remove(&cl, tx * PW + ty, i);
        claim_add(&cl, x * PW + y, i);
        actions[i] = 0;
      }
      free(swap);
      free(who);
    }
    /* (3) vertex conflicts, highest index first; cancellations cascade */
    for (int i = A - 1; i >= 0; --i) {
      if (!s->active[i] || actions[i] == 0) continue;
      int x = s->pos[2 * i], y = s->pos[2 * i + 1];
      int t = (x + MOVES[actions[i]][0]) * PW + (y + MOVES[actions[i]][1]);
      if (cl.cnt[t] > 1) revert_action(c, s, &cl, i, t, actions);
    }
    /* (4) apply (move_without_checks) and rebuild the occupancy */
    for (int i = 0; i < A; ++i) {
      if (!s->active[i]) continue;
      s->pos[2 * i] += MOVES[actions[i]][0];
      s->pos[2 * i + 1] += MOVES[actions[i]][1];
    }
    memset(s->positions, 0, cells);
    for (int i = 0; i < A; ++i)
      if (s->active[i]) s->positions[s->pos[2 * i] * PW + s->pos[2 * i + 1]] = 1;
    free(cl.head);
    free(cl.cnt);
    free(cl.next);
  }
  free(actions);
}

static int on_goal(const orc_inst* s, int i) {
  return s->pos[2 * i] == s->tgt[2 * i] && s->pos[2 * i + 1] == s->tgt[2 * i + 1];
}

/* upstream envs.py :: _get_agents_obs for every agent -> uint8 [A][3][D][D] */
static void write_obs(const orc_cfg* c, const orc_inst* s, uint8_t* obs) {
  const int r = c->r, D = 2 * r + 1, PW = c->PW;
  for (int i = 0; i < c->A; ++i) {
    uint8_t* o = obs + (size_t)i * 3 * D * D;
    int x = s->pos[2 * i], y = s->pos[2 * i + 1];
    for (int a = 0; a < D; ++a)
      for (int b = 0; b < D; ++b) {
        int cell = (x - r + a) * PW + (y - r + b);
        o[a * D + b] = s->obstacles[cell];
        o[D * D + a * D + b] = s->positions[cell];
        o[2 * D * D + a * D + b] = 0;
      }
    int dx = x - s->tgt[2 * i], dy = y - s->tgt[2 * i + 1];
    dx = dx >= 0 ? (dx < r ? dx : r) : (dx > -r ? dx : -r);
    dy = dy >= 0 ? (dy < r ? dy : r) : (dy > -r ? dy : -r);
    o[2 * D * D + (r - dx) * D + (r - dy)] = 1;
  }
}

/*
 * One environment step (upstream Pogema.step / PogemaLifeLong.step / PogemaCoopFinish.step wrapped by
 * MultiTimeLimit.step).  rewards float[A], terminated/truncated/was_on_goal uint8[A], obs may be NULL.
 * Returns 1 if the episode ended (all terminated or all truncated).
 */
int orc_step(const orc_cfg* c, orc_inst* s, const uint8_t* actions, float* rewards, uint8_t* terminated,
             uint8_t* truncated, uint8_t* was_on_goal, uint8_t* obs) {
  const int A = c->A;
  move_agents(c, s, actions);
  int all_was = 1, all_term = 1;
  for (int i = 0; i < A; ++i) {
    was_on_goal[i] = (uint8_t)(on_goal(s, i) && s->active[i]);
    if (!was_on_goal[i]) all_was = 0;
  }
  if (c->on_target == 0) {
    for (int i = 0; i < A; ++i) {
      int g = on_goal(s, i);
      rewards[i] = (g && s->active[i]) ? 1.0f : 0.0f;
      terminated[i] = (uint8_t)g;
    }
    for (int i = 0; i < A; ++i) {
      if (on_goal(s, i) && s->active[i]) { /* hide_agent */
        s->active[i] = 0;
        s->positions[s->pos[2 * i] * c->PW + s->pos[2 * i + 1]] = 0;
      }
    }
  } else if (c->on_target == 2) {
    for (int i = 0; i < A; ++i) {
      int g = on_goal(s, i);
      rewards[i] = (g && s->active[i]) ? 1.0f : 0.0f;
      terminated[i] = 0;
      if (g) { /* generate_new_target: one choice() draw over the agent's component */
        uint32_t k = lemire32(&s->rng[i], (uint32_t)(s->comp_size[i] - 1));
        s->tgt[2 * i] = s->cells[2 * (s->comp_start[i] + k)];
        s->tgt[2 * i + 1] = s->cells[2 * (s->comp_start[i] + k) + 1];
      }
    }
  } else {
    for (int i = 0; i < A; ++i) {
      rewards[i] = all_was ? 1.0f : 0.0f;
      terminated[i] = (uint8_t)all_was;
    }
  }
  s->elapsed += 1;
  int trunc = s->elapsed >= c->max_episode_steps;
  for (int i = 0; i < A; ++i) {
    truncated[i] = (uint8_t)trunc;
    if (!terminated[i]) all_term = 0;
  }
  if (obs) write_obs(c, s, obs);
  return trunc || all_term;
}

void orc_observe(const orc_cfg* c, const orc_inst* s, uint8_t* obs) { write_obs(c, s, obs); }

/* Rebuild the occupancy array from positions + active flags (after the caller set the state). */
void orc_rebuild_positions(const orc_cfg* c, orc_inst* s) {
  memset(s->positions, 0, (size_t)c->PH * c->PW);
  for (int i = 0; i < c->A; ++i)
    if (s->active[i]) s->positions[s->pos[2 * i] * c->PW + s->pos[2 * i + 1]] = 1;
}

/*
 * Batched driver used by the tests and by the CPU baseline: N instances laid out contiguously,
 * T steps, actions uint8 [T][N][A].  With auto_reset an instance whose episode ended is restored
 * from (pos0, tgt0, rng0) before its observation is written (AutoResetWrapper semantics).
 * Outputs (any may be NULL): obs [N][A][3][D][D] of the LAST step, rewards_sum double[N][A],
 * rewards/terminated/truncated of the LAST step; final pos/tgt/active are left in the state arrays.  Returns the number of agent-steps done.
 */
long long orc_run(const orc_cfg* c, int N, int T, int auto_reset, uint8_t* obstacles, int32_t* pos, int32_t* tgt,
                  uint8_t* active, int32_t* elapsed, const int32_t* pos0, const int32_t* tgt0, orc_pcg64* rng,
                  const orc_pcg64* rng0, int32_t* comp_start, int32_t* comp_size, int32_t* cells,
                  long long cells_stride, const uint8_t* actions, uint8_t* obs_last, double* rewards_sum,
                  float* rewards_last, uint8_t* term_last, uint8_t* trunc_last) {
  const int A = c->A, cells_n = c->PH * c->PW, D = 2 * c->r + 1;
  uint8_t* positions = (uint8_t*)malloc(cells_n);
  float* rew = (float*)malloc(sizeof(float) * A);
  uint8_t* term = (uint8_t*)malloc(A);
  uint8_t* trunc = (uint8_t*)malloc(A);
  uint8_t* was = (uint8_t*)malloc(A);
  uint8_t* obs_tmp = (uint8_t*)malloc((size_t)A * 3 * D * D);
  for (int n = 0; n < N; ++n) {
    orc_inst s;
    s.obstacles = obstacles + (size_t)n * cells_n;
    s.positions = positions;
    s.pos = pos + (size_t)n * A * 2;
    s.tgt = tgt + (size_t)n * A * 2;
    s.active = active + (size_t)n * A;
    s.elapsed = elapsed[n];
    s.rng = rng ? rng + (size_t)n * A : NULL;
    s.comp_start = comp_start ? comp_start + (size_t)n * A : NULL;
    s.comp_size = comp_size ? comp_size + (size_t)n * A : NULL;
    s.cells = cells ? cells + (size_t)n * cells_stride * 2 : NULL;
    orc_rebuild_positions(c, &s);
    for (int t = 0; t < T; ++t) {
      uint8_t* o = obs_last ? obs_last + (size_t)n * A * 3 * D * D : obs_tmp;
      int done = orc_step(c, &s, actions + ((size_t)t * N + n) * A, rew, term, trunc, was, o);
      if (rewards_sum)
        for (int i = 0; i < A; ++i) rewards_sum[(size_t)n * A + i] += rew[i];
      if (t == T - 1) {
        if (rewards_last) memcpy(rewards_last + (size_t)n * A, rew, sizeof(float) * A);
        if (term_last) memcpy(term_last + (size_t)n * A, term, A);
        if (trunc_last) memcpy(trunc_last + (size_t)n * A, trunc, A);
      }
      if (done && auto_reset) {
        memcpy(s.pos, pos0 + (size_t)n * A * 2, sizeof(int32_t) * 2 * A);
        memcpy(s.tgt, tgt0 + (size_t)n * A * 2, sizeof(int32_t) * 2 * A);
        memset(s.active, 1, A);
        if (s.rng) memcpy(s.rng, rng0 + (size_t)n * A, sizeof(orc_pcg64) * A);
        s.elapsed = 0;
        orc_rebuild_positions(c, &s);
        write_obs(c, &s, o);
      }
    }
    elapsed[n] = s.elapsed;
  }
  free(positions);
  free(rew);
  free(term);
  free(trunc);
  free(was);
  free(obs_tmp);
  return (long long)N * T * A;
}
