// pgm_gen.cpp - native task generator (host side).  See pgm_gen.h.
//
// Restates, with the numpy streams reproduced bit for bit (pgm_rng.h):
//   upstream generator.py :: generate_obstacles            -> draw_obstacles
//   upstream generator.py :: bfs                           -> label_components
//   upstream generator.py :: generate_positions_and_targets_fast / placing -> place_agents
//   upstream grid.py      :: Grid.__init__ (retry loop, OverflowError)     -> generate_instance
//   upstream grid.py      :: Grid.add_artificial_border    -> pad_and_pack
//   upstream generator.py :: get_components, grid.py :: GridLifeLong.__init__,
//   upstream envs.py      :: PogemaLifeLong._initialize_grid -> build_lifelong
#include "pgm_gen.h"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace pgm {
namespace {

// numpy random_binomial(p, n=1): one inversion draw per cell
// (numpy/random/src/distributions/distributions.c :: random_binomial_inversion).
struct Binomial1 {
  bool zero = false, flip = false;
  double pe = 0, q = 0, qn = 0;
  long bound = 1;
  explicit Binomial1(double p) {
    if (p == 0.0) {
      zero = true;
      return;
    }
    if (p <= 0.5) {
      pe = p;
    } else {
      pe = 1.0 - p;
      flip = true;
    }
    q = 1.0 - pe;
    qn = exp(1.0 * log(q));
    double np = 1.0 * pe;
    double b = np + 10.0 * sqrt(np * q + 1);
    bound = (long)(b < 1.0 ? b : 1.0);
  }
  int draw(Pcg64& g) const {
    if (zero) return 0;
    long X = 0;
    double px = qn;
    double U = pcg64_next_double(g);
    while (U > px) {
      X++;
      if (X > bound) {
        X = 0;
        px = qn;
        U = pcg64_next_double(g);
      } else {
        U -= px;
        px = ((1 - X + 1) * pe * px) / (X * q);
      }
    }
    return (int)(flip ? 1 - X : X);
  }
};

void draw_obstacles(const GenParams& p, Pcg64& g, std::vector<uint8_t>& obst) {
  Binomial1 b(p.density);
  const int n = p.H * p.W;
  obst.resize(n);
  for (int i = 0; i < n; ++i) obst[i] = (uint8_t)b.draw(g);
}

// Row-major scan, labels from 2, 4-connected.  Returns component sizes by label.
void label_components(const uint8_t* obst, int H, int W, std::vector<int32_t>& label,
                      std::vector<int32_t>& comp_size) {
  label.assign((size_t)H * W, 0);
  for (int i = 0; i < H * W; ++i) label[i] = obst[i] ? 1 : 0;
  comp_size.assign(2, 0);
  std::vector<int32_t> queue;
  queue.reserve((size_t)H * W);
  int32_t cur = 2;
  for (int s = 0; s < H * W; ++s) {
    if (label[s] != 0) continue;
    label[s] = cur;
    comp_size.push_back(1);
    queue.clear();
    queue.push_back(s);
    size_t head = 0;
    while (head < queue.size()) {
      int c = queue[head++];
      int cx = c / W, cy = c % W;
      // MOVES order (-1,0) (1,0) (0,-1) (0,1); order does not change the labelling
      if (cx > 0 && label[c - W] == 0) { label[c - W] = cur; comp_size[cur]++; queue.push_back(c - W); }
      if (cx + 1 < H && label[c + W] == 0) { label[c + W] = cur; comp_size[cur]++; queue.push_back(c + W); }
      if (cy > 0 && label[c - 1] == 0) { label[c - 1] = cur; comp_size[cur]++; queue.push_back(c - 1); }
      if (cy + 1 < W && label[c + 1] == 0) { label[c + 1] = cur; comp_size[cur]++; queue.push_back(c + 1); }
    }
    cur++;
  }
}

// generate_positions_and_targets_fast + placing.  starts may come back shorter
// than A (upstream then retries / raises); finishes always has A slots.
void place_agents(const GenParams& p, const std::vector<uint8_t>& obst, uint64_t seed,
                  std::vector<int32_t>& starts, std::vector<int32_t>& finishes) {
  std::vector<int32_t> label, comp;
  label_components(obst.data(), p.H, p.W, label, comp);
  std::vector<int32_t> order;
  order.reserve((size_t)p.H * p.W);
  for (int i = 0; i < p.H * p.W; ++i)
    if (label[i] >= 2) order.push_back(i);
  // Generator.shuffle on a Python list: Fisher-Yates from the back, random_interval(i)
  Pcg64 g;
  pcg64_seed(g, seed);
  for (int64_t i = (int64_t)order.size() - 1; i > 0; --i) {
    int64_t j = (int64_t)pcg64_interval(g, (uint64_t)i);
    std::swap(order[i], order[j]);
  }
  std::vector<std::vector<int32_t>> requests(comp.size());
  int done_requests = 0;
  starts.clear();
  finishes.assign(p.A, -1);
  for (int32_t cell : order) {
    if (label[cell] < 2) continue;
    int id = label[cell];
    label[cell] = 0;
    if (!requests[id].empty()) {
      int tt = requests[id].back();
      requests[id].pop_back();
      finishes[tt] = cell;
      done_requests++;
      continue;
    }
    if ((int)starts.size() >= p.A) {
      if (done_requests >= p.A) break;
      continue;
    }
    if (comp[id] >= 2) {
      comp[id] -= 2;
      requests[id].push_back((int)starts.size());
      starts.push_back(cell);
    }
  }
}

inline uint32_t pack_xy(int x, int y) { return (uint32_t)x | ((uint32_t)y << 16); }

void pad_and_pack(const GenParams& p, const std::vector<uint8_t>& obst, GenInstance& out) {
  const int r = p.r;
  out.PH = p.H + 2 * r;
  out.PW = p.W + 2 * r;
  out.WPR = (out.PW + 31) / 32;
  out.obst_bits.assign((size_t)out.PH * out.WPR, 0u);
  auto set = [&](int x, int y) { out.obst_bits[(size_t)x * out.WPR + (y >> 5)] |= 1u << (y & 31); };
  // empty_outside=True: zeros outside, then a one cell ring at r-1 / P-r
  for (int y = r - 1; y <= out.PW - r; ++y) {
    set(r - 1, y);
    set(out.PH - r, y);
  }
  for (int x = r - 1; x <= out.PH - r; ++x) {
    set(x, r - 1);
    set(x, out.PW - r);
  }
  for (int x = 0; x < p.H; ++x)
    for (int y = 0; y < p.W; ++y)
      if (obst[(size_t)x * p.W + y]) set(x + r, y + r);
}

// Components over the PADDED grid (get_components), per-agent generators, and
// the goal fix-up of GridLifeLong.__init__ (uses Grid.rnd).
void build_lifelong(const GenParams& p, uint64_t seed, Pcg64& grid_rnd, GenInstance& out) {
  const int PH = out.PH, PW = out.PW;
  std::vector<uint8_t> padded((size_t)PH * PW);
  for (int x = 0; x < PH; ++x)
    for (int y = 0; y < PW; ++y)
      padded[(size_t)x * PW + y] = (out.obst_bits[(size_t)x * out.WPR + (y >> 5)] >> (y & 31)) & 1u;
  std::vector<int32_t> label, comp;
  label_components(padded.data(), PH, PW, label, comp);
  // cells of every component that hosts an agent, row-major inside a component
  std::vector<int32_t> comp_offset(comp.size(), -1);
  out.comp_start.assign(p.A, 0);
  out.comp_size.assign(p.A, 0);
  std::vector<int32_t> wanted;
  for (int a = 0; a < p.A; ++a) {
    int x = out.pos[a] & 0xFFFF, y = out.pos[a] >> 16;
    int id = label[(size_t)x * PW + y];
    if (comp_offset[id] < 0) {
      comp_offset[id] = 0;
      wanted.push_back(id);
    }
  }
  int32_t total = 0;
  for (int id : wanted) {
    comp_offset[id] = total;
    total += comp[id];
  }
  out.cells.assign(total, 0u);
  std::vector<int32_t> fill(comp.size(), 0);
  for (int c = 0; c < PH * PW; ++c) {
    int id = label[c];
    if (id >= 2 && comp_offset[id] >= 0) out.cells[comp_offset[id] + fill[id]++] = pack_xy(c / PW, c % PW);
  }
  for (int a = 0; a < p.A; ++a) {
    int x = out.pos[a] & 0xFFFF, y = out.pos[a] >> 16;
    int id = label[(size_t)x * PW + y];
    out.comp_start[a] = comp_offset[id];
    out.comp_size[a] = comp[id];
    int tx = out.tgt[a] & 0xFFFF, ty = out.tgt[a] >> 16;
    if (label[(size_t)tx * PW + ty] != id) {
      // "The start point and the goal are in different components. The goal is changed."
      uint32_t k = pcg64_bounded32(grid_rnd, (uint32_t)(comp[id] - 1));
      out.tgt[a] = out.cells[comp_offset[id] + k];
    }
  }
  // envs.py :: PogemaLifeLong._initialize_grid
  Pcg64 main_rng;
  pcg64_seed(main_rng, seed);
  out.rng.resize(p.A);
  for (int a = 0; a < p.A; ++a) {
    uint32_t s = pcg64_bounded32(main_rng, 0x7FFFFFFFu - 1u);  // integers(iinfo(int32).max)
    pcg64_seed(out.rng[a], (uint64_t)s);
  }
}

void finish_instance(const GenParams& p, uint64_t seed, const std::vector<uint8_t>& obst,
                     const std::vector<int32_t>& starts, const std::vector<int32_t>& finishes,
                     Pcg64& grid_rnd, GenInstance& out) {
  pad_and_pack(p, obst, out);
  out.pos.resize(p.A);
  out.tgt.resize(p.A);
  for (int a = 0; a < p.A; ++a) {
    out.pos[a] = pack_xy(starts[a] / p.W + p.r, starts[a] % p.W + p.r);
    out.tgt[a] = pack_xy(finishes[a] / p.W + p.r, finishes[a] % p.W + p.r);
  }
  out.rng.clear();
  out.comp_start.clear();
  out.comp_size.clear();
  out.cells.clear();
  if (p.lifelong) build_lifelong(p, seed, grid_rnd, out);
}

}  // namespace

void binomial1_constants(double p, int* zero, int* flip, double* qn, double* px1) {
  Binomial1 b(p);
  *zero = b.zero ? 1 : 0;
  *flip = b.flip ? 1 : 0;
  *qn = b.qn;
  *px1 = b.zero ? 0.0 : ((1 - 1 + 1) * b.pe * b.qn) / (1 * b.q);
}

int generate_instance(const GenParams& p, uint64_t seed, GenInstance& out) {
  Pcg64 grid_rnd;  // Grid.rnd = default_rng(seed)
  pcg64_seed(grid_rnd, seed);
  std::vector<uint8_t> obst;
  if (p.map) {
    obst.assign(p.map, p.map + (size_t)p.H * p.W);
  } else {
    Pcg64 g;  // generate_obstacles: a fresh default_rng(seed)
    pcg64_seed(g, seed);
    draw_obstacles(p, g, obst);
  }
  std::vector<int32_t> starts, finishes;
  place_agents(p, obst, seed, starts, finishes);
  if ((int)starts.size() != p.A) {
    for (int attempt = 0; attempt < p.num_retries; ++attempt) {
      if ((int)starts.size() == p.A) break;
      if (!p.map) {
        draw_obstacles(p, grid_rnd, obst);
        place_agents(p, obst, seed, starts, finishes);
      }
    }
  }
  if (starts.empty() || (int)starts.size() != p.A) return -1;
  for (int a = 0; a < p.A; ++a)
    if (finishes[a] < 0) return -1;
  finish_instance(p, seed, obst, starts, finishes, grid_rnd, out);
  return 0;
}

int explicit_instance(const GenParams& p, uint64_t seed, const uint8_t* obstacles,
                      const int32_t* agents_xy, const int32_t* targets_xy, GenInstance& out) {
  Pcg64 grid_rnd;
  pcg64_seed(grid_rnd, seed);
  std::vector<uint8_t> obst(obstacles, obstacles + (size_t)p.H * p.W);
  std::vector<int32_t> starts(p.A), finishes(p.A);
  for (int a = 0; a < p.A; ++a) {
    int sx = agents_xy[2 * a], sy = agents_xy[2 * a + 1];
    int fx = targets_xy[2 * a], fy = targets_xy[2 * a + 1];
    if (sx < 0 || sx >= p.H || sy < 0 || sy >= p.W || fx < 0 || fx >= p.H || fy < 0 || fy >= p.W) return -2;
    obst[(size_t)sx * p.W + sy] = 0;
    obst[(size_t)fx * p.W + fy] = 0;
    starts[a] = sx * p.W + sy;
    finishes[a] = fx * p.W + fy;
  }
  finish_instance(p, seed, obst, starts, finishes, grid_rnd, out);
  return 0;
}

}  // namespace pgm
