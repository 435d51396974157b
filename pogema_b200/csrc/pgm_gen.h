// pgm_gen.h - native task generator (host side): the product's implementation of
// upstream generator.py + grid.py :: Grid.__init__ / add_artificial_border /
// GridLifeLong.__init__ + envs.py :: PogemaLifeLong._initialize_grid.
#pragma once
#include <stdint.h>

#include <vector>

#include "pgm_rng.h"

namespace pgm {

struct GenParams {
  int H = 0, W = 0;          // unpadded map size
  int A = 0;                 // agents
  int r = 0;                 // obs radius
  double density = 0.0;      // GridConfig.density
  bool lifelong = false;     // on_target == restart
  const uint8_t* map = nullptr;  // optional fixed obstacle map [H][W]
  int num_retries = 10;      // upstream Grid.__init__(num_retries=10)
};

// One generated instance, in PADDED coordinates.
struct GenInstance {
  int PH = 0, PW = 0, WPR = 0;
  std::vector<uint32_t> obst_bits;  // [PH][WPR] bit y&31 of word (x*WPR + y/32)
  std::vector<uint32_t> pos, tgt;   // packed x | y << 16, padded
  // lifelong only
  std::vector<Pcg64> rng;               // per agent generator
  std::vector<int32_t> comp_start;      // per agent: offset of its component in `cells`
  std::vector<int32_t> comp_size;       // per agent: size of its component
  std::vector<uint32_t> cells;          // packed padded coords, grouped by component, row-major inside
};

// Constants of numpy's random_binomial(p, n=1) inversion draw, computed with the host libm exactly as
// numpy does: a cell is an obstacle iff (U > qn) [xor flip]; px1 is the next inversion threshold.
void binomial1_constants(double p, int* zero, int* flip, double* qn, double* px1);

// Returns 0, or -1 for the upstream OverflowError ("Can't create task").
int generate_instance(const GenParams& p, uint64_t seed, GenInstance& out);

// Explicit task (GridConfig.map + agents_xy + targets_xy).  obstacles [H][W],
// agents_xy/targets_xy [A][2] unpadded.  Obstacles under starts/finishes are
// cleared (upstream Grid.__init__).
int explicit_instance(const GenParams& p, uint64_t seed, const uint8_t* obstacles,
                      const int32_t* agents_xy, const int32_t* targets_xy, GenInstance& out);

}  // namespace pgm
