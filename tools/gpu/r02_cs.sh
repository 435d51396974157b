#!/bin/bash
# evict-first stores for rewards / terminated / truncated (fast kernel): same-box A/B against the previous library
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   %.2f us  frac %.3f' % (d['ms_per_step']*1e3, d['frac_6541']))"; }
b() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   driver line: %.2f us frac %.3f steady %.3f closed %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['steady_state']['frac'], d['closed_loop']['roofline_frac']))"; }
for rep in 1 2; do
for lib in pogema_b200/_lib/libpgm_b200_prev.so pogema_b200/_lib/libpgm_b200.so; do
  export PGM_B200_LIB=$PWD/$lib
  echo "== $lib"
  echo " c1 many 16 | many 64 | 16 x 4 sets | closed"; q --steps 2048 --many 16; q --steps 2048 --many 64; q --steps 2048 --many 16 --sets 4; q --steps 2048 --graph 16
  b
  echo " c2 many 16 | many 64 | closed"; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --many 16; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --many 64; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --graph 16
  echo " c3 many 16 | closed"; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --many 16; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16
  echo " r3 share many 16 | closed"; q --n 2048 --r 3 --steps 1024 --many 16; q --n 2048 --r 3 --steps 1024 --graph 16
done
done
