#!/bin/bash
# rollouts cut into launches of <= 16 steps: parity suite + same-box A/B against the previous library
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/z_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/z_pytest.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   %.2f us  frac %.3f' % (d['ms_per_step']*1e3, d['frac_6541']))"; }
for rep in 1 2; do
for lib in pogema_b200/_lib/libpgm_b200_prev.so pogema_b200/_lib/libpgm_b200.so; do
  [ -f $lib ] || continue
  export PGM_B200_LIB=$PWD/$lib
  echo "== $lib"
  echo " c1 many 16 / 32 / 64 / closed"; q --steps 2048 --many 16; q --steps 2048 --many 32; q --steps 2048 --many 64; q --steps 2048 --graph 16
  echo " c2 many 16 / 64 / closed"; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --many 16; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --many 64; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --graph 16
  echo " c3 many 16 / 64"; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --many 16;  q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --many 64
  echo " r3 share many 16 / 64 / closed"; q --n 2048 --r 3 --steps 1024 --many 16; q --n 2048 --r 3 --steps 1024 --many 64; q --n 2048 --r 3 --steps 1024 --graph 16
done
done
