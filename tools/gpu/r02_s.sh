#!/bin/bash
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   %.2f us  frac %.3f' % (d['ms_per_step']*1e3, d['frac_6541']))"; }
for rep in 1 2; do
for lib in ab/libpgm_old.so ab/libpgm_vb.so pogema_b200/_lib/libpgm_b200.so; do
  export PGM_B200_LIB=$PWD/$lib
  echo "== $lib"
  echo " c1 closed"; q --steps 2048 --graph 16
  echo " c2 closed"; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --graph 16
  echo " c3 closed"; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16
done
done
