#!/bin/bash
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['plan'].get('fast',{}); print('   %.2f us  frac %.3f  team %s apt %s tpc %s' % (d['ms_per_step']*1e3, d['frac_6541'], f.get('team_threads'), f.get('agents_per_thread'), f.get('teams_per_cta')))"; }
timeout 900 python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
for rep in 1 2; do
for lib in ab/libpgm_old.so pogema_b200/_lib/libpgm_b200.so; do
  export PGM_B200_LIB=$PWD/$lib
  echo "== $lib"
  echo " c1"; q --steps 1024 --many 16; q --steps 1024 --graph 16
  echo " c2"; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --many 16; q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --graph 16
  echo " c3"; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --many 16; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16
  echo " 16384 r3"; q --n 16384 --r 3 --steps 256 --many 16; q --n 16384 --r 3 --steps 256 --graph 16
  echo " r3 share"; q --n 2048 --r 3 --steps 1024 --many 16; q --n 2048 --r 3 --steps 1024 --graph 16
done
done
