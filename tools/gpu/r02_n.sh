#!/bin/bash
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['plan'].get('fast',{}); print('   %.2f us  frac %.3f  team %s apt %s tpc %s grid %s' % (d['ms_per_step']*1e3, d['frac_6541'], f.get('team_threads'), f.get('agents_per_thread'), f.get('teams_per_cta'), f.get('grid')))"; }
echo "== configs[3] shape, instance count vs SM balance (148 SMs, 4 resident per SM): many / closed"
for n in 444 512 592 740 1024; do echo " n=$n"; q --n $n --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --many 16; q --n $n --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16; done
echo "== configs[1] tpc 4 vs 5 vs 6, three runs each (closed / many)"
for rep in 1 2 3; do for tpc in 4 5 6; do echo " tpc $tpc"; PGM_TPC=$tpc q --steps 2048 --graph 16; PGM_TPC=$tpc q --steps 2048 --many 16; done; done
