#!/bin/bash
mkdir -p gpurun_out
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['plan'].get('fast',{}); print('   %.2f us  frac %.3f  team %s apt %s tpc %s' % (d['ms_per_step']*1e3, d['frac_6541'], f.get('team_threads'), f.get('agents_per_thread'), f.get('teams_per_cta')))"; }
echo "== configs[1] closed loop: residency sweep (tpc, max CTAs per SM)"
for cfg in "7 4" "7 3" "7 2" "4 7" "4 5" "4 4" "4 3" "2 10" "2 8" "2 6" "1 16" "1 12"; do
  set -- $cfg; echo " tpc $1 maxcta $2"; PGM_TPC=$1 PGM_FAST_MAXCTA=$2 q --steps 1024 --graph 16
done
echo "== configs[1] closed loop: team 64"
for cfg in "7 2" "4 4" "4 3" "2 8" "2 6" "2 4" "1 8"; do
  set -- $cfg; echo " team 64 tpc $1 maxcta $2"; PGM_FAST_TEAM=64 PGM_TPC=$1 PGM_FAST_MAXCTA=$2 q --steps 1024 --graph 16
done
echo "== configs[1] many (16 steps per launch): team / tpc"
for cfg in "32 7" "32 4" "32 14" "64 7" "64 4"; do
  set -- $cfg; echo " team $1 tpc $2"; PGM_FAST_TEAM=$1 PGM_TPC=$2 q --steps 1024 --many 16
done
echo "== r=3 share (2048 instances): closed loop"
for cfg in "32 7 9" "32 7 2" "32 4 4" "32 2 7" "64 7 9" "64 7 1" "64 4 2" "64 2 4" "64 2 7"; do
  set -- $cfg; echo " team $1 tpc $2 maxcta $3"; PGM_FAST_TEAM=$1 PGM_TPC=$2 PGM_FAST_MAXCTA=$3 q --n 2048 --r 3 --steps 1024 --graph 16
done
echo "== r=3 share many"
for cfg in "32 7" "32 4" "64 7" "64 4"; do
  set -- $cfg; echo " team $1 tpc $2"; PGM_FAST_TEAM=$1 PGM_TPC=$2 q --n 2048 --r 3 --steps 1024 --many 16
done
echo "== configs[2] maze soft/restart closed loop"
for cfg in "128 7 1" "128 2 3" "128 1 7" "128 1 5" "256 1 7" "256 1 4" "64 4 3" "64 2 7"; do
  set -- $cfg; echo " team $1 tpc $2 maxcta $3"; PGM_FAST_TEAM=$1 PGM_TPC=$2 PGM_FAST_MAXCTA=$3 q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --graph 16
done
echo "== configs[2] many"
for cfg in "128 7" "128 1" "256 1" "64 7"; do
  set -- $cfg; echo " team $1 tpc $2"; PGM_FAST_TEAM=$1 PGM_TPC=$2 q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --many 16
done
echo "== configs[3] warehouse closed loop"
for cfg in "256 1 4" "256 1 3" "256 1 2"; do
  set -- $cfg; echo " team $1 tpc $2 maxcta $3"; PGM_FAST_TEAM=$1 PGM_TPC=$2 PGM_FAST_MAXCTA=$3 q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16
done
