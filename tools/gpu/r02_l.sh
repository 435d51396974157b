#!/bin/bash
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['plan'].get('fast',{}); print('   %.2f us  frac %.3f  team %s apt %s tpc %s' % (d['ms_per_step']*1e3, d['frac_6541'], f.get('team_threads'), f.get('agents_per_thread'), f.get('teams_per_cta')))"; }
echo "== is the store phase bound by HBM or by the path into L2?  configs[1], observation ring of 4 buffers (380 MB) vs 1 (95 MB, stays in the 126 MB L2)"
for ring in 4 1; do echo " ring $ring closed / many"; q --steps 1024 --graph 16 --ring $ring; q --steps 1024 --many 16 --ring $ring; done
echo "== r=3 share (9.4 MB per buffer): ring 4 / 1 / 16"
for ring in 4 1; do echo " ring $ring closed"; q --n 2048 --r 3 --steps 1024 --graph 16 --ring $ring; done
echo "== configs[3]: ring 4 vs 1 (190 MB > L2 either way) closed"
for ring in 4 1; do echo " ring $ring closed"; q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16 --ring $ring; done
echo "== front only (no observations), closed loop: c1, c2, c3, r3"
q --steps 1024 --graph 16 --noobs
q --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512 --graph 16 --noobs
q --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256 --graph 16 --noobs
q --n 2048 --r 3 --steps 1024 --graph 16 --noobs
echo "== fuzz with 400 cases"
PGM_FUZZ_CASES=400 timeout 1500 python -m pytest tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
