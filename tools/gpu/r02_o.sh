#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/o_pytest.log | tail -3
PGM_FUZZ_CASES=400 timeout 1500 python -m pytest tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['plan'].get('fast',{}); print('   %.2f us  frac %.3f  team %s apt %s tpc %s' % (d['ms_per_step']*1e3, d['frac_6541'], f.get('team_threads'), f.get('agents_per_thread'), f.get('teams_per_cta')))"; }
echo "== agent counts off 16: 4096 x 60 agents / 100 agents (fast vs generic), many + closed"
for f in 1 0; do echo " PGM_FAST=$f"; PGM_FAST=$f q --agents 60 --steps 512 --many 16; PGM_FAST=$f q --agents 60 --steps 512 --graph 16; PGM_FAST=$f q --agents 100 --size 40 --n 2048 --steps 512 --many 16; PGM_FAST=$f q --agents 100 --size 40 --n 2048 --steps 512 --graph 16; done
echo "== configs[1] regression check"; q --steps 1024 --many 16; q --steps 1024 --graph 16
timeout 600 bash tools/gpu/r02_b_san.sh > gpurun_out/o_san.log 2>&1; grep -c "exit=0" gpurun_out/o_san.log; grep "exit=" gpurun_out/o_san.log | grep -v "exit=0"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_bench.py --n 48 --size 14 --agents 21 --r 3 --steps 6 --max-steps 5 2>&1 | grep -E "ERROR SUMMARY"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_bench.py --n 48 --size 14 --agents 21 --r 5 --steps 16 --many 8 --max-steps 5 2>&1 | grep -E "ERROR SUMMARY"
