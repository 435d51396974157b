#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1
tail -4 gpurun_out/g_pytest.log
timeout 900 python tools/bench_configs.py > gpurun_out/g_configs.json 2> gpurun_out/g_configs.err
python - <<'PY'
import json
for ln in open('gpurun_out/g_configs.json'):
    try: d = json.loads(ln)
    except Exception: continue
    print('  %-62s many %7.2f us %.3f | closed %7.2f us %.3f | %s' % (d['config'][:62], d['steps_per_launch_16']['us_per_step'], d['steps_per_launch_16']['frac_of_measured_hbm'],
          d['one_launch_per_step']['us_per_step'], d['one_launch_per_step']['frac_of_measured_hbm'], d['plan'].get('fast')))
PY
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); f=d['plan'].get('fast',{}); print('   %.2f us  frac %.3f  team %s apt %s tpc %s' % (d['ms_per_step']*1e3, d['frac_6541'], f.get('team_threads'), f.get('agents_per_thread'), f.get('teams_per_cta')))"; }
echo "== 16384 instances r=3 / r=5: team 32 vs 64, many + closed"
for t in 32 64; do for r in 3 5; do echo " team $t r $r"; PGM_FAST_TEAM=$t q --n 16384 --r $r --steps 256 --many 16; PGM_FAST_TEAM=$t q --n 16384 --r $r --steps 256 --graph 16; done; done
echo "== configs[1] tpc sweep team 64"
for tpc in 2 3 4 5 6 8; do echo " tpc $tpc"; PGM_TPC=$tpc q --steps 1024 --many 16; PGM_TPC=$tpc q --steps 1024 --graph 16; done
for k in 20; do
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/g_bench_driver.json 2> gpurun_out/g_bench_driver.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g_bench_driver.json').read().strip().splitlines()[-1])
print('driver line: us/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'steady', round(d['roofline']['steady_state']['frac'],3),
      'closed', d['closed_loop'] and round(d['closed_loop']['roofline_frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'dram', d['host_dram']['nt_fill_GBps_all_ranks'], d['host_dram']['frac_of_ceiling'])
PY
done
