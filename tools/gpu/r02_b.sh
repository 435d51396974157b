#!/bin/bash
# round 2, call B: parity suite with the fast step kernel + A/B timings of all configurations (fast on / off)
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
tail -15 gpurun_out/b_pytest.log
timeout 1500 bash tools/gpu/r02_b_san.sh > gpurun_out/b_san.log 2>&1; grep -c "exit=0" gpurun_out/b_san.log; grep "exit=" gpurun_out/b_san.log | grep -v "exit=0"
for f in 1 0; do
  PGM_FAST=$f timeout 900 python tools/bench_configs.py > gpurun_out/b_configs_fast$f.json 2> gpurun_out/b_configs_fast$f.err
  tail -2 gpurun_out/b_configs_fast$f.err
done
python - <<'PY'
import json
for f in (1, 0):
    print('PGM_FAST =', f)
    for ln in open(f'gpurun_out/b_configs_fast{f}.json'):
        try: d = json.loads(ln)
        except Exception: continue
        print('  %-62s many %7.2f us %.3f | closed %7.2f us %.3f | %s' % (d['config'][:62], d['steps_per_launch_16']['us_per_step'], d['steps_per_launch_16']['frac_of_measured_hbm'],
              d['one_launch_per_step']['us_per_step'], d['one_launch_per_step']['frac_of_measured_hbm'], d['plan'].get('fast', d['plan'])))
PY
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_driver.json 2> gpurun_out/b_bench_driver.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_bench_driver.json').read().strip().splitlines()[-1])
print('driver line: us/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'steady', round(d['roofline']['steady_state']['frac'],3),
      'closed', d['closed_loop'] and round(d['closed_loop']['roofline_frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'shard', d['sharding_check'], 'dram', d['host_dram'])
for c in d.get('configs') or []:
    print('   ', c['config'][:60], {k: (round(v['us_per_step'],2), round(v['roofline_frac'],3)) for k,v in c.items() if isinstance(v, dict) and 'us_per_step' in v})
PY
