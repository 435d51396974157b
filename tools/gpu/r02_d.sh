#!/bin/bash
# round 2, call D: fast kernel v2 (pipelined action loads, TMA fills) - parity, sanitizer, A/B incl. stagger knob
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/d_pytest.log 2>&1
tail -4 gpurun_out/d_pytest.log
timeout 1500 bash tools/gpu/r02_b_san.sh > gpurun_out/d_san.log 2>&1; grep -c "exit=0" gpurun_out/d_san.log; grep "exit=" gpurun_out/d_san.log | grep -v "exit=0"
PGM_FAST=1 timeout 900 python tools/bench_configs.py > gpurun_out/d_configs.json 2> gpurun_out/d_configs.err
for st in 100 250 500 1000; do
  echo "stagger $st"
  PGM_STAGGER_NS=$st python tools/quick_bench.py --steps 512 --graph 16 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  c1 closed', d['ms_per_step']*1e3, d['frac_6541'])"
  PGM_STAGGER_NS=$st python tools/quick_bench.py --n 2048 --r 3 --steps 512 --graph 16 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  r3 closed', d['ms_per_step']*1e3, d['frac_6541'])"
done
python - <<'PY'
import json
for ln in open('gpurun_out/d_configs.json'):
    try: d = json.loads(ln)
    except Exception: continue
    print('  %-62s many %7.2f us %.3f | closed %7.2f us %.3f' % (d['config'][:62], d['steps_per_launch_16']['us_per_step'], d['steps_per_launch_16']['frac_of_measured_hbm'],
          d['one_launch_per_step']['us_per_step'], d['one_launch_per_step']['frac_of_measured_hbm']))
PY
python tools/phase_timeline.py > gpurun_out/d_timeline_c1.txt 2>&1; cat gpurun_out/d_timeline_c1.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/d_bench_driver.json 2> gpurun_out/d_bench_driver.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench_driver.json').read().strip().splitlines()[-1])
print('driver line: us/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'steady', round(d['roofline']['steady_state']['frac'],3),
      'closed', d['closed_loop'] and round(d['closed_loop']['roofline_frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'dram', d['host_dram']['nt_fill_GBps_all_ranks'], d['host_dram']['frac_of_ceiling'])
PY
