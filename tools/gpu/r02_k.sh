#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/k_pytest.log 2>&1
tail -4 gpurun_out/k_pytest.log
timeout 1200 bash tools/gpu/r02_b_san.sh > gpurun_out/k_san.log 2>&1; grep -c "exit=0" gpurun_out/k_san.log; grep "exit=" gpurun_out/k_san.log | grep -v "exit=0"
timeout 900 python tools/bench_configs.py > gpurun_out/k_configs.json 2> gpurun_out/k_configs.err
python - <<'PY'
import json
for ln in open('gpurun_out/k_configs.json'):
    try: d = json.loads(ln)
    except Exception: continue
    print('  %-62s many %7.2f us %.3f | closed %7.2f us %.3f | %s' % (d['config'][:62], d['steps_per_launch_16']['us_per_step'], d['steps_per_launch_16']['frac_of_measured_hbm'],
          d['one_launch_per_step']['us_per_step'], d['one_launch_per_step']['frac_of_measured_hbm'], d['plan'].get('fast')))
PY
