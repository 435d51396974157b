#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/w_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/w_pytest.log | tail -2
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; tail -3 gpurun_out/w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench.json').read().strip().splitlines()[-1])
print('value %.2f G, us/step %.2f, frac %.3f, steady %.3f, closed %.3f, e2e %.1f M, kernel %s' % (d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['steady_state']['frac'], d['closed_loop']['roofline_frac'], d['e2e']['value']/1e6, d['roofline']['kernel']))
for g in d['closed_loop'].get('groups', []): print('  groups', g['groups'], round(g['us_per_step'],2), round(g['roofline_frac'],3))
for c in d['configs']:
    print('  ', c['config'][:40], {k: (round(v['us_per_step'],2), round(v['roofline_frac'],3)) for k,v in c.items() if isinstance(v,dict) and 'us_per_step' in v})
PY
