#!/bin/bash
# what the driver runs at round end, in one call: parity suite, smoke(), both bench arms
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/final_pytest.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -3 gpurun_out/final_ref.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -3 gpurun_out/final_bench.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/final_ref.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('same config:', r['config'] == d['config'])
print('value %.2f G, us/step %.2f, frac %.3f, steady %.3f, closed %.3f, e2e %.1f M (ratio vs reference arm %.0fx), launches %s, clocks %s' % (d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['steady_state']['frac'], d['closed_loop']['roofline_frac'], d['e2e']['value']/1e6, d['e2e']['value']/r['value'], d['gpu_launches'], d['clocks']))
print('sharding', d['sharding_check']['status'], 'host_dram', round(d['host_dram']['frac_of_ceiling'],2), 'cpu_baseline', d.get('cpu_baseline',{}).get('value'))
PY
