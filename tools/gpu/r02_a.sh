#!/bin/bash
# round 2, call A: parity suite + the new bench line in the driver's own invocation and launch-split variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -5 gpurun_out/a_pytest.log
for spl in 16 20 10; do
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --steps-per-launch $spl --no-cpu-baseline --no-configs > gpurun_out/a_bench_k20_spl$spl.json 2> gpurun_out/a_bench_k20_spl$spl.err
done
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/a_bench_driver.json 2> gpurun_out/a_bench_driver.err
tail -3 gpurun_out/a_bench_driver.err
timeout 600 python bench.py --no-cpu-baseline --no-configs > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unparsed', e); continue
    print(f, 'us/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'steady', round(d['roofline']['steady_state']['frac'],3),
          'closed', d['closed_loop'] and round(d['closed_loop']['roofline_frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'bits', round(d['e2e_bits']['value']/1e6,1),
          'shard', d['sharding_check'] and d['sharding_check']['status'], 'dram', d['host_dram']['nt_fill_GBps_all_ranks'], d['host_dram']['frac_of_ceiling'])
    for c in d.get('configs') or []:
        print('   ', c['config'][:60], {k: (round(v['us_per_step'],2), round(v['roofline_frac'],3)) for k,v in c.items() if isinstance(v, dict) and 'us_per_step' in v})
PY
