#!/bin/bash
# 8 GPUs: the bench line under torchrun (sharding check, per-config records, e2e on a shared host)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/i_bench_8gpu.json 2> gpurun_out/i_bench_8gpu.err
tail -3 gpurun_out/i_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/i_bench_8gpu.json').read().strip().splitlines()[-1])
print('8 GPUs: value', d['value']/1e9, 'us/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'steady', round(d['roofline']['steady_state']['frac'],3),
      'closed', d['closed_loop'] and round(d['closed_loop']['roofline_frac'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'bits', round(d['e2e_bits']['value']/1e6,1), 'shard', d['sharding_check'], 'dram', d['host_dram'])
for c in d.get('configs') or []:
    print('   ', c['config'][:66], {k: (round(v['us_per_step'],2), round(v['roofline_frac'],3)) for k,v in c.items() if isinstance(v, dict) and 'us_per_step' in v})
PY
