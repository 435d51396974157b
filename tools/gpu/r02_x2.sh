#!/bin/bash
# the driver's multi-GPU invocation at N=2 (torchrun), both arms
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/x2_ref.json 2> gpurun_out/x2_ref.err; tail -2 gpurun_out/x2_ref.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/x2_bench.json 2> gpurun_out/x2_bench.err; tail -4 gpurun_out/x2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x2_bench.json').read().strip().splitlines()[-1])
print('n_gpus %d value %.2f G, us/step %.2f, frac %.3f, closed %.3f, groups %s, e2e %.1f M, e2e_bits %.1f M, sharding %s' % (d['n_gpus'], d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['frac'], d['closed_loop']['roofline_frac'], [(g['groups'], round(g['roofline_frac'],3)) for g in d['closed_loop']['groups']], d['e2e']['value']/1e6, d['e2e_bits']['value']/1e6, d['sharding_check']['status']))
for c in d['configs']:
    print('  ', c['config'][:50], {k: (round(v['us_per_step'],2), round(v['roofline_frac'],3)) for k,v in c.items() if isinstance(v,dict) and 'us_per_step' in v})
PY
