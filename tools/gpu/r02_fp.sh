#!/bin/bash
# is the ~5 % loss of 32 / 64 steps per launch the drift of the per-instance timelines, or the memory footprint of the
# longer rollout's action / reward / flag tensors?  16 steps per launch over 1 / 4 / 8 sets of those tensors.
q() { python tools/quick_bench.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   %.2f us  frac %.3f' % (d['ms_per_step']*1e3, d['frac_6541']))"; }
for rep in 1 2; do
echo " c1: many 16 | 16 x 1 set | 16 x 4 sets | 16 x 8 sets | many 64 | many 64 x 1 set"
q --steps 2048 --many 16; q --steps 2048 --many 16 --sets 1; q --steps 2048 --many 16 --sets 4; q --steps 2048 --many 16 --sets 8; q --steps 2048 --many 64; q --steps 2048 --many 64 --sets 1
echo " c1 many 16 / 64 with a ring of 8 observation buffers"
q --steps 2048 --many 16 --ring 8; q --steps 2048 --many 64 --ring 8
done
