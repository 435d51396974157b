#!/bin/bash
# parity suite + closed loop with several groups of instances (tools/two_groups.py)
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/v_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/v_pytest.log | tail -2
echo "== closed loop, G groups (one launch per step and group)"
t() { python tools/two_groups.py "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   groups %d x %d: %.2f us per step of all groups, frac %.3f  %s' % (d['groups'], d['instances_per_group'], d['ms_per_step_of_all_groups']*1e3, d['frac_6541'], d['plan']))"; }
for rep in 1 2; do
echo " c1"; for g in 1 2 4; do t --groups $g; done
echo " c2"; for g in 1 2; do t --groups $g --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512; done
echo " c3"; for g in 1 2; do t --groups $g --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256; done
echo " r3 share"; for g in 1 2; do t --groups $g --n 2048 --r 3 --steps 1024; done
done
