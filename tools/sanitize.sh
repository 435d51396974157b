#!/bin/bash
# compute-sanitizer passes over small runs of every collision system / on_target mode (GPU box).
set -x
for tool in memcheck racecheck synccheck initcheck; do
  for mode in "priority finish" "block_both nothing" "soft restart"; do
    set -- $mode
    compute-sanitizer --tool $tool --error-exitcode 9 python tools/quick_bench.py --n 48 --size 12 --agents 20 --r 3 \
        --coll $1 --ot $2 --steps 6 --max-steps 5 > gpurun_out/san_${tool}_$1.log 2>&1
    echo "$tool $1 $2 exit=$?"
    tail -3 gpurun_out/san_${tool}_$1.log
  done
done
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_bench.py --n 48 --size 12 --agents 20 --r 4 --coll soft --ot restart --steps 32 --many 8 --max-steps 5 > gpurun_out/san_memcheck_many.log 2>&1; echo "memcheck many exit=$?"; tail -2 gpurun_out/san_memcheck_many.log
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_bench.py --n 6 --size 300 --agents 40 --r 5 --coll soft --ot restart --steps 4 --max-steps 5 > gpurun_out/san_memcheck_buckets.log 2>&1; echo "memcheck tile-buckets exit=$?"; tail -2 gpurun_out/san_memcheck_buckets.log
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/quick_bench.py --n 6 --size 300 --agents 40 --r 5 --coll priority --ot finish --steps 4 --max-steps 5 > gpurun_out/san_racecheck_buckets.log 2>&1; echo "racecheck tile-buckets exit=$?"; tail -2 gpurun_out/san_racecheck_buckets.log
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/quick_bench.py --n 24 --size 40 --agents 300 --r 5 --coll priority --ot finish --steps 4 --max-steps 5 > gpurun_out/san_racecheck_team.log 2>&1; echo "racecheck big-team exit=$?"; tail -2 gpurun_out/san_racecheck_team.log
# single-step launches with two observation batches (64 agents on a warp), and the packed host transport (obs_format 3 + flag copies)
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/quick_bench.py --n 24 --size 16 --agents 64 --r 3 --steps 6 --max-steps 5 > gpurun_out/san_racecheck_twobatch.log 2>&1; echo "racecheck two-batch exit=$?"; tail -2 gpurun_out/san_racecheck_twobatch.log
for tool in memcheck racecheck initcheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python tools/e2e_probe.py --n 40 --size 12 --agents 20 --r 3 --steps 6 --threads 3 > gpurun_out/san_${tool}_packed.log 2>&1; echo "$tool packed exit=$?"; tail -2 gpurun_out/san_${tool}_packed.log
done
