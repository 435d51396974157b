#!/bin/bash
# compute-sanitizer over the fast step kernel: every collision system, single / multi-step launches, APT 1/2/4
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for mode in "priority finish 20 32" "block_both nothing 48 32" "soft restart 100 32 30" "priority restart 128 64 30"; do
    set -- $mode
    PGM_FAST_TEAM=$4 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/quick_bench.py --n 48 --size ${5:-14} --agents $3 --r 3 \
        --coll $1 --ot $2 --steps 6 --max-steps 5 > gpurun_out/b_san_${tool}_$1_$3.log 2>&1
    echo "$tool $1 $2 A=$3 exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|frac_6541" gpurun_out/b_san_${tool}_$1_$3.log | cut -c1-200
    PGM_FAST_TEAM=$4 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/quick_bench.py --n 48 --size ${5:-14} --agents $3 --r 5 \
        --coll $1 --ot $2 --steps 16 --many 8 --max-steps 5 > gpurun_out/b_san_${tool}_$1_$3_many.log 2>&1
    echo "$tool $1 $2 A=$3 many exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/b_san_${tool}_$1_$3_many.log | cut -c1-200
  done
done
