#!/bin/bash
# round 2, call C: where does the fast kernel spend its time?  phase timelines + ncu --set full captures
mkdir -p gpurun_out
python tools/phase_timeline.py > gpurun_out/c_timeline_c1.txt 2>&1; cat gpurun_out/c_timeline_c1.txt
python tools/phase_timeline.py --n 2048 --r 3 > gpurun_out/c_timeline_r3.txt 2>&1; cat gpurun_out/c_timeline_r3.txt
NCU="ncu --set full --clock-control none --import-source on -k regex:pgm_fast"
# one single-step launch of configs[1] (closed-loop form): 20 warm-up steps in quick_bench, then the timed ones
$NCU -s 30 -c 1 -o gpurun_out/c_prof_single python tools/quick_bench.py --steps 32 > gpurun_out/c_ncu_single.log 2>&1
# one 16-step launch of configs[1]
$NCU -s 22 -c 1 -o gpurun_out/c_prof_many python tools/quick_bench.py --steps 64 --many 16 > gpurun_out/c_ncu_many.log 2>&1
# r=3 share, 16-step launch
$NCU -s 22 -c 1 -o gpurun_out/c_prof_r3 python tools/quick_bench.py --n 2048 --r 3 --steps 64 --many 16 > gpurun_out/c_ncu_r3.log 2>&1
# configs[3] warehouse, 16-step launch
$NCU -s 22 -c 1 -o gpurun_out/c_prof_wh python tools/quick_bench.py --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 32 --many 16 > gpurun_out/c_ncu_wh.log 2>&1
tail -2 gpurun_out/c_ncu_*.log
ls -la gpurun_out/c_*
