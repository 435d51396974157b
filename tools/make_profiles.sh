#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one
# `ncu --set full` capture of the step kernel, the phase timeline.  Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 256 --warmup 8 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 256 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-graph > gpurun_out/bench_under_ncu.log 2>&1
# launch #18 of pgm_step_kernel = the second 16-step launch (16 warm-up single steps + 1 multi-step launch skipped)
ncu --set full --clock-control none --import-source on -k regex:pgm_step_kernel -s 17 -c 1 -o gpurun_out/prof_step \
    python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-graph > gpurun_out/bench_under_ncu2.log 2>&1
python tools/phase_timeline.py > gpurun_out/phase_timeline.txt 2>&1
ls -la gpurun_out
