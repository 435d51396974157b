#!/bin/bash
# GPU session 1 of this continuation: full GPU test suite, bench line, knob experiments.
mkdir -p gpurun_out
exec 2>&1
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)|L3" ; free -g | head -2
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
g++ -O3 -std=c++17 -Ipogema_b200/csrc tools/prototypes/hostexpand_bench.cpp pogema_b200/csrc/pgm_hostexpand.cpp -lpthread -o /tmp/heb
for t in 4 8 16 32 64; do /tmp/heb $t 1; done; /tmp/heb 16 4
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
echo "== closed loop, obs batch"
for b in 64 32 16; do PGM_OBS_BATCH=$b python tools/quick_bench.py --graph 16 --steps 2048; done
echo "== 16 steps per launch, obs batch"
for b in 64 32; do PGM_OBS_BATCH=$b python tools/quick_bench.py --many 16 --steps 2048; done
echo "== r=3 N=2048 team"
for t in 32 64; do python tools/quick_bench.py --n 2048 --r 3 --team $t --many 16 --steps 4096; python tools/quick_bench.py --n 2048 --r 3 --team $t --graph 16 --steps 4096; done
echo "== N=2048 r=5 team"
for t in 32 64; do python tools/quick_bench.py --n 2048 --r 5 --team $t --graph 16 --steps 4096; done
