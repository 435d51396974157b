#!/bin/bash
mkdir -p gpurun_out
exec 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest.log
for t in 0 12 14 15 16; do python tools/e2e_probe.py --threads $t; done
PGM_HOST_SPIN_US=100 python tools/e2e_probe.py --threads 16
PGM_HOST_SPIN_US=100 python tools/e2e_probe.py --threads 15
for c in 4 16; do PGM_STREAM_CHUNKS=$c python tools/e2e_probe.py --threads 16; done
python tools/e2e_probe.py --threads 16 --pageable
python tools/e2e_probe.py --threads 16 --fmt f32
python tools/quick_bench.py --graph 16 --steps 2048
python tools/quick_bench.py --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --graph 16 --steps 1024
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
