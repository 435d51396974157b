#!/bin/bash
mkdir -p gpurun_out
for st in 0 400 1000 2000; do
  echo "=== stagger $st"
  PGM_STAGGER_NS=$st python tools/phase_timeline.py 2>&1 | tail -16
done
echo "=== no observations (front only), closed loop"
python tools/quick_bench.py --steps 512 --graph 16 --noobs
echo "=== bits format (48 B/agent), closed loop"
python tools/quick_bench.py --steps 512 --graph 16 --fmt bits
echo "=== tpc 8 / tpc 4"
PGM_TPC=8 python tools/quick_bench.py --steps 512 --graph 16
PGM_TPC=4 python tools/quick_bench.py --steps 512 --graph 16
PGM_TPC=14 python tools/quick_bench.py --steps 512 --graph 16
PGM_FAST_TEAM=64 python tools/quick_bench.py --steps 512 --graph 16
