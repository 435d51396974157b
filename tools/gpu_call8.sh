#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest.log
run() { (cd $1 && python tools/quick_bench.py $2 | grep -o '"ms_per_step.*'); }
for i in 1 2; do for v in _oldtree .; do echo "$v many"; run $v "--many 16 --steps 4096"; done; done
python tools/e2e_probe.py --threads 16
bash tools/make_profiles.sh > gpurun_out/make_profiles.log 2>&1
tail -3 gpurun_out/make_profiles.log
python tools/bench_configs.py > gpurun_out/configs.json 2>&1; cat gpurun_out/configs.json | cut -c1-100
