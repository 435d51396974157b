#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/final2_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/final2_pytest.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/make_profiles_r02.sh > gpurun_out/final2_profiles.log 2>&1
tail -3 gpurun_out/final2_profiles.log
