#!/bin/bash
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/q_pytest.log 2>&1; grep -E "passed|failed|error" gpurun_out/q_pytest.log | tail -2
bash tools/make_profiles_r02.sh > gpurun_out/q_profiles.log 2>&1
tail -3 gpurun_out/q_profiles.log
