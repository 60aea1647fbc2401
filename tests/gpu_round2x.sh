#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2x_stage255.log
: > $L
for envs in "FH_ODD_V=4" "FH_ODD_V=3" "FH_ODD_V=2" "FH_ODD_TRW=4"; do
  env $envs BN=255 BD=3 BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1
done
env BN=255 BD=6 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-400 $L
timeout 1500 python -m pytest tests/test_gpu_odd.py -q --timeout 900 -x > gpurun_out/r2x_pytest_odd.log 2>&1; tail -n 5 gpurun_out/r2x_pytest_odd.log
