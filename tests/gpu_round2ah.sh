#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ah_stage1024.log
: > $L
for v in 1 0; do
  FH_REG3_1024=$v BN0=32 BN1=64 BN2=1024 BD=3 timeout 300 python tests/stage_time.py >> $L 2>&1
  FH_REG3_1024=$v BN0=16 BN1=64 BN2=1024 BD=6 timeout 300 python tests/stage_time.py >> $L 2>&1
done
cut -c1-360 $L
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 900 -x -k "axis_length_512" > gpurun_out/r2ah_pytest.log 2>&1; tail -n 3 gpurun_out/r2ah_pytest.log
