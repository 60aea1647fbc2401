#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ai_stage1024.log
: > $L
for v in 1 0; do
  FH_REG3_1024=$v BN0=128 BN1=1024 BN2=1024 BD=3 timeout 300 python tests/stage_time.py >> $L 2>&1
done
cut -c1-360 $L
