#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ae_stage_tlb.log
: > $L
BN=512 BN1=64 timeout 300 python tests/stage_time.py >> $L 2>&1
BN=512 BN2=64 timeout 300 python tests/stage_time.py >> $L 2>&1
BN=512 BN0=64 timeout 300 python tests/stage_time.py >> $L 2>&1
BN=256 BN1=1024 timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-330 $L
