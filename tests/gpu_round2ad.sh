#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ad_stage512.log
: > $L
BN=512 timeout 300 python tests/stage_time.py >> $L 2>&1
BN=512 BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-330 $L
BN=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mid_green_512 -s 3 -c 1 -f -o gpurun_out/r2ad_k_mid_green_512 python tests/stage_time.py > gpurun_out/r2ad_ncu_mid512.log 2>&1
ls -la gpurun_out/r2ad*.ncu-rep
