#!/bin/bash
mkdir -p gpurun_out
for p in 0 1; do
FH_MID2_PREF=$p timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mid2 -s 3 -c 1 -o gpurun_out/r2b_mid2_pref$p -f python tests/stage_time.py > gpurun_out/r2b_ncu_pref$p.log 2>&1
done
