#!/bin/bash
mkdir -p gpurun_out
for k in k_fwd_last_reg3 k_inv_last_reg3; do
  BN=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2ab_$k python tests/stage_time.py > gpurun_out/r2ab_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2ab*.ncu-rep
