#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2j_slab_check.log 2>&1
grep -c " ok" gpurun_out/r2j_slab_check.log; grep "FAIL\|Error\|error" gpurun_out/r2j_slab_check.log | head -5
for J in 3 4 6; do
FH_KBLOCKS=$J SLAB_X=kblock timeout 900 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2j_kblock_J$J.log 2>&1
grep "^mode\|^profile" gpurun_out/r2j_kblock_J$J.log | cut -c1-300
done
