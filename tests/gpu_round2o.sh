#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2o_slab_check.log 2>&1
grep -c " ok" gpurun_out/r2o_slab_check.log; grep "FAIL\|Error\|error" gpurun_out/r2o_slab_check.log | head -5
FH_KBLOCKS=2 SLAB_X=kblock timeout 900 $TR tests/slab_check.py --notest --time 512 > gpurun_out/r2o_kblock.log 2>&1
grep "^mode" gpurun_out/r2o_kblock.log | cut -c1-200
