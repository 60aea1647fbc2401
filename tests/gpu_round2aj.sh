#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2aj_slab_check1.log 2>&1
grep -c " ok" gpurun_out/r2aj_slab_check1.log; grep "FAIL\|Error\|error" gpurun_out/r2aj_slab_check1.log | head -5; grep "1024" gpurun_out/r2aj_slab_check1.log | cut -c1-150
