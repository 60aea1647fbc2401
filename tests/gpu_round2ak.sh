#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
SLAB_X=p2p timeout 600 $TR tests/slab_check.py --notest --time-scalar 1024 --profile > gpurun_out/r2ak_config5_1024.log 2>&1
grep "mode\|profile\|Error\|error" gpurun_out/r2ak_config5_1024.log | cut -c1-330 | head
