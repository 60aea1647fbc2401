#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2ac_stage512.log
: > $L
BN=512 timeout 300 python tests/stage_time.py >> $L 2>&1
BN=512 BD=3 timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-330 $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_sizes.py -q --timeout 900 -x -k "512" > gpurun_out/r2ac_pytest512.log 2>&1; tail -n 2 gpurun_out/r2ac_pytest512.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2ac_slab_check1.log 2>&1
grep -c " ok" gpurun_out/r2ac_slab_check1.log; grep "FAIL\|Error\|error" gpurun_out/r2ac_slab_check1.log | head -5
