#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2f_slab_check.log 2>&1
grep -c " ok" gpurun_out/r2f_slab_check.log; grep "FAIL\|Error\|error" gpurun_out/r2f_slab_check.log | head -5
for J in 1 2 4 8; do
FH_PUSH_CHUNKS=$J SLAB_X=push timeout 900 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2f_slab_time512_J$J.log 2>&1
grep "^mode\|^profile" gpurun_out/r2f_slab_time512_J$J.log
done
timeout 600 python -m pytest tests/test_gpu_potential.py -q --timeout 600 > gpurun_out/r2f_pytest_potential.log 2>&1; tail -n 2 gpurun_out/r2f_pytest_potential.log
