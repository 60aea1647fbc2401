#!/bin/bash
# 2-GPU session: push-mode exchange, correctness + timing at 512^3
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2e_slab_check.log 2>&1
grep -c " ok" gpurun_out/r2e_slab_check.log; grep "FAIL\|Error\|error" gpurun_out/r2e_slab_check.log | head -5
SLAB_X=push,p2p timeout 900 $TR tests/slab_check.py --notest --time 512 --stages --profile > gpurun_out/r2e_slab_time512.log 2>&1
grep "^mode\|^stages\|^profile" gpurun_out/r2e_slab_time512.log
timeout 600 python -m pytest tests/test_gpu_potential.py -q --timeout 600 > gpurun_out/r2e_pytest_potential.log 2>&1; tail -n 2 gpurun_out/r2e_pytest_potential.log
