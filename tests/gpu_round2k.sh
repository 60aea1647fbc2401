#!/bin/bash
# 8-GPU session: exchange modes at 512^3 elasticity, the bench line, config 5 (1024^3 scalar)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tests/slab_check.py > gpurun_out/r2k_slab_check8.log 2>&1
grep -c " ok" gpurun_out/r2k_slab_check8.log; grep "FAIL\|Error\|error" gpurun_out/r2k_slab_check8.log | head -5
FH_KBLOCKS=3 SLAB_X=p2p,push,kblock timeout 900 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2k_time512.log 2>&1
grep "^mode\|^profile" gpurun_out/r2k_time512.log | cut -c1-300
FH_KBLOCKS=4 SLAB_X=kblock timeout 600 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2k_time512_k4.log 2>&1
grep "^mode\|^profile" gpurun_out/r2k_time512_k4.log | cut -c1-300
FH_KBLOCKS=6 SLAB_X=kblock timeout 600 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2k_time512_k6.log 2>&1
grep "^mode\|^profile" gpurun_out/r2k_time512_k6.log | cut -c1-300
timeout 900 $TR bench.py --gpus 8 > gpurun_out/r2k_bench8.json 2> gpurun_out/r2k_bench8.err
cut -c1-400 gpurun_out/r2k_bench8.json
SLAB_X=p2p timeout 900 $TR tests/slab_check.py --notest --time-scalar 1024 --profile > gpurun_out/r2k_config5_1024.log 2>&1
grep "^mode\|^profile" gpurun_out/r2k_config5_1024.log | cut -c1-300
