#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2l_slab_check.log 2>&1
grep -c " ok" gpurun_out/r2l_slab_check.log; grep "FAIL\|Error\|error" gpurun_out/r2l_slab_check.log | head -5
for J in 3 6; do
FH_KBLOCKS=$J SLAB_X=kblock timeout 900 $TR tests/slab_check.py --notest --time 512 > gpurun_out/r2l_kblock_J$J.log 2>&1
grep "^mode" gpurun_out/r2l_kblock_J$J.log | cut -c1-200
done
FH_SLAB_GRAPH=0 FH_KBLOCKS=3 SLAB_X=kblock timeout 900 $TR tests/slab_check.py --notest --time 512 > gpurun_out/r2l_kblock_nograph.log 2>&1
grep "^mode" gpurun_out/r2l_kblock_nograph.log | cut -c1-200
SLAB_X=p2p,push timeout 900 $TR tests/slab_check.py --notest --time 512 > gpurun_out/r2l_p2p_push_graph.log 2>&1
grep "^mode" gpurun_out/r2l_p2p_push_graph.log | cut -c1-200
