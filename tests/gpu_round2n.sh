#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
for J in 3 4 2; do
FH_KBLOCKS=$J SLAB_X=kblock timeout 600 $TR tests/slab_check.py --notest --time 512 > gpurun_out/r2n_kblock_J$J.log 2>&1
grep "^mode" gpurun_out/r2n_kblock_J$J.log | cut -c1-200
done
FH_SLAB_GRAPH=0 FH_KBLOCKS=3 SLAB_X=kblock timeout 600 $TR tests/slab_check.py --notest --time 512 > gpurun_out/r2n_kblock_J3_nograph.log 2>&1
grep "^mode" gpurun_out/r2n_kblock_J3_nograph.log | cut -c1-200
timeout 900 $TR bench.py --gpus 8 > gpurun_out/r2n_bench8.json 2> gpurun_out/r2n_bench8.err
cut -c1-300 gpurun_out/r2n_bench8.json
