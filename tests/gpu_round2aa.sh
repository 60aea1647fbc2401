#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR tests/slab_check.py > gpurun_out/r2aa_slab_check2.log 2>&1
grep -c " ok" gpurun_out/r2aa_slab_check2.log; grep "FAIL\|Error\|error" gpurun_out/r2aa_slab_check2.log | head -5; grep "512" gpurun_out/r2aa_slab_check2.log | cut -c1-150 | head -12
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2aa_bench_2gpu.json 2> gpurun_out/r2aa_bench_2gpu.err
cut -c1-600 gpurun_out/r2aa_bench_2gpu.json; tail -3 gpurun_out/r2aa_bench_2gpu.err
