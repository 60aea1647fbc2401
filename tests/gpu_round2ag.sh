#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2ag_bench_8gpu.json 2> gpurun_out/r2ag_bench_8gpu.err
cut -c1-400 gpurun_out/r2ag_bench_8gpu.json; tail -3 gpurun_out/r2ag_bench_8gpu.err
