#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2t_stage_times.log
: > $L
timeout 300 python tests/stage_time.py >> $L 2>&1
BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1
BN=128 timeout 300 python tests/stage_time.py >> $L 2>&1
BD=3 timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-420 $L
timeout 1200 python -m pytest tests/test_gpu_bench_sizes.py tests/test_gpu_parity.py -q --timeout 900 -x > gpurun_out/r2t_pytest.log 2>&1; tail -n 3 gpurun_out/r2t_pytest.log
