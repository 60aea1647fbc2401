#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2y_stage255.log
: > $L
env BN=255 BD=3 BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1
env BN=255 BD=3 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
env BN=255 BD=6 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-400 $L
timeout 1500 python -m pytest tests/test_gpu_odd.py tests/test_gpu_reference_suite.py -q --timeout 900 -x > gpurun_out/r2y_pytest_a.log 2>&1; tail -n 5 gpurun_out/r2y_pytest_a.log
timeout 1500 python -m pytest tests/test_gpu_bench_sizes.py tests/test_gpu_parity.py -q --timeout 900 -x -k "255 or deferred or odd or 63 or 45" > gpurun_out/r2y_pytest_b.log 2>&1; tail -n 3 gpurun_out/r2y_pytest_b.log
