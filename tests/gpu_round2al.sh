#!/bin/bash
mkdir -p gpurun_out
timeout 140 compute-sanitizer --tool memcheck python tests/sanitize_new_kernels.py > gpurun_out/r2al_sanitizer_memcheck.log 2>&1; tail -n 12 gpurun_out/r2al_sanitizer_memcheck.log | cut -c1-200
