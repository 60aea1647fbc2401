#!/bin/bash
mkdir -p gpurun_out
timeout 120 compute-sanitizer --tool racecheck python tests/sanitize_new_kernels.py > gpurun_out/r2am_sanitizer_racecheck.log 2>&1; tail -n 6 gpurun_out/r2am_sanitizer_racecheck.log | cut -c1-200
