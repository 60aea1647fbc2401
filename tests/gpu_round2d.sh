#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_potential.py tests/test_gpu_dropin.py tests/test_gpu_slab.py -q --timeout 900 > gpurun_out/r2d_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_sizes.py -q --timeout 600 -k "golden_examples or c3 or c1b or c2_shaped" >> gpurun_out/r2d_pytest.log 2>&1
grep -n "passed\|failed" gpurun_out/r2d_pytest.log
