#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_odd.py -q --timeout 100 -x -k "download or cube" > gpurun_out/r2an_pytest.log 2>&1; tail -n 2 gpurun_out/r2an_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
