#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2z_stage256.log
: > $L
FH_MID2=2 timeout 300 python tests/stage_time.py >> $L 2>&1
timeout 300 python tests/stage_time.py >> $L 2>&1
FH_MID2=2 BD=3 timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-330 $L
timeout 900 python -m pytest tests/test_gpu_round2.py -q --timeout 900 -x -k "mid2" > gpurun_out/r2z_pytest.log 2>&1; tail -n 3 gpurun_out/r2z_pytest.log
